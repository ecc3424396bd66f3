"""Runs tests/test_gemm_gpu.py N times in ONE process (timing-dependent races show up as intermittent failures; this is
how the output-staging slot reuse of the skinny-K TMA epilogue at N = 144 was found):  python tools/stress_gemm_tests.py [N]"""
import sys, pytest
bad = 0
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 10):
    rc = pytest.main(["tests/test_gemm_gpu.py", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider"])
    print("RUN", i, "rc", rc, flush=True)
    bad += rc != 0
print("BAD", bad)
