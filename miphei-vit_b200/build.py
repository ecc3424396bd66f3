"""Builds libmiphei_b200.so (all CUDA kernels + the C ABI) in-tree with nvcc for sm_100a.

Run as a script or call build(); objects are cached under csrc/build/ and rebuilt when a source or header changes.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libmiphei_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _newest_header():
    t = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".h", ".cuh")):
                t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def build(verbose=False, force=False, prof=False, variant=None, defines=()):
    """prof=True builds the diagnostic twin libmiphei_b200_prof.so (-DMV_GEMM_PROFILE=1: per-role cycle counters in the
    GEMM kernel, tools/gemm_roles.py); the production library carries no instrumentation."""
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    bdir = os.path.join(CSRC, "build_prof" if prof else "build")
    flags = FLAGS + (["-DMV_GEMM_PROFILE=1"] if prof else [])
    out = OUT.replace(".so", "_prof.so") if prof else OUT
    if variant:  # experiment builds (A/B of compile-time constants): libmiphei_b200_<variant>.so, selected with MIPHEI_B200_LIB
        bdir = os.path.join(CSRC, "build_" + variant)
        flags = FLAGS + ["-D" + d for d in defines]
        out = OUT.replace(".so", "_%s.so" % variant)
    os.makedirs(bdir, exist_ok=True)
    hdr_t = _newest_header()
    jobs = []
    objs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(bdir, s[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [NVCC] + flags + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(obj + ".log", "w") as f:
            f.write(log)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, log))
        return src, log

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, log in ex.map(compile_one, jobs):
                if verbose:
                    print("== " + os.path.basename(src))
                    print(log)
    if jobs or not os.path.exists(out):
        cmd = [NVCC, "-shared", "-o", out] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return out


if __name__ == "__main__":
    var = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")]
    defs = [a[2:] for a in sys.argv if a.startswith("-D")]
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv, prof="--prof" in sys.argv, variant=var[0] if var else None,
                defines=defs))
