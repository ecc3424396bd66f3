"""Graph-replayed eval forward at batch B (used by tools/ab_lib.sh for same-box A/B of two library builds)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda", 0)
model = bench.build_model(dev).eval()
eng = model.engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
x = bench.synth_batch(torch, B, 256, 5, dev).to(dev)
def run(tag, n=30):
    for ws in eng._ws.values():
        ws.graph = None
    for _ in range(4): eng.infer(x, reuse_output=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): eng.infer(x, reuse_output=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("%-28s %.3f ms  %.1f tiles/s" % (tag, ms, B / ms * 1e3), flush=True)
run("eval forward B=%d" % B)
