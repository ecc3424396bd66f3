import os, sys, time, torch
sys.path.insert(0, "/root/repo")
import bench
dev = torch.device("cuda", 0)
model = bench.build_model(dev).eval()
eng = model.engine
B, S = 16, 256
hosts = [bench.synth_batch(torch, B, S, 5 + i, dev).pin_memory() for i in range(3)]
xd = hosts[0].to(dev)
eng.infer(xd, reuse_output=True); torch.cuda.synchronize()
def timeit(name, fn, n=30):
    fn(4); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(n); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("%-40s %.1f tiles/s  %.3f ms/step" % (name, B * n / dt, dt / n * 1e3), flush=True)
def resident(n):
    for _ in range(n): eng.infer(xd, reuse_output=True)
timeit("resident", resident)
oh = torch.empty((B, 16, S, S), dtype=torch.uint8).pin_memory()
def seq(n):
    for i in range(n):
        y = eng.infer(hosts[i % 3].to(dev, non_blocking=True), out_dtype=torch.uint8, reuse_output=True)
        oh.copy_(y, non_blocking=True)
timeit("sequential fp32", seq)
for depth in (2, 3):
    def st(n):
        for o in eng.infer_stream((hosts[i % 3] for i in range(n)), depth=depth): pass
    timeit("stream fp32 depth %d" % depth, st)
raw = [torch.randint(0, 255, (B, S, S, 3), dtype=torch.uint8).pin_memory() for _ in range(3)]
def st8(n):
    for o in eng.infer_stream((raw[i % 3] for i in range(n))): pass
timeit("stream u8", st8)
def h2d_only(n):
    for i in range(n): eng._workspace(B).x_in.copy_(hosts[i % 3], non_blocking=True)
timeit("h2d only (12.6 MB)", h2d_only)
def d2h_only(n):
    for i in range(n): oh.copy_(eng._workspace(B).out_u8, non_blocking=True)
timeit("d2h only (16.8 MB)", d2h_only)
