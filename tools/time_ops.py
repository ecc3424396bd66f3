"""In-pipeline (warm L2, real operand residency) GPU time per op: wraps every miphei_vit_b200.ops entry point with CUDA
events, runs the eval forward (no graphs) or training steps, and prints the per-op totals of one step.

  python tools/time_ops.py infer 16        python tools/time_ops.py train 32 [depth]
"""
import os
import sys
from collections import OrderedDict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from miphei_vit_b200 import ops  # noqa: E402
from miphei_vit_b200.generators.mipheivit import get_vitmatte  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "infer"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 40
records = []
enabled = [False]


def wrap(name, fn):
    def inner(*a, **k):
        if not enabled[0]:
            return fn(*a, **k)
        key = name
        if name == "gemm":
            A, Bm = a[0], a[1]
            mode = k.get("mode", 0)
            conv = k.get("conv")
            if conv is not None:
                key = "gemm mode%d conv A%s B%s" % (mode, tuple(A.shape), tuple(Bm.shape))
            else:
                key = "gemm mode%d A%s B%s%s" % (mode, tuple(A.shape), tuple(Bm.shape), " act%d" % k["act"] if k.get("act") else "")
        elif a and hasattr(a[0], "shape"):
            key = "%s %s" % (name, tuple(a[0].shape))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        records.append((key, e0, e1))
        return r
    return inner


for n in dir(ops):
    f = getattr(ops, n)
    if callable(f) and not n.startswith("_") and getattr(f, "__module__", "") == ops.__name__:
        setattr(ops, n, wrap(n, f))

with torch.device("cuda"):
    m = get_vitmatte("hoptimus0", 256, 16, use_lora=True, pretrained=False, depth=depth)
m = m.cuda()
x = torch.randn(B, 3, 256, 256, device="cuda")
if what == "infer":
    m.eval()
    m.engine.use_graphs = False
    run = lambda: m.engine.infer(x, reuse_output=True)  # noqa: E731
else:
    from miphei_vit_b200.trainer import Trainer
    y = torch.rand(B, 16, 256, 256, device="cuda") * 1.8 - 0.9
    tr = Trainer(m, marker_weights=torch.linspace(1, 10, 16), batch_size=B, use_graph=False)  # per-op events need eager launches
    run = lambda: tr.step(x, y)  # noqa: E731
for _ in range(3):
    run()
torch.cuda.synchronize()
iters = 3
enabled[0] = True
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(iters):
    run()
t1.record()
torch.cuda.synchronize()
agg = OrderedDict()
for key, e0, e1 in records:
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1) * 1e3
tot = sum(v[1] for v in agg.values()) / iters
print("# %s B=%d depth=%d: wall %.1f us/step, sum of op times %.1f us/step" % (what, B, depth, t0.elapsed_time(t1) * 1e3 / iters, tot))
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-95s n=%4d  %9.1f us/step  %7.1f us/call  %5.1f%%" % (k, c // iters, t / iters, t / c, 100 * t / iters / tot))
