"""Is the (eagerly launched) training step CPU-bound? Host time to ISSUE one step vs GPU time per step."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from miphei_vit_b200.trainer import Trainer
from miphei_vit_b200 import lib
dev = torch.device("cuda", 0)
model = bench.build_model(dev)
B = 32
x = torch.randn(B, 3, 256, 256, device=dev)
y = torch.rand(B, 16, 256, 256, device=dev) * 1.8 - 0.9
tr = Trainer(model, marker_weights=torch.linspace(1, 10, 16), batch_size=B, total_steps=1000)
for _ in range(3):
    tr.step(x, y)
torch.cuda.synchronize()
lib.reset_launch_count()
issue = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    torch.cuda.synchronize()   # empty queue: the issue time below is pure host work
    t0 = time.perf_counter()
    tr.step(x, y)
    issue.append(time.perf_counter() - t0)
e1.record()
torch.cuda.synchronize()
n_launch = lib.launch_count() / 5
e0.record()
for _ in range(5):
    tr.step(x, y)
e1.record(); torch.cuda.synchronize()
print("host issue time per step %.1f ms (min %.1f), library launches per step %d; GPU step (back to back) %.2f ms" % (
    1e3 * sum(issue) / len(issue), 1e3 * min(issue), n_launch, e0.elapsed_time(e1) / 5))
