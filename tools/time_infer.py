"""Quick timing of the eval forward at full size (CUDA events; graph replay)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from miphei_vit_b200.generators.mipheivit import get_vitmatte  # noqa: E402
from miphei_vit_b200 import lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
t0 = time.time()
torch.manual_seed(0)
with torch.device("cuda"):
    m = get_vitmatte("hoptimus0", 256, 16, use_lora=True, pretrained=False)
m = m.cuda().eval()
with torch.no_grad():
    for blk in m.encoder.vit.blocks:
        blk.ls1.gamma.uniform_(0.05, 0.5)
        blk.ls2.gamma.uniform_(0.05, 0.5)
print("build %.1fs" % (time.time() - t0), flush=True)
x = torch.randn(B, 3, 256, 256, device="cuda")
for use_graphs, split in ((False, False), (True, False), (True, True)):
    m.engine.use_graphs = use_graphs
    m.engine.split_streams = split
    for w in m.engine._ws.values():
        w.graph = None
    for _ in range(3):
        y = m.engine.infer(x, reuse_output=True)
    torch.cuda.synchronize()
    lib.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 10
    for _ in range(n):
        y = m.engine.infer(x, reuse_output=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("split=%s graphs=%s B=%d: %.3f ms/forward, %.1f tiles/s, %.1f TFLOP/s (793.4 GF/tile), launches/fwd %d" % (
        split, use_graphs, B, ms, B / ms * 1e3, B * 793.4 / ms, lib.launch_count() // n), flush=True)
print("finite", torch.isfinite(y).all().item(), float(y.abs().mean()))
