"""One eval forward without CUDA graphs, for `ncu` launch lists / captures (never a bench number)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from miphei_vit_b200.generators.mipheivit import get_vitmatte  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 40
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
with torch.device("cuda"):
    m = get_vitmatte("hoptimus0", 256, 16, use_lora=True, pretrained=False, depth=depth)
m = m.cuda().eval()
m.engine.use_graphs = False
x = torch.randn(B, 3, 256, 256, device="cuda")
for _ in range(iters):
    y = m.engine.infer(x, reuse_output=True)
torch.cuda.synchronize()
print("done", float(y.abs().mean()))
