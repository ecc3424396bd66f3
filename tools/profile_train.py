"""One training step (after one warm-up step) for ncu launch lists."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from miphei_vit_b200.generators.mipheivit import get_vitmatte
from miphei_vit_b200.trainer import Trainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 40
with torch.device("cuda"):
    m = get_vitmatte("hoptimus0", 256, 16, use_lora=True, pretrained=False, depth=depth)
m = m.cuda()
x = torch.randn(B, 3, 256, 256, device="cuda")
y = torch.rand(B, 16, 256, 256, device="cuda") * 1.8 - 0.9
tr = Trainer(m, marker_weights=torch.linspace(1, 10, 16), batch_size=B, use_graph=False)  # eager launches for ncu
for _ in range(2):
    l = tr.step(x, y)
torch.cuda.synchronize()
print("done", l.item())
