"""BASELINE configs[3] on one GPU: HEMIT-style 3-channel head on 512-px tiles (36x36 + 5 = 1301 tokens, attention-heavier),
training step (fwd + bwd + weighted MSE + clip + Adam), batch 8, full-size ViT-g/14; plus the eval forward at batch 8."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from miphei_vit_b200.trainer import Trainer
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
model = bench.build_model(dev, out_chans=3, img=512)
x = torch.randn(B, 3, 512, 512, device=dev)
y = torch.rand(B, 3, 512, 512, device=dev) * 1.8 - 0.9
model.eval()
for _ in range(3):
    model.engine.infer(x, reuse_output=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    model.engine.infer(x, reuse_output=True)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("512 px, 3 ch, batch %d: eval forward %.2f ms = %.1f tiles/s = %.0f TFLOP/s (3443.6 GF/tile)" % (B, ms, B / ms * 1e3, B * 3443.6 / ms))
tr = Trainer(model, marker_weights=torch.ones(3), batch_size=B, total_steps=1000)
for _ in range(3):
    tr.step(x, y)
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    loss = tr.step(x, y)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("512 px, 3 ch, batch %d: training step %.2f ms = %.1f tiles/s = %.0f TFLOP/s (7379.4 GF/tile), loss %.4f" % (
    B, ms, B / ms * 1e3, B * 7379.4 / ms, float(loss)))
