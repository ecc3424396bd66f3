"""HBM-bound decoder kernels vs the measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs)."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from miphei_vit_b200 import ops
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
bf = torch.bfloat16
def t(name, fn, byts, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print("%-44s %8.1f us  %6.0f GB/s  %.2f of HBM peak" % (name, us, byts / us / 1e3, byts / us / 1e3 / peak), flush=True)
for (hw, C) in [(16, 1536), (32, 256), (64, 128), (128, 64)]:
    x = torch.randn(B, hw, hw, C, device="cuda").to(bf)
    out = torch.empty(B, 2 * hw, 2 * hw, C, device="cuda", dtype=bf)
    t("upsample2x %dx%d C=%d" % (hw, hw, C), lambda: ops.upsample2x(x, out=out), x.numel() * 2 * 5)
    d = torch.randn(B, 2 * hw, 2 * hw, C, device="cuda").to(bf)
    o2 = torch.empty(B, hw, hw, C, device="cuda", dtype=bf)
    t("upsample2x_bwd %dx%d C=%d" % (hw, hw, C), lambda: ops.upsample2x_bwd(d, out=o2), x.numel() * 2 * 5)
M = B * 256 * 256
f2 = torch.randn(M, 32, device="cuda").to(bf)
t("transpose_bf16 [M,32]", lambda: ops.transpose_bf16(f2), M * 32 * 2 * 2)
x = torch.randn(10528, 1536, device="cuda")
w = torch.ones(1536, device="cuda"); b_ = torch.zeros(1536, device="cuda")
y = torch.empty(10528, 1536, device="cuda", dtype=bf)
t("layernorm_fwd M=10528", lambda: ops.layernorm_fwd(x, w, b_, out=y), 10528 * 1536 * 6)
t("layernorm_fwd M=5264", lambda: ops.layernorm_fwd(x[:5264], w, b_, out=y[:5264]), 5264 * 1536 * 6)
dy = torch.randn(10528, 1536, device="cuda").to(bf); dres = torch.randn(10528, 1536, device="cuda")
dx = torch.empty_like(x); dxb = torch.empty_like(y)
t("layernorm_bwd M=10528", lambda: ops.layernorm_bwd(x, w, dy, dres=dres, out=dx, out_bf16=dxb), 10528 * 1536 * (4 + 2 + 4 + 4 + 2))
pred = torch.rand(B, 16, 256, 256, device="cuda"); tgt = torch.rand(B, 16, 256, 256, device="cuda")
t("loss_fwd_bwd", lambda: ops.loss_fwd_bwd(pred, tgt, torch.ones(16, device="cuda")), pred.numel() * 12)
