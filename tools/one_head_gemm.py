"""One skinny head-backward GEMM launched a few times (for an ncu capture of that kernel alone):
   python tools/one_head_gemm.py [e|T] [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from miphei_vit_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "e"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
M = B * 256 * 256
bf = torch.bfloat16
f2 = torch.randn(M, 32, device="cuda").to(bf)
if which == "e":
    w = torch.randn(256, 64, device="cuda").to(bf)
    sc, sh = torch.rand(256, device="cuda") + 0.5, torch.randn(256, device="cuda")
    du = torch.randn(M, 16, device="cuda").to(bf)
    out = torch.empty(M, 256, device="cuda", dtype=bf)
    fn = lambda: ops.gemm(f2, w[:, :32], scale=sc, shift=sh, act=ops.ACT_GATE_MASK, in2=du, out=out)  # noqa: E731
else:
    w = torch.randn(144, 64, device="cuda").to(bf)
    out = torch.empty(M, 144, device="cuda", dtype=bf)
    fn = lambda: ops.gemm(f2, w[:, :32], out=out)  # noqa: E731
for _ in range(3):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    fn()
e1.record()
torch.cuda.synchronize()
print("%s: %.1f us" % (which, e0.elapsed_time(e1) / 5 * 1e3))
