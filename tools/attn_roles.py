"""Phase / role cycle shares of the attention forward kernel from the diagnostic build (-DMV_GEMM_PROFILE=1):
   MIPHEI_B200_LIB=$PWD/miphei-vit_b200/libmiphei_b200_prof.so python tools/attn_roles.py [B] [N]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from miphei_vit_b200 import lib, ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 329
H = 24
qkv = (torch.randn(B * N, 3 * H * 64, device="cuda") * 0.5).bfloat16()
out = torch.empty(B * N, H * 64, dtype=torch.bfloat16, device="cuda")
L = lib.init(0)
buf = torch.zeros(16 * 2 * 148, dtype=torch.int64, device="cuda")
for _ in range(300):  # also brings the SM clock up
    ops.attn_fwd(qkv, B, N, H, out=out)
torch.cuda.synchronize()
L.mv_attn_set_profile_buffer(ctypes.c_void_p(buf.data_ptr()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.attn_fwd(qkv, B, N, H, out=out)
e1.record()
torch.cuda.synchronize()
L.mv_attn_set_profile_buffer(ctypes.c_void_p(0))
p = buf.view(-1, 16)[:148].double().cpu()
life = p[:, 7].mean().item()
names = ["wait S", "ld S + max", "exchange barrier", "wait PV + rescale", "exp pass", "st wait + arrive", "item epilogue"]
print("B=%d N=%d: %.1f us; softmax warp (slot 0) lifetime %.0f cycles, %.1f items per CTA" % (B, N, e0.elapsed_time(e1) * 1e3, life, p[:, 11].mean().item()))
for i, n in enumerate(names):
    print("  %-18s %8.0f cycles  %5.1f %%" % (n, p[:, i].mean().item(), 100 * p[:, i].mean().item() / life))
ml = p[:, 10].mean().item()
print("MMA thread lifetime %.0f cycles: waiting for Q / K / V %.1f %%, for P %.1f %%, for the S columns %.1f %%"
      % (ml, 100 * p[:, 8].mean().item() / ml, 100 * p[:, 9].mean().item() / ml, 100 * p[:, 12].mean().item() / ml))
print("            issuing S MMAs + commits %.1f %%, issuing PV MMAs + commits %.1f %%" % (100 * p[:, 13].mean().item() / ml, 100 * p[:, 14].mean().item() / ml))
