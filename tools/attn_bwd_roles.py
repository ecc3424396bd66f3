"""Phase / role cycle shares of the two attention backward kernels from the diagnostic build (-DMV_GEMM_PROFILE=1):
   MIPHEI_B200_LIB=$PWD/miphei-vit_b200/libmiphei_b200_prof.so python tools/attn_bwd_roles.py [B] [N]"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from miphei_vit_b200 import lib, ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 329
H = 24
qkv = (torch.randn(B * N, 3 * H * 64, device="cuda") * 0.5).bfloat16()
do = torch.randn(B * N, H * 64, device="cuda").bfloat16()
out, lse = ops.attn_fwd(qkv, B, N, H, want_lse=True)
dq = ops.attn_bwd(qkv, out, do, lse, B, N, H)
dsum = torch.empty((B, H, N), dtype=torch.float32, device="cuda")
L = lib.init(0)
G = 296 if os.environ.get("MV_ATTN_BWD_V") == "1" else 148  # CTAs per kernel (version 1: two per SM)
buf = torch.zeros(16 * 2 * G, dtype=torch.int64, device="cuda")
for _ in range(200):  # also brings the SM clock up
    ops.attn_bwd(qkv, out, do, lse, B, N, H, dqkv=dq, dsum=dsum)
torch.cuda.synchronize()
L.mv_attn_set_profile_buffer(ctypes.c_void_p(buf.data_ptr()))
ops.attn_bwd(qkv, out, do, lse, B, N, H, dqkv=dq, dsum=dsum)
torch.cuda.synchronize()
L.mv_attn_set_profile_buffer(ctypes.c_void_p(0))
p = buf.view(-1, 16).double().cpu()
for k, name in enumerate(("dK/dV kernel", "dQ kernel")):
    q = p[k * G:(k + 1) * G]
    q = q[q[:, 5] > 0]
    life, tiles = q[:, 5].mean().item(), q[:, 6].mean().item()
    print("%s  B=%d N=%d: math warp lifetime %.0f cycles, %.1f tile pairs per CTA (%.0f cycles per tile pair)" % (name, B, N, life, tiles, life / tiles))
    for i, n in enumerate(["statistics staging", "wait X, Y", "ld + math", "st + arrive", "item epilogue"]):
        print("    %-20s %8.0f cycles  %5.1f %%" % (n, q[:, i].mean().item(), 100 * q[:, i].mean().item() / life))
    ml = q[:, 12].mean().item()
    print("  MMA thread lifetime %.0f cycles: waiting for operands %.1f %%, issuing X/Y %.1f %%, waiting for P/dS %.1f %%, issuing accumulations %.1f %%"
          % (ml, *[100 * q[:, 8 + i].mean().item() / ml for i in range(4)]))
