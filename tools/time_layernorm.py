"""Steady-state LayerNorm forward time (CUDA graph of 40 launches over rotating buffers): python tools/time_layernorm.py
   MV_LN_STREAM=0 selects the register-resident kernel for every M."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from miphei_vit_b200 import ops
D = 1536
w = torch.ones(D, device="cuda"); b = torch.zeros(D, device="cuda")
for M in (5264, 10528, 2632):
    xs = [torch.randn(M, D, device="cuda") for _ in range(4)]
    ys = [torch.empty(M, D, device="cuda", dtype=torch.bfloat16) for _ in range(4)]
    for i in range(4): ops.layernorm_fwd(xs[i], w, b, out=ys[i])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(40): ops.layernorm_fwd(xs[i % 4], w, b, out=ys[i % 4])
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 200 * 1e3
    print("layernorm_fwd M=%d: %.1f us  %.0f GB/s" % (M, us, M * D * 6 / us / 1e3), flush=True)
