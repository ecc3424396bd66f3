"""Sweep of the SwiGLU forward / backward GEMM configurations (tile width, CTA pairs) on the encoder shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from miphei_vit_b200 import ops  # noqa: E402

D, H = 1536, 4096
nbuf = 6  # rotate operands so that weights / activations are not L2-resident across iterations


def run(name, fn, flops, iters=24):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print("%-44s %8.1f us  %7.1f TFLOP/s" % (name, us, flops / us / 1e6), flush=True)


for M in (5264, 10528):
    xs = [torch.randn(M, D, device="cuda").bfloat16() for _ in range(nbuf)]
    w1 = [(torch.randn(2 * H, D, device="cuda") * 0.03).bfloat16() for _ in range(nbuf)]
    b1 = torch.randn(2 * H, device="cuda") * 0.1
    u = torch.empty(M, H, device="cuda", dtype=torch.bfloat16)
    h = torch.empty(M, 2 * H, device="cuda", dtype=torch.bfloat16)
    for pair in (1, 2):
        run("M=%d fc1+SwiGLU fwd pair=%d" % (M, pair),
            lambda i: ops.gemm(xs[i % nbuf], w1[i % nbuf], mode=ops.GEMM_SWIGLU, shift=b1, out=u, pair=pair), 2.0 * M * D * 2 * H)
        run("M=%d fc1+SwiGLU fwd +save h pair=%d" % (M, pair),
            lambda i: ops.gemm(xs[i % nbuf], w1[i % nbuf], mode=ops.GEMM_SWIGLU, shift=b1, out=u, aux=h, pair=pair), 2.0 * M * D * 2 * H)
    w2t = [(torch.randn(H, D, device="cuda") * 0.03).bfloat16() for _ in range(nbuf)]  # dU = dY[M, D] . W2[D, H] -> B = W2^T [H, D]
    hs = [torch.randn(M, 2 * H, device="cuda").bfloat16() for _ in range(2)]
    dh = torch.empty(M, 2 * H, device="cuda", dtype=torch.bfloat16)
    for bn in (128, 256):
        for pair in (0, 2):
            run("M=%d SwiGLU bwd bn=%d pair=%d" % (M, bn, pair),
                lambda i: ops.gemm(xs[i % nbuf], w2t[i % nbuf], mode=ops.GEMM_SWIGLU_BWD, in2=hs[i % 2], out=dh, block_n=bn, pair=pair),
                2.0 * M * D * H)
    # correctness of the new configurations against the default one
    ref = ops.gemm(xs[0], w2t[0], mode=ops.GEMM_SWIGLU_BWD, in2=hs[0]).float()
    for bn, pair in ((256, 0), (256, 2), (128, 2)):
        got = ops.gemm(xs[0], w2t[0], mode=ops.GEMM_SWIGLU_BWD, in2=hs[0], block_n=bn, pair=pair).float()
        print("  bwd bn=%d pair=%d max abs diff vs default %.3e" % (bn, pair, (got - ref).abs().max().item()))
    r1 = ops.gemm(xs[0], w1[0], mode=ops.GEMM_SWIGLU, shift=b1, pair=1).float()
    g1 = ops.gemm(xs[0], w1[0], mode=ops.GEMM_SWIGLU, shift=b1, pair=2).float()
    print("  fwd pair=2 max abs diff vs default %.3e" % (g1 - r1).abs().max().item())
