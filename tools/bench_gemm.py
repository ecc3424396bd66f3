"""Micro-benchmark of mv_gemm_bf16 on the encoder shapes (CUDA events, L2-sized rotation of operands)."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from miphei_vit_b200 import ops  # noqa: E402


def bench(M, N, K, mode=0, bn=0, iters=20, compare=True, pair=0):
    nbuf = 4
    As = [torch.randn(M, K, device="cuda").bfloat16() for _ in range(nbuf)]
    Bs = [(torch.randn(N, K, device="cuda") * 0.05).bfloat16() for _ in range(nbuf)]
    bias = torch.randn(N, device="cuda")
    out = None
    kw = dict(mode=mode, block_n=bn, pair=pair)
    if mode == ops.GEMM_SWIGLU:
        kw["shift"] = bias
    for i in range(3):
        out = ops.gemm(As[i % nbuf], Bs[i % nbuf], **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        ops.gemm(As[i % nbuf], Bs[i % nbuf], out=out, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    res = {"M": M, "N": N, "K": K, "mode": mode, "bn": bn, "pair": pair, "ms": round(ms, 4), "tflops": round(2.0 * M * N * K / ms / 1e9, 1)}
    if compare and mode == 0:
        for i in range(3):
            torch.matmul(As[i % nbuf], Bs[i % nbuf].t())
        torch.cuda.synchronize()
        e0.record()
        for i in range(iters):
            torch.matmul(As[i % nbuf], Bs[i % nbuf].t())
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / iters
        res["cublas_tflops"] = round(2.0 * M * N * K / ms2 / 1e9, 1)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    for M in (5264, 10528):
        for bn in (256, 192, 0):
            bench(M, 4608, 1552, bn=bn, compare=(bn == 256))
            bench(M, 1536, 1536, bn=bn, compare=(bn == 256))
            bench(M, 1536, 4096, bn=bn, compare=(bn == 256))
