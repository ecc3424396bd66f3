"""Decoder-shaped GEMM launches in isolation (wgrad, skinny K=32 epilogue-bound shapes) for timing / ncu captures.

  python tools/bench_decoder_gemms.py [B] [iters]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from miphei_vit_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = "cuda"
bf = torch.bfloat16


def timeit(name, fn, flops=None, bytes_=None):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    extra = ""
    if flops:
        extra += " %.0f TFLOP/s" % (flops / us / 1e6)
    if bytes_:
        extra += " %.0f GB/s" % (bytes_ / us / 1e3)
    print("%-40s %9.1f us%s" % (name, us, extra), flush=True)


def wgrad(name, cout, c0, c1, hw, stride=1):
    M = B * hw * hw
    ld = (M + 7) // 8 * 8
    dzT = torch.randn(cout, ld, device=dev).to(bf)
    s0 = torch.randn(B, hw * stride, hw * stride, c0, device=dev).to(bf)
    s1 = torch.randn(B, hw * stride, hw * stride, c1, device=dev).to(bf) if c1 else None
    kp = 9 * (((c0 + 63) // 64) + ((c1 + 63) // 64)) * 64
    dwp = torch.zeros(cout, kp, device=dev)
    timeit(name, lambda: ops.gemm(dzT[:, :M], s0, mode=ops.GEMM_NN_ATOMIC, conv=dict(stride=stride, a2=s1), out=dwp),
           flops=2.0 * cout * 9 * (c0 + c1) * M)


wgrad("wgrad fu3 (32 <- 8|64 @256)", 32, 8, 64, 256)
wgrad("wgrad fu2 (64 <- 48|128 @128)", 64, 48, 128, 128)
wgrad("wgrad fu1 (128 <- 96|256 @64)", 128, 96, 256, 64)
wgrad("wgrad fu0 (256 <- 192|1536 @32)", 256, 192, 1536, 32)

M = B * 256 * 256
f2 = torch.randn(M, 32, device=dev).to(bf)
w256 = torch.randn(256, 64, device=dev).to(bf)
sc, sh = torch.rand(256, device=dev) + 0.5, torch.randn(256, device=dev)
w2, b2 = torch.randn(256, device=dev), torch.randn(16, device=dev)
du = torch.randn(M, 16, device=dev).to(bf)
e = torch.empty(M, 256, device=dev, dtype=bf)
timeit("e = mask(f W1^T) [M,256] K=32", lambda: ops.gemm(f2, w256[:, :32], scale=sc, shift=sh, act=ops.ACT_GATE_MASK, in2=du, out=e),
       bytes_=M * (64 + 512 + 32))
T = torch.empty(M, 144, device=dev, dtype=bf)
w144 = torch.randn(144, 64, device=dev).to(bf)
timeit("T = f W3t^T [M,144] K=32", lambda: ops.gemm(f2, w144[:, :32], out=T), bytes_=M * (64 + 288))
stats = torch.zeros(2, 256, device=dev)
timeit("head stats [M,256] K=32 no_out", lambda: ops.gemm(f2, w256[:, :32], shift=sh, colstats=stats, no_out=True), bytes_=M * 64)
gate = torch.zeros(M, 16, device=dev, dtype=bf)
timeit("HEAD_GATE", lambda: ops.gemm(f2, w256[:, :32], mode=ops.GEMM_HEAD_GATE, scale=sc, shift=sh, in2=w2, resid=b2, out=gate),
       bytes_=M * (64 + 32))
f4 = f2.view(B, 256, 256, 32)
wc = torch.randn(16, 576, device=dev).to(bf)
b3 = torch.randn(16, device=dev)
pred = torch.empty(B, 16, 256, 256, device=dev)
timeit("HEAD_CONV", lambda: ops.gemm(f4, wc, mode=ops.GEMM_HEAD_CONV, conv=dict(stride=1), shift=b3, in2=gate, out=pred),
       bytes_=M * (64 + 32 + 64))
img8 = torch.randn(B, 256, 256, 8, device=dev).to(bf)
up = torch.randn(B, 256, 256, 64, device=dev).to(bf)
wb3 = torch.randn(32, 9 * 128, device=dev).to(bf)
z = torch.empty(M, 32, device=dev)
st = torch.zeros(2, 32, device=dev)
timeit("fu3 conv fwd (32 <- 8|64 @256) + stats", lambda: ops.gemm(img8, wb3, conv=dict(stride=1, a2=up), colstats=st, out=z),
       flops=2.0 * M * 32 * 9 * 67, bytes_=M * (16 + 128 + 128))
dz = torch.randn(B, 256, 256, 32, device=dev).to(bf)
wd3 = torch.randn(64, 9 * 64, device=dev).to(bf)
dx = torch.empty(M, 64, device=dev, dtype=bf)
timeit("fu3 dgrad (64 <- 32 @256)", lambda: ops.gemm(dz, wd3, conv=dict(stride=1), out=dx), flops=2.0 * M * 64 * 9 * 32,
       bytes_=M * (64 + 128))
