"""Which role bounds each encoder GEMM: per-CTA cycle counters from mv_gemm_set_profile_buffer (diagnostic build path).

Fractions are of the CTA lifetime: prod_wait = producer blocked on a free smem slot (MMA-bound when high),
mma_wait_ops = MMA warp blocked on operands (load-bound), mma_wait_acc = blocked on a free accumulator (epilogue-bound),
epi_wait / epi_busy = epilogue warp 2 waiting for / draining an accumulator."""
import ctypes
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the counters exist only in the diagnostic build: python miphei-vit_b200/build.py --prof
os.environ.setdefault("MIPHEI_B200_LIB", os.path.join(_ROOT, "miphei-vit_b200", "libmiphei_b200_prof.so"))

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from miphei_vit_b200 import lib as L, ops  # noqa: E402

D, H = 1536, 4096
nbuf = 5
buf = torch.zeros(16 * 2 * 148, dtype=torch.int64, device="cuda")


def prof(name, fn, flops, grid_hint=148):
    """Steady-state GPU time per launch from a CUDA graph of 20 launches (no CPU launch overhead, clocks stay up); the
    counters are those of the last launch of the graph."""
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    buf.zero_()
    L.load().mv_gemm_set_profile_buffer(ctypes.c_void_p(buf.data_ptr()))
    n = 20
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(n):
            fn(i)
    L.load().mv_gemm_set_profile_buffer(ctypes.c_void_p(0))
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (3 * n) * 1e3
    b = buf.view(-1, 16).double().cpu()
    b = b[b[:, 5] > 0]
    t_first = float(b[:, 8].min())
    tl = "entry spread %.1f us, prologue %.1f us, first CTA end %.1f us, last CTA end %.1f us (median %.1f)" % (
        (float(b[:, 8].max()) - t_first) / 1e3, float((b[:, 9] - b[:, 8]).mean()) / 1e3, (float(b[:, 10].min()) - t_first) / 1e3,
        (float(b[:, 10].max()) - t_first) / 1e3, (float(b[:, 10].median()) - t_first) / 1e3)
    life = b[:, 5]
    f = lambda c: float((b[:, c] / life).mean())  # noqa: E731
    print("%-34s %7.1f us %6.0f TF/s | ctas %3d life %6.0f clk | prod_wait %.2f mma_wait_ops %.2f mma_wait_acc %.2f | epi_wait %.2f epi_busy %.2f | tiles/cta %.1f" % (
        name, us, flops / us / 1e6, b.shape[0], float(life.mean()), f(0), f(1), f(2), f(3), f(4), float(b[:, 6].mean())), flush=True)
    print("      " + tl, flush=True)


for M in (5264, 10528):
    xs = [torch.randn(M, D, device="cuda").bfloat16() for _ in range(nbuf)]
    us_ = [torch.randn(M, H, device="cuda").bfloat16() for _ in range(3)]
    wqkv = [(torch.randn(3 * D, D, device="cuda") * 0.03).bfloat16() for _ in range(nbuf)]
    wproj = [(torch.randn(D, D, device="cuda") * 0.03).bfloat16() for _ in range(nbuf)]
    w1 = [(torch.randn(2 * H, D, device="cuda") * 0.03).bfloat16() for _ in range(nbuf)]
    w2 = [(torch.randn(D, H, device="cuda") * 0.03).bfloat16() for _ in range(nbuf)]
    bq, b1, g = torch.randn(3 * D, device="cuda"), torch.randn(2 * H, device="cuda"), torch.rand(D, device="cuda")
    xres = [torch.randn(M, D, device="cuda") for _ in range(3)]
    qkv = torch.empty(M, 3 * D, device="cuda", dtype=torch.bfloat16)
    u = torch.empty(M, H, device="cuda", dtype=torch.bfloat16)
    for bn in (0, 256, 192, 128):
        prof("M=%d QKV bn=%d" % (M, bn), lambda i: ops.gemm(xs[i % nbuf], wqkv[i % nbuf], shift=bq, out=qkv, block_n=bn), 2.0 * M * D * 3 * D)
    for bn in (0, 256, 192, 128):
        prof("M=%d proj+res bn=%d" % (M, bn), lambda i: ops.gemm(xs[i % nbuf], wproj[i % nbuf], scale=g, shift=g, resid=xres[i % 3], out=xres[i % 3], block_n=bn), 2.0 * M * D * D)
    prof("M=%d proj+res bn=256 nopair" % M, lambda i: ops.gemm(xs[i % nbuf], wproj[i % nbuf], scale=g, shift=g, resid=xres[i % 3], out=xres[i % 3], block_n=256, pair=1), 2.0 * M * D * D)
    prof("M=%d proj bf16-out" % M, lambda i: ops.gemm(xs[i % nbuf], wproj[i % nbuf], shift=g, out=xs[(i + 1) % nbuf]), 2.0 * M * D * D)
    for bn in (0, 256, 192):
        prof("M=%d fc2+res bn=%d" % (M, bn), lambda i: ops.gemm(us_[i % 3], w2[i % nbuf], scale=g, shift=g, resid=xres[i % 3], out=xres[i % 3], block_n=bn), 2.0 * M * D * H)
    prof("M=%d fc1 swiglu" % M, lambda i: ops.gemm(xs[i % nbuf], w1[i % nbuf], mode=ops.GEMM_SWIGLU, shift=b1, out=u), 2.0 * M * D * 2 * H)
