"""Role-stall counters (diagnostic build) for the decoder's implicit-GEMM convolutions: weight gradients, forward, dgrad."""
import ctypes, os, sys
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("MIPHEI_B200_LIB", os.path.join(_ROOT, "miphei-vit_b200", "libmiphei_b200_prof.so"))
import torch
sys.path.insert(0, _ROOT)
from miphei_vit_b200 import lib as L, ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
bf = torch.bfloat16
buf = torch.zeros(16 * 2 * 148, dtype=torch.int64, device="cuda")


def prof(name, fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    buf.zero_()
    L.load().mv_gemm_set_profile_buffer(ctypes.c_void_p(buf.data_ptr()))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(4):
            fn()
    L.load().mv_gemm_set_profile_buffer(ctypes.c_void_p(0))
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 4 * 1e3
    b = buf.view(-1, 16).double().cpu(); b = b[b[:, 5] > 0]
    life = b[:, 5]
    f = lambda c: float((b[:, c] / life).mean())
    print("%-36s %8.1f us | ctas %3d life %8.0f clk | prod_wait %.2f mma_wait_ops %.2f mma_wait_acc %.2f | epi_wait %.2f epi_busy %.2f | tiles/cta %.1f" % (
        name, us, b.shape[0], float(life.mean()), f(0), f(1), f(2), f(3), f(4), float(b[:, 6].mean())), flush=True)


def wgrad(name, cout, c0, c1, hw):
    M = B * hw * hw
    dzT = torch.randn(cout, M, device="cuda").to(bf)
    s0 = torch.randn(B, hw, hw, c0, device="cuda").to(bf)
    s1 = torch.randn(B, hw, hw, c1, device="cuda").to(bf) if c1 else None
    kp = 9 * (((c0 + 63) // 64) + ((c1 + 63) // 64)) * 64
    dwp = torch.zeros(cout, kp, device="cuda")
    prof(name, lambda: ops.gemm(dzT, s0, mode=ops.GEMM_NN_ATOMIC, conv=dict(stride=1, a2=s1), out=dwp))


wgrad("wgrad fu3 (32 <- 8|64 @256)", 32, 8, 64, 256)
wgrad("wgrad fu2 (64 <- 48|128 @128)", 64, 48, 128, 128)
wgrad("wgrad fu0 (256 <- 192|1536 @32)", 256, 192, 1536, 32)
M = B * 256 * 256
img8 = torch.randn(B, 256, 256, 8, device="cuda").to(bf)
up = torch.randn(B, 256, 256, 64, device="cuda").to(bf)
wb3 = torch.randn(32, 9 * 128, device="cuda").to(bf)
z = torch.empty(M, 32, device="cuda")
st = torch.zeros(2, 32, device="cuda")
prof("fu3 conv fwd (32 <- 8|64 @256)", lambda: ops.gemm(img8, wb3, conv=dict(stride=1, a2=up), colstats=st, out=z))
dz = torch.randn(B, 256, 256, 32, device="cuda").to(bf)
wd3 = torch.randn(64, 9 * 64, device="cuda").to(bf)
dx = torch.empty(M, 64, device="cuda", dtype=bf)
prof("fu3 dgrad (64 <- 32 @256)", lambda: ops.gemm(dz, wd3, conv=dict(stride=1), out=dx))
f2 = torch.randn(M, 32, device="cuda").to(bf)
w256 = torch.randn(256, 64, device="cuda").to(bf)
sc, sh = torch.rand(256, device="cuda") + 0.5, torch.randn(256, device="cuda")
du = torch.randn(M, 16, device="cuda").to(bf)
e = torch.empty(M, 256, device="cuda", dtype=bf)
prof("e = mask(f W1^T) [M,256] K=32", lambda: ops.gemm(f2, w256[:, :32], scale=sc, shift=sh, act=ops.ACT_GATE_MASK, in2=du, out=e))
w2, b2 = torch.randn(256, device="cuda"), torch.randn(16, device="cuda")
gate = torch.zeros(M, 16, device="cuda", dtype=bf)
prof("HEAD_GATE", lambda: ops.gemm(f2, w256[:, :32], mode=ops.GEMM_HEAD_GATE, scale=sc, shift=sh, in2=w2, resid=b2, out=gate))
wc = torch.randn(16, 576, device="cuda").to(bf)
b3 = torch.randn(16, device="cuda")
pred = torch.empty(B, 16, 256, 256, device="cuda")
prof("HEAD_CONV", lambda: ops.gemm(f2.view(B, 256, 256, 32), wc, mode=ops.GEMM_HEAD_CONV, conv=dict(stride=1), shift=b3, in2=gate, out=pred))
