"""Steady-state time of attention forward / backward (CUDA graph of 10 launches, rotating operands).
   `--fwd` times the forward kernel only (timing experiments with variant libraries whose results are wrong)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from miphei_vit_b200 import ops
for (B, N, H) in [(16, 329, 24), (32, 329, 24), (8, 1301, 24)]:
    qkvs = [(torch.randn(B * N, 3 * H * 64, device="cuda") * 2).bfloat16() for _ in range(3)]
    dos = [torch.randn(B * N, H * 64, device="cuda").bfloat16() for _ in range(3)]
    out, lse = ops.attn_fwd(qkvs[0], B, N, H, want_lse=True)
    dq = ops.attn_bwd(qkvs[0], out, dos[0], lse, B, N, H)
    dsum = torch.empty((B, H, N), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    def graph_time(fn, n=10):
        fn(0); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(n): fn(i)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): g.replay()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (3 * n) * 1e3
    tf = graph_time(lambda i: ops.attn_fwd(qkvs[i % 3], B, N, H, out=out, lse=lse))
    if "--fwd" in sys.argv:
        print("B=%d N=%d H=%d: fwd %.1f us (%.0f TF/s)" % (B, N, H, tf, 4.0 * N * N * 64 * H * B / tf / 1e6), flush=True)
        continue
    tb = graph_time(lambda i: ops.attn_bwd(qkvs[i % 3], out, dos[i % 3], lse, B, N, H, dqkv=dq, dsum=dsum))
    fl = 4.0 * N * N * 64 * H * B
    print("B=%d N=%d H=%d: fwd %.1f us (%.0f TF/s)  bwd %.1f us (%.0f TF/s at 2.5x fwd flops)" % (B, N, H, tf, fl / tf / 1e6, tb, 2.5 * fl / tb / 1e6), flush=True)
