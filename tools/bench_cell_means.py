"""Per-nucleus mean extraction: kernel time vs the HBM roofline and vs the reference's per-image torch.unique / scatter_add_
loop on the same GPU (restated here with torch ops — the reference function itself is not importable on the GPU box)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from miphei_vit_b200 import ops


def synthetic_nuclei(batch, size, n_cells, seed=0):
    """Label maps with `n_cells` random discs per image (later discs overwrite earlier ones)."""
    g = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:size, 0:size]
    out = np.zeros((batch, size, size), dtype=np.int64)
    for b in range(batch):
        for k in range(n_cells):
            cy, cx, r = g.integers(0, size), g.integers(0, size), g.integers(2, 9)
            out[b][(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = 1 + 7 * k + 1000003 * b
    return torch.from_numpy(out)


B, C, S = 32, 16, 256
pred = torch.rand(B, C, S, S, device="cuda")
target = torch.rand(B, C, S, S, device="cuda")
nuclei = synthetic_nuclei(B, S, 250, seed=1).cuda()
def torch_loop():
    outs = []
    for b in range(B):
        nb = nuclei[b]
        m = nb > 0
        flat = nb[m]
        if flat.numel() == 0:
            continue
        u, inv = torch.unique(flat, return_inverse=True)
        pf = pred[b].permute(1, 2, 0)[m]
        tf = target[b].permute(1, 2, 0)[m]
        ps = torch.zeros((u.shape[0], C), device="cuda").scatter_add_(0, inv.unsqueeze(1).expand(-1, C), pf)
        ts = torch.zeros((u.shape[0], C), device="cuda").scatter_add_(0, inv.unsqueeze(1).expand(-1, C), tf)
        n = torch.zeros(u.shape[0], device="cuda").scatter_add_(0, inv, torch.ones_like(flat, dtype=torch.float32))
        outs.append((ps / n[:, None], ts / n[:, None], u))
    return outs
def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
t_k = timeit(lambda: ops.cell_means(pred, target, nuclei))
t_t = timeit(torch_loop, 3)
byts = B * S * S * (2 * C * 4 + 8)
print("cell means B=%d C=%d %dx%d: kernel path %.1f us (%.0f GB/s algorithmic, incl. the row-count sync and pack), "
      "reference-style torch loop %.1f us -> %.1fx" % (B, C, S, S, t_k, byts / t_k / 1e3, t_t, t_t / t_k))
