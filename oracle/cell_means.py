"""CPU restatement of MeanCellExtrator.extract_mean (reference src/utils.py:49-121) — TEST INFRASTRUCTURE ONLY.

Per image: labels > 0 -> sorted unique ids (torch.unique semantics), per-id sums of the C prediction / target channels and
pixel counts, means = sums / counts; rows of image 0 first.  float64 accumulation (the reference's scatter_add_ runs in
the input dtype with an unspecified order, so it is matched to fp32 tolerance, ids / counts exactly).
Pinned against the reference class itself in tests/test_cell_means_cpu.py (imported with pytorch_lightning / hydra /
wandb stubbed) and against tests/golden/cell_means.pt."""
import numpy as np
import torch


def extract_mean(pred, target, nuclei):
    B, C, H, W = pred.shape
    nuclei = nuclei.reshape(B, H * W).cpu().numpy()
    p = pred.reshape(B, C, H * W).double().cpu().numpy()
    t = target.reshape(B, C, H * W).double().cpu().numpy()
    pm, tm, ids, cnts = [], [], [], []
    for b in range(B):
        lab = nuclei[b]
        keep = lab > 0
        if not keep.any():
            continue
        u, inv = np.unique(lab[keep], return_inverse=True)
        n = np.bincount(inv, minlength=len(u)).astype(np.float64)
        sp = np.stack([np.bincount(inv, weights=p[b, c][keep], minlength=len(u)) for c in range(C)], 1)
        st = np.stack([np.bincount(inv, weights=t[b, c][keep], minlength=len(u)) for c in range(C)], 1)
        pm.append(sp / n[:, None])
        tm.append(st / n[:, None])
        ids.append(u.astype(np.int64))
        cnts.append(n)
    if not ids:
        z = torch.zeros((0, C), dtype=pred.dtype)
        return z, z.clone(), torch.zeros(0, dtype=torch.int64), torch.zeros(0)
    return (torch.from_numpy(np.concatenate(pm)).to(pred.dtype), torch.from_numpy(np.concatenate(tm)).to(pred.dtype),
            torch.from_numpy(np.concatenate(ids)), torch.from_numpy(np.concatenate(cnts)).float())


def synthetic_nuclei(batch, size, n_cells, seed=0, id_offset=1, empty=()):
    """Label maps with `n_cells` random discs per image (later discs overwrite earlier ones), ids = id_offset + k * 7."""
    g = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:size, 0:size]
    out = np.zeros((batch, size, size), dtype=np.int64)
    for b in range(batch):
        if b in empty:
            continue
        for k in range(n_cells):
            cy, cx, r = g.integers(0, size), g.integers(0, size), g.integers(2, 9)
            out[b][(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = id_offset + 7 * k + 1000003 * b
    return torch.from_numpy(out)
