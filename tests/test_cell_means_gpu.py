"""GPU parity of the per-nucleus mean kernel (mv_cell_means) against the oracle and the reference-generated goldens."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cell_means as oc  # noqa: E402
from test_cell_means_cpu import GOLDEN, _case  # noqa: E402

pytestmark = pytest.mark.gpu


def _check(pred, target, nuclei, **kw):
    from miphei_vit_b200 import ops
    rp, rt, rid, rcnt = oc.extract_mean(pred, target, nuclei)
    gp, gt, gid, nu, gcnt = ops.cell_means(pred.cuda(), target.cuda(), nuclei.cuda(), return_counts=True, **kw)
    assert torch.equal(gid.cpu(), rid)                      # ids: bit-exact, ascending per image, batch order
    assert torch.equal(gcnt.cpu(), rcnt)                    # pixel counts: exact
    per_img = [int((nuclei[b] > 0).any()) and len(torch.unique(nuclei[b][nuclei[b] > 0])) for b in range(nuclei.shape[0])]
    assert nu.cpu().tolist() == per_img
    assert torch.allclose(gp.cpu(), rp, rtol=1e-5, atol=1e-6)  # fp32 sums in a different order than the oracle's fp64
    assert torch.allclose(gt.cpu(), rt, rtol=1e-5, atol=1e-6)
    return gid


@pytest.mark.parametrize("seed,B,C,S,n_cells,empty", [(1, 3, 5, 64, 12, ()), (2, 2, 16, 256, 300, ()), (3, 4, 16, 128, 40, (0, 2)),
                                                       (4, 2, 3, 64, 0, ()), (5, 1, 1, 32, 5, ())])
def test_cell_means_matches_oracle(seed, B, C, S, n_cells, empty):
    pred, target, nuclei = _case(seed, B=B, C=C, S=S, n_cells=n_cells, empty=empty)
    _check(pred, target, nuclei)


def test_cell_means_matches_reference_golden():
    from miphei_vit_b200 import ops
    g = torch.load(GOLDEN, map_location="cpu", weights_only=False)
    for case in g["cases"]:
        pred, target, nuclei = _case(case["seed"], B=case["B"], C=case["C"], S=case["S"], n_cells=case["n_cells"],
                                     empty=tuple(case["empty"]))
        gp, gt, gid, nu = ops.cell_means(pred.cuda(), target.cuda(), nuclei.cuda())
        assert torch.equal(gid.cpu(), case["ids"])
        assert torch.allclose(gp.cpu(), case["pred_means"], rtol=1e-5, atol=1e-6)
        assert torch.allclose(gt.cpu(), case["target_means"], rtol=1e-5, atol=1e-6)


def test_cell_means_int32_labels_large_ids_and_table_growth():
    pred, target, nuclei = _case(7, B=2, C=4, S=128, n_cells=200)
    nuclei = nuclei + (nuclei > 0) * 2_000_000_000          # ids beyond int32 stay exact in int64
    _check(pred, target, nuclei, cap=64)                    # 64 rows are too few: the wrapper grows the table and re-runs
    small = (nuclei % 1000).int()                           # int32 label maps are accepted as they are
    _check(pred, target, small.long())
    from miphei_vit_b200 import ops
    a = ops.cell_means(pred.cuda(), target.cuda(), small.cuda())
    b = ops.cell_means(pred.cuda(), target.cuda(), small.long().cuda())
    assert torch.equal(a[2], b[2]) and torch.allclose(a[0], b[0])


def test_mean_cell_extrator_mirror_of_reference_class():
    from miphei_vit_b200.cells import MeanCellExtrator
    pred, target, nuclei = _case(9, B=2, C=6, S=64, n_cells=15)
    m = MeanCellExtrator(scale_factor=1.0)
    pm, tm, ids = m(pred.cuda(), target.cuda(), nuclei.cuda())
    rp, rt, rid, _ = oc.extract_mean(pred, target, nuclei)
    assert torch.equal(ids.cpu(), rid) and torch.allclose(pm.cpu(), rp, rtol=1e-5, atol=1e-6)
    pm0, tm0, _ = m(pred.cuda(), None, nuclei.cuda())       # target=None -> zeros, like the reference
    assert float(tm0.abs().max()) == 0.0 and torch.allclose(pm0, pm)
    with pytest.raises(ValueError):
        MeanCellExtrator(scale_factor=1.5)


def _torch_means(pred, nuclei):
    """differentiable torch restatement (torch.unique + index_add per image), used only to check gradients"""
    B, C, H, W = pred.shape
    rows = []
    for b in range(B):
        lab = nuclei[b].reshape(-1)
        keep = lab > 0
        if not keep.any():
            continue
        u, inv = torch.unique(lab[keep], return_inverse=True)
        vals = pred[b].reshape(C, -1)[:, keep].t()
        s = torch.zeros((u.numel(), C), dtype=pred.dtype, device=pred.device).index_add(0, inv, vals)
        n = torch.zeros(u.numel(), dtype=pred.dtype, device=pred.device).index_add(0, inv, torch.ones_like(inv, dtype=pred.dtype))
        rows.append(s / n[:, None])
    return torch.cat(rows)


def test_mean_cell_extrator_is_differentiable_like_the_reference():
    """training_step feeds the extractor's output to the cell loss (src/models.py:120-131): gradients must reach pred."""
    from miphei_vit_b200.cells import MeanCellExtrator
    pred, target, nuclei = _case(11, B=3, C=6, S=64, n_cells=20, empty=(1,))
    m = MeanCellExtrator(scale_factor=1.0)
    p = pred.cuda().requires_grad_(True)
    t = target.cuda().requires_grad_(True)
    pm, tm, ids = m(p, t, nuclei.cuda())
    wgt = torch.randn(pm.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
    ((pm - tm) * wgt).sum().backward()
    pr = pred.cuda().double().requires_grad_(True)
    ref = _torch_means(pr, nuclei.cuda())
    (ref * wgt.double()).sum().backward()
    assert torch.allclose(p.grad.double(), pr.grad, rtol=1e-5, atol=1e-8)
    assert torch.allclose(t.grad.double(), -pr.grad, rtol=1e-5, atol=1e-8)
    assert float(p.grad[1].abs().max()) == 0.0            # empty image: no gradient
    # half-resolution path keeps the autograd chain through F.interpolate
    p2 = pred.cuda().requires_grad_(True)
    pm2, _, _ = MeanCellExtrator(scale_factor=0.5)(p2, None, nuclei.cuda())
    pm2.sum().backward()
    assert p2.grad is not None and float(p2.grad.abs().sum()) > 0


def test_cell_means_thousands_of_nuclei_use_the_global_workspace():
    """more nuclei per tile than the shared-memory tables hold (the reference handles any count): 4096 labels per image"""
    S, B, C = 256, 2, 16
    g = torch.Generator().manual_seed(5)
    pred, target = torch.rand((B, C, S, S), generator=g), torch.rand((B, C, S, S), generator=g)
    yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
    lab = (yy // 4) * (S // 4) + (xx // 4) + 1          # 4096 labels of 16 pixels each
    nuclei = torch.stack([lab, lab.flip(0) * 3])
    nuclei[1][:8] = 0
    gid = _check(pred, target, nuclei)
    assert gid.numel() == 4096 + 4096 - 2 * 64
