"""Pins the oracle restatement (oracle/model.py) to the reference's own code, imported from /root/reference.
Skipped where the reference tree is absent (the GPU box); the committed goldens cover that case."""
import copy
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model as om  # noqa: E402
from oracle import ref_import  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present")

CFG = om.Config(img_size=128, embed_dim=128, depth=2, num_heads=2, hidden=256, out_chans=3)


def _inputs(B=2):
    x = om.normalize_tiles(om.synthetic_tiles_u8(B, CFG.img_size, seed=11))
    y = om.synthetic_targets(B, CFG.out_chans, CFG.img_size, seed=12)
    return x, y


def test_state_dict_keys_match_reference():
    sd = om.init_state_dict(CFG, seed=3)
    ref = ref_import.build_reference_model(CFG)
    rsd = ref.state_dict()
    assert set(sd.keys()) == set(rsd.keys())
    for k in sd:
        assert tuple(sd[k].shape) == tuple(rsd[k].shape), k
    train = {n for n, p in ref.named_parameters() if p.requires_grad}
    assert train == set(om.trainable_keys(sd))


def test_forward_eval_matches_reference():
    sd = om.init_state_dict(CFG, seed=3)
    ref = ref_import.build_reference_model(CFG, sd).eval()
    x, _ = _inputs()
    with torch.no_grad():
        want = ref(x)
        got = om.miphei_forward(sd, x, CFG, training=False)
    assert got.shape == want.shape == (2, 3, 128, 128)
    assert (got - want).abs().max().item() < 2e-5
    assert want.abs().max().item() > 1e-3


def test_train_step_matches_reference():
    """forward in train mode (BN batch stats), WeightedMSE, backward, clip, Adam(0.5, 0.999, 1e-7), LambdaLR."""
    sd = om.init_state_dict(CFG, seed=5)
    ref = ref_import.build_reference_model(CFG, copy.deepcopy(sd)).train()
    x, y = _inputs()
    w = torch.tensor([1.0, 2.5, 0.7])
    loss_mod = ref_import.load()["loss"].WeightedMSELoss(50.0, w)
    params = [p for p in ref.parameters() if p.requires_grad]
    base_lr, total = 2e-4 * 2 ** 0.5, 1000
    opt = torch.optim.Adam(ref.parameters(), lr=base_lr, betas=(0.5, 0.999), eps=1e-7)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: om.lr_lambda(s, total, 2))
    state = {}
    for it in range(3):
        opt.zero_grad()
        loss_ref = loss_mod(y, ref(x))
        loss_ref.backward()
        gn_ref = torch.nn.utils.clip_grad_norm_(params, 1.0)
        ref_grads = {n: p.grad.clone() for n, p in ref.named_parameters() if p.requires_grad}
        opt.step()
        sched.step()
        loss, raw, gn, _ = om.train_step(sd, state, x, y, CFG, w, base_lr, total, 50.0, warmup_steps=2)
        assert abs(loss.item() - loss_ref.item()) < 1e-4 * max(1.0, abs(loss_ref.item()))
        assert abs(gn.item() - gn_ref.item()) < 1e-3 * gn_ref.item()
        for n, g in ref_grads.items():
            # reference grads are post-clip; oracle returns raw grads. A conv bias feeding a train-mode BatchNorm
            # (psi.0.bias) has an analytically zero gradient: only rounding noise, skipped.
            coef = min(1.0, 1.0 / (gn_ref.item() + 1e-6))
            if g.norm().item() < 1e-6 * min(1.0, gn_ref.item()):
                continue
            assert om.cosine(raw[n] * coef, g) > 0.99999, n
    rsd = ref.state_dict()
    for k in sd:
        if sd[k].dtype.is_floating_point:
            assert (sd[k].detach() - rsd[k]).abs().max().item() < 1e-4, k
        else:
            assert int(sd[k]) == int(rsd[k]), k


def test_losses_match_reference():
    loss = ref_import.load()["loss"]
    g = torch.Generator().manual_seed(0)
    a = torch.rand((2, 3, 16, 16), generator=g) * 2 - 1
    b = torch.rand((2, 3, 16, 16), generator=g) * 2 - 1
    assert abs(loss.get_mae_loss(3.0)(a, b).item() - om.mae_loss(a, b, 3.0).item()) < 1e-6
    assert abs(loss.get_mse_loss(3.0)(a, b).item() - om.mse_loss(a, b, 3.0).item()) < 1e-6
    assert abs(loss.L1_L2_Loss(3.0)(b, a).item() - om.l1_l2_loss(b, a, 3.0).item()) < 1e-6


def test_lr_lambda_matches_reference_source():
    """utils.py imports hydra/wandb and cannot be imported; the scheduler body is executed from its source text."""
    src = open(os.path.join(ref_import.REF_ROOT, "src", "utils.py")).read()
    start = src.index("def pix2pix_lr_scheduler")
    end = src.index("\ndef ", start + 10) if "\ndef " in src[start + 10:] else len(src)
    ns = {"torch": torch}
    exec(src[start:end], ns)
    for total in (2000, 999, 801):
        fn = ns["pix2pix_lr_scheduler"](total, 400, total // 2)  # call site: src/models.py:363-369
        for s in range(0, total + 5, 7):
            assert abs(fn(s) - om.lr_lambda(s, total, 400)) < 1e-12, (total, s)
