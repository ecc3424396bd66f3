"""The oracle restatement must reproduce the golden vectors generated from the reference's own code
(tests/golden/make_golden.py). Runs everywhere (CPU only, no /root/reference needed)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model as om  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), map_location="cpu", weights_only=False)


@pytest.mark.parametrize("name", ["tiny128", "small16ch"])
def test_oracle_reproduces_reference_golden(name):
    g = load_golden(name)
    cfg = om.Config(**g["config"])
    sd = om.init_state_dict(cfg, seed=g["weight_seed"], perturb=True)
    x = om.normalize_tiles(om.synthetic_tiles_u8(g["batch"], cfg.img_size, seed=g["input_seed"]))
    y = om.synthetic_targets(g["batch"], cfg.out_chans, cfg.img_size, seed=g["target_seed"])
    with torch.no_grad():
        col = {}
        pred = om.miphei_forward(sd, x, cfg, training=False, collect=col)
    assert (col["features"] - g["features_eval"]).abs().max().item() < 2e-5
    assert (pred - g["pred_eval"]).abs().max().item() < 2e-5
    state = {}
    for it in range(3):
        loss, raw, gn, pred_t = om.train_step(sd, state, x, y, cfg, g["marker_weights"], g["base_lr"],
                                              g["total_steps"], 50.0, warmup_steps=g["warmup_steps"])
        assert abs(loss.item() - g["losses"][it]) < 2e-4 * abs(g["losses"][it]), (it, loss.item())
        assert abs(gn.item() - g["grad_norms"][it]) < 1e-3 * g["grad_norms"][it]
        if it == 0:
            assert (pred_t - g["pred_train0"]).abs().max().item() < 2e-5
            for k, ref in g["grads0"].items():
                if g["grad0_norms"][k] < 1e-6:
                    continue
                assert om.cosine(raw[k], ref) > 0.99999, k
            for k, n in g["grad0_norms"].items():
                if n > 1e-4:
                    assert abs(float(raw[k].norm()) - n) < 2e-3 * n, k
    for k, n in g["param_norms_after"].items():
        assert abs(float(sd[k].detach().float().norm()) - n) <= 1e-4 * max(1.0, n), k
    for k, v in g["bn_after"].items():
        assert (sd[k] - v).abs().max().item() < 1e-5, k
