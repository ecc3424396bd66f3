"""End-to-end training parity on B200: loss, gradients (cosine >= 0.999, north-star bar) and optimiser steps against the
CPU oracle and the committed reference goldens."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model as om  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GRAD_COS_MIN = 0.999


def build(cfg, sd):
    from miphei_vit_b200.generators.mipheivit import get_vitmatte

    m = get_vitmatte("hoptimus0", cfg.img_size, cfg.out_chans, use_lora=True, embed_dim=cfg.embed_dim, depth=cfg.depth,
                     num_heads=cfg.num_heads, hidden=cfg.hidden)
    m.load_state_dict(sd)
    return m.cuda().train()


@pytest.mark.parametrize("name", ["tiny128", "small16ch"])
def test_training_matches_reference_golden(name):
    from miphei_vit_b200 import ops
    from miphei_vit_b200.trainer import Trainer

    g = torch.load(os.path.join(GOLDEN, name + ".pt"), map_location="cpu", weights_only=False)
    cfg = om.Config(**g["config"])
    sd = om.init_state_dict(cfg, seed=g["weight_seed"], perturb=True)
    model = build(cfg, sd)
    x = om.normalize_tiles(om.synthetic_tiles_u8(g["batch"], cfg.img_size, seed=g["input_seed"])).cuda()
    y = om.synthetic_targets(g["batch"], cfg.out_chans, cfg.img_size, seed=g["target_seed"]).cuda()
    tr = Trainer(model, marker_weights=g["marker_weights"], base_lr=g["base_lr"], total_steps=g["total_steps"],
                 warmup_steps=g["warmup_steps"])
    # --- first forward/backward: loss, prediction and gradients vs the reference's own run
    tr.gflat.zero_()
    pred = model(x)
    loss, dpred = ops.loss_fwd_bwd(pred.detach().float().contiguous(), y, tr.marker_weights, lambda_factor=50.0)
    pred.backward(dpred.to(pred.dtype))
    assert abs(loss.item() - g["losses"][0]) < 2e-2 * g["losses"][0]
    assert om.pearson(pred.detach().float().cpu(), g["pred_train0"]) >= 0.999
    grads = {n: p.grad.detach().float().cpu().clone() for n, p in tr.order}
    for k, ref in g["grads0"].items():
        if g["grad0_norms"][k] < 1e-5:
            continue
        c = om.cosine(grads[k], ref)
        assert c >= 0.95, (k, c)  # single small tensors: bf16 rounding amplified by cancelling sums (global bar below)
    keys = [k for k in g["grads0"] if g["grad0_norms"][k] >= 1e-5]
    allc = om.cosine(torch.cat([grads[k].flatten() for k in keys]), torch.cat([g["grads0"][k].flatten() for k in keys]))
    assert allc >= GRAD_COS_MIN, allc
    gn = torch.sqrt(sum((v.double() ** 2).sum() for v in grads.values())).item()
    assert abs(gn - g["grad_norms"][0]) < 3e-2 * g["grad_norms"][0]
    # --- three optimiser steps: losses track the reference trajectory
    tr2_model = build(cfg, sd)
    tr2 = Trainer(tr2_model, marker_weights=g["marker_weights"], base_lr=g["base_lr"], total_steps=g["total_steps"],
                  warmup_steps=g["warmup_steps"])
    losses = [tr2.step(x, y).item() for _ in range(3)]
    for a, b in zip(losses, g["losses"]):
        assert abs(a - b) < 3e-2 * b, (losses, g["losses"])
    assert abs(tr2.norm[0].item() - g["grad_norms"][2]) < 5e-2 * g["grad_norms"][2]
    assert losses[2] < losses[1]
    # parameters moved like the reference's (norm of every trainable tensor after 3 steps)
    psd = tr2_model.state_dict()
    for k in om.trainable_keys(sd):
        n = g["param_norms_after"][k]
        assert abs(float(psd[k].float().norm()) - n) <= 2e-2 * max(n, 1e-3), k


def test_training_gradients_match_oracle_global_cosine():
    """all trainable gradients concatenated (LoRA of every block + decoder): cosine >= 0.999 vs the fp32 oracle."""
    from miphei_vit_b200 import ops
    from miphei_vit_b200.trainer import Trainer

    cfg = om.Config(img_size=256, embed_dim=256, depth=4, num_heads=4, hidden=512, out_chans=16)
    sd = om.init_state_dict(cfg, seed=33, perturb=True)
    x = om.normalize_tiles(om.synthetic_tiles_u8(2, cfg.img_size, seed=5))
    y = om.synthetic_targets(2, cfg.out_chans, cfg.img_size, seed=6)
    w = torch.linspace(1.0, 10.0, cfg.out_chans)
    osd = {k: v.clone() for k, v in sd.items()}
    keys = om.trainable_keys(osd)
    for k in keys:
        osd[k].requires_grad_(True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref_pred = om.miphei_forward(osd, x, cfg, training=True)
    ref_loss = om.weighted_mse_loss(y, ref_pred, w, 50.0)
    gref = dict(zip(keys, torch.autograd.grad(ref_loss, [osd[k] for k in keys])))
    model = build(cfg, sd)
    tr = Trainer(model, marker_weights=w, batch_size=2, total_steps=100)
    tr.gflat.zero_()
    pred = model(x.cuda())
    loss, dpred = ops.loss_fwd_bwd(pred.detach().float().contiguous(), y.cuda(), tr.marker_weights, lambda_factor=50.0)
    pred.backward(dpred.to(pred.dtype))
    got = {n: p.grad.detach().float().cpu() for n, p in tr.order}
    allc = om.cosine(torch.cat([got[k].flatten() for k in keys]), torch.cat([gref[k].flatten() for k in keys]))
    lora = [k for k in keys if ".lora_" in k]
    lorac = om.cosine(torch.cat([got[k].flatten() for k in lora]), torch.cat([gref[k].flatten() for k in lora]))
    print("gradient cosine: all %.6f, LoRA only %.6f; loss %.5f vs %.5f" % (allc, lorac, loss.item(), ref_loss.item()))
    # north-star bar on the full gradient; the LoRA subset alone (tiny norms, reached through the whole bf16 decoder
    # backward) is reported and held to 0.99
    assert allc >= GRAD_COS_MIN and lorac >= 0.99
    assert abs(loss.item() - ref_loss.item()) < 2e-2 * ref_loss.item()
