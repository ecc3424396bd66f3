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
GRAD_COS_MIN = 0.999        # north-star bar: all trainable gradients, and the LoRA matrices (the only encoder gradients) alone
PER_TENSOR_COS_MIN = 0.99   # each of the ~300 trainable tensors on its own


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
        assert c >= PER_TENSOR_COS_MIN, (k, c)  # every single trainable tensor
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
    # north-star bar on the full gradient AND on the LoRA subset alone (the only encoder gradients: they carry the error of
    # the whole encoder backward and of the decoder's data gradient)
    assert allc >= GRAD_COS_MIN and lorac >= GRAD_COS_MIN
    assert abs(loss.item() - ref_loss.item()) < 2e-2 * ref_loss.item()


def _oracle_grads(cfg, sd, x, y, w):
    osd = {k: v.clone() for k, v in sd.items()}
    keys = om.trainable_keys(osd)
    for k in keys:
        osd[k].requires_grad_(True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    pred = om.miphei_forward(osd, x, cfg, training=True)
    loss = om.weighted_mse_loss(y, pred, w, 50.0)
    return keys, dict(zip(keys, torch.autograd.grad(loss, [osd[k] for k in keys]))), loss.detach(), pred.detach()


def _cos_groups(keys, got, ref):
    cat = lambda d, ks: torch.cat([d[k].flatten() for k in ks])  # noqa: E731
    lora = [k for k in keys if ".lora_" in k]
    dec = [k for k in keys if ".lora_" not in k]
    return om.cosine(cat(got, keys), cat(ref, keys)), om.cosine(cat(got, dec), cat(ref, dec)), om.cosine(cat(got, lora), cat(ref, lora))


def test_training_full_depth_matches_oracle():
    """BASELINE configs[2] geometry — the real 40-block ViT-g/14 (1.13 B parameters), 256 px, 16 channels — one training
    forward / backward at batch 2 against the fp32 CPU oracle: loss, prediction, and the gradient cosine of all trainables,
    of the decoder, of the LoRA matrices alone and of the LoRA matrices block by block (the error of a 40-block bf16
    backward accumulates towards block 0)."""
    from miphei_vit_b200 import ops
    from miphei_vit_b200.trainer import Trainer

    cfg = om.Config()
    sd = om.init_state_dict(cfg, seed=0, perturb=True)
    x = om.normalize_tiles(om.synthetic_tiles_u8(2, cfg.img_size, seed=77))
    y = om.synthetic_targets(2, cfg.out_chans, cfg.img_size, seed=78)
    w = torch.linspace(1.0, 10.6, cfg.out_chans)
    keys, gref, ref_loss, ref_pred = _oracle_grads(cfg, sd, x, y, w)
    model = build(cfg, sd)
    tr = Trainer(model, marker_weights=w, batch_size=2, total_steps=100)
    tr.gflat.zero_()
    pred = model(x.cuda())
    loss, dpred = ops.loss_fwd_bwd(pred.detach().float().contiguous(), y.cuda(), tr.marker_weights, lambda_factor=50.0)
    pred.backward(dpred)
    got = {n: p.grad.detach().float().cpu() for n, p in tr.order}
    allc, decc, lorac = _cos_groups(keys, got, gref)
    per_block = []
    for i in range(cfg.depth):
        ks = [k for k in keys if ".blocks.%d.attn" % i in k]
        per_block.append(om.cosine(torch.cat([got[k].flatten() for k in ks]), torch.cat([gref[k].flatten() for k in ks])))
    print("depth-40 training parity: loss %.5f vs %.5f, pearson(pred) %.6f, gradient cosine all %.6f decoder %.6f LoRA %.6f; "
          "per-block LoRA min %.5f (block %d), block 0 %.5f, block 39 %.5f" % (
              loss.item(), ref_loss.item(), om.pearson(pred.detach().float().cpu(), ref_pred), allc, decc, lorac,
              min(per_block), per_block.index(min(per_block)), per_block[0], per_block[-1]))
    assert abs(loss.item() - ref_loss.item()) < 2e-2 * ref_loss.item()
    assert om.pearson(pred.detach().float().cpu(), ref_pred) >= 0.999
    assert allc >= GRAD_COS_MIN and decc >= GRAD_COS_MIN and lorac >= GRAD_COS_MIN, (allc, decc, lorac)
    assert min(per_block) >= 0.995, per_block


def test_plain_autograd_route_without_trainer():
    """The advertised drop-in route: generator(x) -> loss -> loss.backward() with a stock torch.optim.Adam over
    generator.parameters() (src/models.py:134-139, 361-362) and NO Trainer / flat buffers: gradients arrive in p.grad
    through AccumulateGrad and match the oracle; frozen parameters get none."""
    cfg = om.Config(img_size=256, embed_dim=256, depth=4, num_heads=4, hidden=512, out_chans=16)
    sd = om.init_state_dict(cfg, seed=33, perturb=True)
    x = om.normalize_tiles(om.synthetic_tiles_u8(2, cfg.img_size, seed=5))
    y = om.synthetic_targets(2, cfg.out_chans, cfg.img_size, seed=6)
    w = torch.linspace(1.0, 10.0, cfg.out_chans)
    keys, gref, ref_loss, _ = _oracle_grads(cfg, sd, x, y, w)
    model = build(cfg, sd)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, betas=(0.5, 0.999), eps=1e-7)
    pred = model(x.cuda())
    loss = om.weighted_mse_loss(y.cuda(), pred, w.cuda(), 50.0)
    opt.zero_grad()
    loss.backward()
    named = dict(model.named_parameters())
    got = {k: named[k].grad.detach().float().cpu() for k in keys}
    assert all(p.grad is None for n, p in named.items() if n not in keys)
    allc, decc, lorac = _cos_groups(keys, got, gref)
    print("plain autograd route: gradient cosine all %.6f decoder %.6f LoRA %.6f" % (allc, decc, lorac))
    assert abs(loss.item() - ref_loss.item()) < 2e-2 * ref_loss.item()
    assert allc >= GRAD_COS_MIN and lorac >= GRAD_COS_MIN
    before = {k: named[k].detach().clone() for k in keys}
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()
    assert any(not torch.equal(before[k], named[k].detach()) for k in keys)
    with torch.no_grad():  # the engine notices the update (version counters) and re-packs
        p2 = model(x.cuda())
    assert not torch.equal(p2, pred.detach())


def test_reference_training_route_autocast_fp16_gradscaler():
    """What Lightning's "16-mixed" manual optimisation does around the generator (src/models.py:134-138, 361-371;
    configs/config.yaml:23): torch.autocast(fp16) forward, scaled backward, unscale + clip_grad_norm_(1.0), torch Adam
    (0.5, 0.999, 1e-7), LambdaLR per step — three steps against the golden trajectory of the reference's own fp32 run."""
    from miphei_vit_b200.trainer import lr_lambda

    g = torch.load(os.path.join(GOLDEN, "small16ch.pt"), map_location="cpu", weights_only=False)
    cfg = om.Config(**g["config"])
    sd = om.init_state_dict(cfg, seed=g["weight_seed"], perturb=True)
    model = build(cfg, sd)
    x = om.normalize_tiles(om.synthetic_tiles_u8(g["batch"], cfg.img_size, seed=g["input_seed"])).cuda()
    y = om.synthetic_targets(g["batch"], cfg.out_chans, cfg.img_size, seed=g["target_seed"]).cuda()
    w = g["marker_weights"].cuda()
    opt = torch.optim.Adam(model.parameters(), lr=g["base_lr"], betas=(0.5, 0.999), eps=1e-7)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: lr_lambda(s, g["total_steps"], g["warmup_steps"]))
    scaler = torch.amp.GradScaler("cuda")
    losses, norms = [], []
    for _ in range(3):
        with torch.autocast("cuda", dtype=torch.float16):
            pred = model(x)
            assert pred.dtype == torch.float16
            loss = om.weighted_mse_loss(y, pred.float(), w, 50.0)
        opt.zero_grad()
        scaler.scale(loss).backward()
        scaler.unscale_(opt)
        norms.append(float(torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)))
        scaler.step(opt)
        scaler.update()
        sched.step()
        losses.append(loss.item())
    print("autocast route losses", losses, "golden", g["losses"], "norms", norms, g["grad_norms"])
    for a, b in zip(losses, g["losses"]):
        assert abs(a - b) < 3e-2 * b, (losses, g["losses"])
    for a, b in zip(norms, g["grad_norms"]):
        assert abs(a - b) < 5e-2 * b, (norms, g["grad_norms"])
    psd = model.state_dict()
    for k in om.trainable_keys(sd):
        n = g["param_norms_after"][k]
        assert abs(float(psd[k].float().norm()) - n) <= 2e-2 * max(n, 1e-3), k


def test_stale_tape_raises_instead_of_returning_wrong_gradients():
    cfg = om.Config(img_size=128, embed_dim=128, depth=2, num_heads=2, hidden=256, out_chans=3)
    model = build(cfg, om.init_state_dict(cfg, seed=2, perturb=True))
    x = om.normalize_tiles(om.synthetic_tiles_u8(2, cfg.img_size, seed=1)).cuda()
    p1 = model(x)
    p2 = model(x)  # same batch size: overwrites the activations p1's graph saved
    with pytest.raises(RuntimeError, match="stale"):
        p1.sum().backward()
    p2.sum().backward()
    with pytest.raises(RuntimeError):
        p2.sum().backward()  # second backward through a consumed graph
    # an eval / no-grad call at the same batch size between forward and backward does NOT disturb the tape
    p3 = model(x)
    model.eval()
    with torch.no_grad():
        model(x)
    model.train()
    p3.sum().backward()


def test_trainer_graph_step_matches_autograd_step():
    """Trainer.step (captured CUDA graphs, no autograd), its eager twin and Trainer.step_autograd (model(x) / backward())
    are the same optimisation: same gradients after the first step, same loss trajectory, same BatchNorm buffers and step
    counters over 4 steps. (Weight-gradient sums are fp32 atomics: run-to-run last-bit noise, which Adam's early sign-like
    updates amplify on near-zero gradients — parameters are compared by direction, not bit by bit.)"""
    from miphei_vit_b200.trainer import Trainer

    cfg = om.Config(img_size=128, embed_dim=128, depth=3, num_heads=2, hidden=256, out_chans=5)
    sd = om.init_state_dict(cfg, seed=8, perturb=True)
    x = om.normalize_tiles(om.synthetic_tiles_u8(4, cfg.img_size, seed=3)).cuda()
    y = om.synthetic_targets(4, cfg.out_chans, cfg.img_size, seed=4).cuda()
    w = torch.linspace(1.0, 3.0, cfg.out_chans)
    runs = []
    for mode in ("graph", "eager", "autograd"):
        model = build(cfg, sd)
        tr = Trainer(model, marker_weights=w, base_lr=2e-3, total_steps=50, warmup_steps=2, use_graph=(mode == "graph"))
        step = tr.step_autograd if mode == "autograd" else tr.step
        losses, g1 = [], None
        for it in range(4):
            losses.append(float(step(x, y).item()))
            if it == 1:  # the graph path replays a captured graph from its second step on
                g1 = tr.gflat.detach().clone()
        runs.append((losses, tr.flat.detach().clone(), int(tr.step_dev.item()), g1,
                     {k: v.clone() for k, v in model.state_dict().items() if "running" in k or "num_batches" in k}))
    ref = runs[0]
    for losses, flat, nstep, g1, bufs in runs[1:]:
        assert nstep == 4
        for a, b in zip(losses, ref[0]):
            assert abs(a - b) < 2e-3 * abs(b), (losses, ref[0])
        assert om.cosine(g1.cpu(), ref[3].cpu()) > 0.9995
        assert om.cosine(flat.cpu(), ref[1].cpu()) > 0.9995
        for k, v in bufs.items():  # (Adam's early sign-like updates amplify last-bit gradient noise: statistics compared by direction)
            if "num_batches" in k:
                assert int(v) == int(ref[4][k]), k
            else:
                assert om.cosine(v.float().cpu(), ref[4][k].float().cpu()) > 0.995, k
    assert ref[0][3] < ref[0][1]
    assert int(ref[4]["decoder.fusion_blks.0.conv.bn.num_batches_tracked"]) == 4
