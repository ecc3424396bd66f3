import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import model as om
from miphei_vit_b200 import ops
from miphei_vit_b200.trainer import Trainer
from miphei_vit_b200.generators.mipheivit import get_vitmatte
name = sys.argv[1] if len(sys.argv) > 1 else "tiny128"
g = torch.load("tests/golden/%s.pt" % name, map_location="cpu", weights_only=False)
cfg = om.Config(**g["config"])
sd = om.init_state_dict(cfg, seed=g["weight_seed"], perturb=True)
m = get_vitmatte("hoptimus0", cfg.img_size, cfg.out_chans, use_lora=True, embed_dim=cfg.embed_dim, depth=cfg.depth, num_heads=cfg.num_heads, hidden=cfg.hidden)
m.load_state_dict(sd); m = m.cuda().train()
x = om.normalize_tiles(om.synthetic_tiles_u8(g["batch"], cfg.img_size, seed=g["input_seed"]))
y = om.synthetic_targets(g["batch"], cfg.out_chans, cfg.img_size, seed=g["target_seed"])
osd = {k: v.clone() for k, v in sd.items()}
keys = om.trainable_keys(osd)
for k in keys: osd[k].requires_grad_(True)
col = {}
rp = om.miphei_forward(osd, x, cfg, training=True, collect=col)
rl = om.weighted_mse_loss(y, rp, g["marker_weights"], 50.0)
feat = col["features"]; feat.retain_grad()
gref = dict(zip(keys, torch.autograd.grad(rl, [osd[k] for k in keys], retain_graph=True)))
dfeat_ref = torch.autograd.grad(rl, feat)[0]
tr = Trainer(m, marker_weights=g["marker_weights"], base_lr=g["base_lr"], total_steps=1000, warmup_steps=2)
tr.gflat.zero_()
# hook the feature gradient
import miphei_vit_b200.autograd as ag
orig = ag.encoder_backward_head
cap = {}
def wrapped(eng, tape, dmap):
    cap["dmap"] = dmap.float().cpu().clone()
    return orig(eng, tape, dmap)
ag.encoder_backward_head = wrapped
pred = m(x.cuda())
loss, dpred = ops.loss_fwd_bwd(pred.detach().float().contiguous(), y.cuda(), tr.marker_weights, lambda_factor=50.0)
pred.backward(dpred.to(pred.dtype))
print("loss", loss.item(), rl.item(), "pred pearson", om.pearson(pred.detach().float().cpu(), rp.detach()))
print("dfeat cosine", om.cosine(cap["dmap"].permute(0, 3, 1, 2), dfeat_ref))
allg = torch.cat([p.grad.float().cpu().flatten() for n, p in tr.order]); allr = torch.cat([gref[n].flatten() for n, p in tr.order])
print("GLOBAL cosine", om.cosine(allg, allr))
for n, p in tr.order:
    if "segmentation_head" in n and "head_0" not in n and "head_1." not in n: continue
    if True:
        print("%-60s cos %.5f  |g| %.3e ref %.3e" % (n, om.cosine(p.grad.float().cpu(), gref[n]), float(p.grad.norm()), float(gref[n].norm())))
