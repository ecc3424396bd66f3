"""Generates tests/golden/*.pt by running the REFERENCE's own code (decoder / heads / LoRA / loss imported verbatim
from /root/reference through oracle/ref_import.py, wrapped around the timm-shaped ViT) on seeded synthetic inputs.

Run here (the container that has /root/reference):   python tests/golden/make_golden.py
The fixtures travel to the GPU box, where /root/reference does not exist.
Weights are NOT stored: they are regenerated from the seed by oracle.model.init_state_dict (torch CPU RNG, same image).
"""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import model as om  # noqa: E402
from oracle import ref_import  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (config kwargs, weight seed, batch)
    "tiny128": (dict(img_size=128, embed_dim=128, depth=2, num_heads=2, hidden=256, out_chans=3), 21, 2),
    "small16ch": (dict(img_size=128, embed_dim=256, depth=3, num_heads=4, hidden=512, out_chans=16), 22, 2),
}


def main():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for name, (ckw, seed, B) in CASES.items():
        cfg = om.Config(**ckw)
        sd = om.init_state_dict(cfg, seed=seed, perturb=True)
        x_u8 = om.synthetic_tiles_u8(B, cfg.img_size, seed=seed + 100)
        x = om.normalize_tiles(x_u8)
        y = om.synthetic_targets(B, cfg.out_chans, cfg.img_size, seed=seed + 200)
        w = torch.linspace(1.0, 4.0, cfg.out_chans)
        ref = ref_import.build_reference_model(cfg, copy.deepcopy(sd))
        out = {"config": cfg.as_dict(), "weight_seed": seed, "batch": B, "input_seed": seed + 100,
               "target_seed": seed + 200, "marker_weights": w}
        ref.eval()
        with torch.no_grad():
            out["features_eval"] = ref.encoder(x).clone()
            out["pred_eval"] = ref(x).clone()
        # one training step with the reference loss + torch Adam exactly as src/models.py:134-139,361-362
        ref.train()
        loss_mod = ref_import.load()["loss"].WeightedMSELoss(50.0, w)
        params = [p for p in ref.parameters() if p.requires_grad]
        base_lr, total, warm = 2e-4 * B ** 0.5, 1000, 2
        opt = torch.optim.Adam(ref.parameters(), lr=base_lr, betas=(0.5, 0.999), eps=1e-7)
        sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: om.lr_lambda(s, total, warm))
        losses, gnorms = [], []
        grads0 = None
        for it in range(3):
            opt.zero_grad()
            pred = ref(x)
            loss = loss_mod(y, pred)
            loss.backward()
            if it == 0:
                out["pred_train0"] = pred.detach().clone()
                grads0 = {n: p.grad.detach().clone() for n, p in ref.named_parameters() if p.requires_grad}
            gnorms.append(float(torch.nn.utils.clip_grad_norm_(params, 1.0)))
            opt.step()
            sched.step()
            losses.append(float(loss.detach()))
        out["losses"] = losses
        out["grad_norms"] = gnorms
        out["base_lr"], out["total_steps"], out["warmup_steps"] = base_lr, total, warm
        # keep a handful of full gradients (first/last block LoRA, one conv of each kind) + every grad's norm
        keep = [k for k in grads0 if (".blocks.0." in k or ".blocks.%d." % (cfg.depth - 1) in k)]
        keep += ["decoder.convstream.convs.0.conv.weight", "decoder.fusion_blks.3.conv.conv.weight",
                 "decoder.fusion_blks.3.conv.bn.weight", "decoder.segmentation_head_0.1.weight",
                 "decoder.segmentation_head_0.0.psi.3.weight", "decoder.segmentation_head_1.0.psi.0.weight"]
        out["grads0"] = {k: grads0[k] for k in keep}
        out["grad0_norms"] = {k: float(g.norm()) for k, g in grads0.items()}
        fsd = ref.state_dict()
        out["param_norms_after"] = {k: float(v.float().norm()) for k, v in fsd.items()}
        out["bn_after"] = {k: v.clone() for k, v in fsd.items() if "fusion_blks.3.conv.bn.running" in k}
        path = os.path.join(HERE, name + ".pt")
        # predictions in fp16 would lose the parity digits: keep fp32, they are small
        torch.save(out, path)
        print(name, "->", path, os.path.getsize(path) // 1024, "KiB", "loss", losses, "gnorm", gnorms)


if __name__ == "__main__":
    main()
