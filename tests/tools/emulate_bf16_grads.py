"""CPU experiment (test tooling): which bf16 storage points of the training path cost LoRA-gradient accuracy?

Re-runs the oracle graph with identity nodes that round the FORWARD value and / or the GRADIENT flowing through them to
bf16 at the places where the CUDA path stores a bf16 tensor, and prints the cosine of the LoRA / decoder gradients
against the clean fp32 oracle.  Usage: python tests/tools/emulate_bf16_grads.py [tiny|small|mid] [flag ...]
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import model as om  # noqa: E402


class _Round(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, fwd, bwd):
        ctx.bwd = bwd
        if fwd == 2:
            return t.half().float()
        return t.bfloat16().float() if fwd else t.clone()

    @staticmethod
    def backward(ctx, g):
        return (g.bfloat16().float() if ctx.bwd else g), None, None


FLAGS = set()


def R(t, name):
    """storage point `name`: flags name+'.f' (forward value rounded) / name+'.b' (gradient rounded)."""
    f, b = (name + ".f") in FLAGS or "all.f" in FLAGS, (name + ".b") in FLAGS or "all.b" in FLAGS
    if (name + ".h") in FLAGS or "all.h" in FLAGS:
        f = 2
    if not (f or b):
        return t
    return _Round.apply(t, f, b)


def vit(sd, x, cfg):
    v = "encoder.vit."
    D, nh = cfg.embed_dim, cfg.num_heads
    hd = D // nh
    B = x.shape[0]
    t = F.conv2d(x, sd[v + "patch_embed.proj.weight"], sd[v + "patch_embed.proj.bias"], stride=om.PATCH)
    t = t.flatten(2).transpose(1, 2) + sd[v + "pos_embed"]
    t = torch.cat([sd[v + "cls_token"].expand(B, -1, -1), sd[v + "reg_token"].expand(B, -1, -1), t], dim=1)
    N = t.shape[1]
    for i in range(cfg.depth):
        b = v + "blocks.%d." % i
        h = R(F.layer_norm(t, (D,), sd[b + "norm1.weight"], sd[b + "norm1.bias"], om.LN_EPS), "ln1")
        qkv = F.linear(h, sd[b + "attn.qkv.qkv.weight"], sd[b + "attn.qkv.qkv.bias"])
        tq = R(h @ sd[b + "attn.qkv.lora_q.A"], "loraT")
        tv = R(h @ sd[b + "attn.qkv.lora_v.A"], "loraT")
        qkv = torch.cat([qkv[..., :D] + tq @ sd[b + "attn.qkv.lora_q.B"], qkv[..., D:2 * D],
                         qkv[..., 2 * D:] + tv @ sd[b + "attn.qkv.lora_v.B"]], dim=-1)
        qkv = R(qkv, "qkv")
        qkv = qkv.reshape(B, N, 3, nh, hd).permute(2, 0, 3, 1, 4)
        q, k, val = qkv[0], qkv[1], qkv[2]
        att = R(((q @ k.transpose(-2, -1)) * (hd ** -0.5)).softmax(dim=-1), "P")
        o = R((att @ val).transpose(1, 2).reshape(B, N, D), "o")
        o = F.linear(o, sd[b + "attn.proj.weight"], sd[b + "attn.proj.bias"])
        t = t + sd[b + "ls1.gamma"] * o
        t = R(t, "resid")
        h = R(F.layer_norm(t, (D,), sd[b + "norm2.weight"], sd[b + "norm2.bias"], om.LN_EPS), "ln2")
        h = R(F.linear(h, sd[b + "mlp.fc1.weight"], sd[b + "mlp.fc1.bias"]), "h")
        x1, x2 = h.chunk(2, dim=-1)
        h = R(F.silu(x1) * x2, "u")
        h = F.linear(h, sd[b + "mlp.fc2.weight"], sd[b + "mlp.fc2.bias"])
        t = t + sd[b + "ls2.gamma"] * h
        t = R(t, "resid")
    return R(F.layer_norm(t, (D,), sd[v + "norm.weight"], sd[v + "norm.bias"], om.LN_EPS), "tok")


def forward(sd, x, cfg):
    tok = vit(sd, x, cfg)
    B, g = x.shape[0], cfg.grid
    f = tok[:, om.NUM_PREFIX:].permute(0, 2, 1).reshape(B, cfg.embed_dim, g, g)
    tgt = cfg.img_size / 16
    f = R(F.interpolate(f, scale_factor=(tgt / g, tgt / g), mode="bicubic"), "fmap")
    d = "decoder."
    x = R(x, "img")
    details = [x]
    y = x
    for i in range(3):
        z = F.conv2d(y, sd[d + "convstream.convs.%d.conv.weight" % i], None, stride=2, padding=1)
        z = R(z, "z")
        y = R(F.relu(om._bn(z, sd, d + "convstream.convs.%d.bn." % i, True)), "y")
        details.append(y)
    for i in range(4):
        up = R(F.interpolate(f, scale_factor=2, mode="bilinear", align_corners=False), "up")
        z = F.conv2d(torch.cat([details[3 - i], up], dim=1), sd[d + "fusion_blks.%d.conv.conv.weight" % i], None, padding=1)
        z = R(z, "z")
        f = R(F.relu(om._bn(z, sd, d + "fusion_blks.%d.conv.bn." % i, True)), "y%d" % i)
    outs = []
    h = 0
    while (d + "segmentation_head_%d.1.weight" % h) in sd:
        p = d + "segmentation_head_%d." % h
        a = F.conv2d(f, sd[p + "0.psi.0.weight"], sd[p + "0.psi.0.bias"])
        a = R(F.relu(om._bn(a, sd, p + "0.psi.1.", True)), "hr")
        u = R(F.conv2d(a, sd[p + "0.psi.3.weight"], sd[p + "0.psi.3.bias"]), "hu")
        gate = R(torch.sigmoid(u), "gate")
        fg = R(f * gate, "fg")
        s = R(F.conv2d(fg, sd[p + "1.weight"], sd[p + "1.bias"], padding=1), "hs")
        outs.append(torch.tanh(s))
        h += 1
    return torch.cat(outs, dim=1)


def grads(sd, x, y, w, cfg):
    sd = {k: v.clone() for k, v in sd.items()}
    keys = om.trainable_keys(sd)
    for k in keys:
        sd[k].requires_grad_(True)
    sd0 = sd
    sd = dict(sd)
    for k in list(sd):
        if not sd[k].is_floating_point() or sd[k].dim() < 2:
            continue
        if k.startswith("decoder.") and k.endswith("weight"):
            sd[k] = R(sd[k], "wdec")
        elif ".lora_" in k:
            sd[k] = R(sd[k], "wlora")
        elif k.startswith("encoder.") and k.endswith("weight"):
            sd[k] = R(sd[k], "wvit")
    pred = forward(sd, x, cfg)
    loss = om.weighted_mse_loss(y, pred, w, 50.0)
    return keys, dict(zip(keys, torch.autograd.grad(loss, [sd0[k] for k in keys])))


def report(keys, g, ref, label):
    lora = [k for k in keys if ".lora_" in k]
    dec = [k for k in keys if ".lora_" not in k]
    cat = lambda d, ks: torch.cat([d[k].flatten() for k in ks])  # noqa: E731
    line = "%-28s all %.6f  dec %.6f  LoRA %.6f" % (label, om.cosine(cat(g, keys), cat(ref, keys)),
                                                    om.cosine(cat(g, dec), cat(ref, dec)), om.cosine(cat(g, lora), cat(ref, lora)))
    nb = len(lora) // 4
    per = []
    for i in (0, nb - 1):
        ks = [k for k in lora if ".blocks.%d." % i in k]
        per.append("blk%d %.5f" % (i, om.cosine(cat(g, ks), cat(ref, ks))))
    print(line + "   " + " ".join(per), flush=True)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    cfgs = {"tiny": om.Config(img_size=128, embed_dim=128, depth=2, num_heads=2, hidden=256, out_chans=3),
            "small": om.Config(img_size=256, embed_dim=256, depth=4, num_heads=4, hidden=512, out_chans=16),
            "mid": om.Config(img_size=256, embed_dim=384, depth=12, num_heads=6, hidden=1024, out_chans=16)}
    cfg = cfgs[which]
    torch.set_num_threads(os.cpu_count() or 1)
    sd = om.init_state_dict(cfg, seed=33, perturb=True)
    x = om.normalize_tiles(om.synthetic_tiles_u8(2, cfg.img_size, seed=5))
    y = om.synthetic_targets(2, cfg.out_chans, cfg.img_size, seed=6)
    w = torch.linspace(1.0, 10.0, cfg.out_chans)
    FLAGS.clear()
    keys, ref = grads(sd, x, y, w, cfg)
    trials = sys.argv[2:] or ["all.b", "all.f", "all.f,all.b", "hs.b", "fg.b", "gate.b", "hu.b", "hr.b", "y3.b", "y2.b", "y1.b",
                              "y0.b", "z.b", "up.b", "fmap.b", "tok.b", "resid.b", "u.b", "h.b", "ln2.b", "o.b", "P.b", "qkv.b",
                              "ln1.b", "loraT.b", "loraT.f"]
    for t in trials:
        FLAGS.clear()
        FLAGS.update(t.split(","))
        _, g = grads(sd, x, y, w, cfg)
        report(keys, g, ref, t)


if __name__ == "__main__":
    main()
