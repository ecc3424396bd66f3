"""CPU ORACLE — test infrastructure, not product code.

A plain-PyTorch fp32, purely functional restatement of the MIPHEI-ViT generator hot path, written against the
reference's state-dict key names.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it; the product package (miphei-vit_b200/) never does.

What each function follows (all paths relative to /root/reference):
  vit_forward        timm==1.0.15 `vit_giant_patch14_reg4_dinov2` as instantiated at
                     src/generators/foundation_models.py:53-57 (timm is an un-vendored third-party dependency,
                     requirements.txt:17; its published semantics are restated in SURVEY.md Appendix B) with the
                     LoRA-wrapped qkv of src/generators/lora.py:16-18,29-33.
  encoder_forward    Encoder.forward, src/generators/mipheivit.py:153-163.
  decoder_forward    Detail_Capture.forward 207-220, ConvStream 66-73, Fusion_Block 88-93, Basic_Conv3x3 36-41 (same
                     file); SegmentationHead / AttentionBlock src/generators/unet.py:407-438.
  init_state_dict    timm init_weights('') + lora.py:11-13 + initialize_decoder_head unet.py:522-531, then the
                     perturbation SURVEY.md fact 9 calls for (LayerScale gamma, LoRA B, BN running stats).
  weighted_mse_loss  WeightedMSELoss.forward src/loss.py:54-57;  mae/mse/l1_l2: src/loss.py:35-44,113-123.
  lr_lambda          pix2pix_lr_scheduler src/utils.py:217-230.
  train_step         ModelModule.training_step optimiser section src/models.py:134-139 with
                     configure_optimizers 361-362 (Adam betas (0.5, 0.999), eps 1e-7, no weight decay).

Parity pinning: the decoder / LoRA / loss restatements are checked against the reference's own code imported from
/root/reference (oracle/ref_import.py; tests/test_oracle_vs_reference.py) and against committed golden vectors
generated from it (tests/golden/).  The ViT restatement is pinned against transformers' independent
Dinov2WithRegistersModel (tests/test_oracle_vit_vs_hf.py) because timm is not installable here; the reference holds
no golden vectors or numerical tests of its own for this path.
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

PATCH = 14
NUM_PREFIX = 5  # cls + 4 register tokens
LN_EPS = 1e-6
BN_EPS = 1e-5
BN_MOMENTUM = 0.1
LORA_RANK = 8
LORA_ALPHA = 1.0
CONVSTREAM_CH = [48, 96, 192]
FUSION_CH = [256, 128, 64, 32]


class Config:
    """Model geometry. The real model is Config() (ViT-g/14: 1536 wide, 40 deep, 24 heads, SwiGLU hidden 4096)."""

    def __init__(self, img_size=256, embed_dim=1536, depth=40, num_heads=24, hidden=4096, out_chans=16):
        self.img_size = img_size
        self.embed_dim = embed_dim
        self.depth = depth
        self.num_heads = num_heads
        self.hidden = hidden  # SwiGLU hidden width; fc1 has 2*hidden outputs, fc2 has hidden inputs
        self.out_chans = out_chans
        assert embed_dim % num_heads == 0
        assert img_size % 16 == 0

    @property
    def grid(self):
        return self.img_size // PATCH

    @property
    def tokens(self):
        return self.grid * self.grid + NUM_PREFIX

    def as_dict(self):
        return dict(img_size=self.img_size, embed_dim=self.embed_dim, depth=self.depth, num_heads=self.num_heads,
                    hidden=self.hidden, out_chans=self.out_chans)


# --------------------------------------------------------------------------------------------- initialisation
def _trunc_normal(shape, std, gen):
    t = torch.empty(shape, dtype=torch.float32)
    # timm trunc_normal_(std=.02) truncates at +-2 (absolute), i.e. 100 sigma: effectively a plain normal
    t.normal_(0.0, std, generator=gen)
    return t.clamp_(-2.0, 2.0)


def init_state_dict(cfg, seed=0, perturb=True):
    """Deterministic random-init weights under the reference's state-dict keys (fp32, CPU)."""
    g = torch.Generator().manual_seed(seed)
    D, H = cfg.embed_dim, cfg.hidden
    sd = OrderedDict()
    v = "encoder.vit."
    sd[v + "cls_token"] = torch.empty(1, 1, D).normal_(0, 1e-6, generator=g)
    sd[v + "reg_token"] = torch.empty(1, 4, D).normal_(0, 1e-6, generator=g)
    sd[v + "pos_embed"] = _trunc_normal((1, cfg.grid * cfg.grid, D), 0.02, g)
    fan_in = 3 * PATCH * PATCH
    bound = 1.0 / math.sqrt(fan_in)  # nn.Conv2d default init (timm leaves PatchEmbed.proj at the torch default)
    sd[v + "patch_embed.proj.weight"] = (torch.rand((D, 3, PATCH, PATCH), generator=g) * 2 - 1) * bound
    sd[v + "patch_embed.proj.bias"] = (torch.rand((D,), generator=g) * 2 - 1) * bound
    for i in range(cfg.depth):
        b = v + "blocks.%d." % i
        sd[b + "norm1.weight"] = torch.ones(D)
        sd[b + "norm1.bias"] = torch.zeros(D)
        sd[b + "attn.qkv.qkv.weight"] = _trunc_normal((3 * D, D), 0.02, g)
        sd[b + "attn.qkv.qkv.bias"] = torch.zeros(3 * D)
        for nm in ("lora_q", "lora_v"):
            sd[b + "attn.qkv.%s.A" % nm] = torch.randn((D, LORA_RANK), generator=g) / math.sqrt(LORA_RANK)
            sd[b + "attn.qkv.%s.B" % nm] = torch.zeros(LORA_RANK, D)
        sd[b + "attn.proj.weight"] = _trunc_normal((D, D), 0.02, g)
        sd[b + "attn.proj.bias"] = torch.zeros(D)
        sd[b + "ls1.gamma"] = torch.full((D,), 1e-5)
        sd[b + "norm2.weight"] = torch.ones(D)
        sd[b + "norm2.bias"] = torch.zeros(D)
        sd[b + "mlp.fc1.weight"] = _trunc_normal((2 * H, D), 0.02, g)
        sd[b + "mlp.fc1.bias"] = torch.zeros(2 * H)
        sd[b + "mlp.fc2.weight"] = _trunc_normal((D, H), 0.02, g)
        sd[b + "mlp.fc2.bias"] = torch.zeros(D)
        sd[b + "ls2.gamma"] = torch.full((D,), 1e-5)
    sd[v + "norm.weight"] = torch.ones(D)
    sd[v + "norm.bias"] = torch.zeros(D)

    def conv_bn(prefix_conv, prefix_bn, cin, cout):
        sd[prefix_conv + "weight"] = torch.empty(cout, cin, 3, 3).normal_(0, 0.02, generator=g)
        bn(prefix_bn, cout)

    def bn(prefix_bn, c):
        sd[prefix_bn + "weight"] = torch.empty(c).normal_(1.0, 0.02, generator=g)
        sd[prefix_bn + "bias"] = torch.zeros(c)
        sd[prefix_bn + "running_mean"] = torch.zeros(c)
        sd[prefix_bn + "running_var"] = torch.ones(c)
        sd[prefix_bn + "num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    d = "decoder."
    chans = [3] + CONVSTREAM_CH
    for i in range(3):
        conv_bn(d + "convstream.convs.%d.conv." % i, d + "convstream.convs.%d.bn." % i, chans[i], chans[i + 1])
    fus = [D] + FUSION_CH
    for i in range(4):
        cin = fus[i] + chans[-(i + 1)]
        conv_bn(d + "fusion_blks.%d.conv.conv." % i, d + "fusion_blks.%d.conv.bn." % i, cin, fus[i + 1])
    for h in range(cfg.out_chans):
        p = d + "segmentation_head_%d." % h
        sd[p + "0.psi.0.weight"] = torch.empty(16, 32, 1, 1).normal_(0, 0.02, generator=g)
        sd[p + "0.psi.0.bias"] = torch.zeros(16)
        bn(p + "0.psi.1.", 16)
        sd[p + "0.psi.3.weight"] = torch.empty(1, 16, 1, 1).normal_(0, 0.02, generator=g)
        sd[p + "0.psi.3.bias"] = torch.zeros(1)
        sd[p + "1.weight"] = torch.empty(1, 32, 3, 3).normal_(0, 0.02, generator=g)
        sd[p + "1.bias"] = torch.zeros(1)

    if perturb:
        # random-init parity is blind unless LayerScale, LoRA B, biases and BN statistics are made non-trivial
        for k in list(sd.keys()):
            t = sd[k]
            if k.endswith("gamma"):
                sd[k] = torch.rand(t.shape, generator=g) * 0.45 + 0.05
            elif k.endswith(("lora_q.B", "lora_v.B")):
                sd[k] = torch.empty(t.shape).normal_(0, 0.02, generator=g)
            elif k.endswith("running_mean"):
                sd[k] = torch.empty(t.shape).normal_(0, 0.1, generator=g)
            elif k.endswith("running_var"):
                sd[k] = torch.rand(t.shape, generator=g) + 0.5
            elif k.endswith(".bias") and t.dim() == 1:
                sd[k] = torch.empty(t.shape).normal_(0, 0.02, generator=g)
            elif k.endswith(("cls_token", "reg_token")):
                sd[k] = torch.empty(t.shape).normal_(0, 0.02, generator=g)
            elif k.endswith(("norm1.weight", "norm2.weight", "norm.weight")):
                sd[k] = 1.0 + torch.empty(t.shape).normal_(0, 0.05, generator=g)
    return sd


def trainable_keys(sd):
    """LoRA A/B of every block + every decoder parameter (apply_lora freezes the rest: lora.py:66-83)."""
    out = []
    for k, t in sd.items():
        if ".lora_" in k:
            out.append(k)
        elif k.startswith("decoder.") and not k.endswith(("running_mean", "running_var", "num_batches_tracked")):
            out.append(k)
    return out


# --------------------------------------------------------------------------------------------- synthetic data
HE_MEAN = (0.707223, 0.578729, 0.703617)  # src/dataset.py:601
HE_STD = (0.211883, 0.230117, 0.177517)
RGB_MEAN = (211.1, 194.7, 213.8)  # channel_stats.json RGB statistics
RGB_STD = (30.1, 36.4, 26.4)


def synthetic_tiles_u8(batch, size, seed=1234):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((batch, 3, size, size), generator=g)
    x = x * torch.tensor(RGB_STD).view(1, 3, 1, 1) + torch.tensor(RGB_MEAN).view(1, 3, 1, 1)
    return x.clamp_(0, 255).round_().to(torch.uint8)


def normalize_tiles(u8):
    """H-Optimus input normalisation, src/dataset.py:596-606 (mean/std scaled by 255)."""
    mean = torch.tensor(HE_MEAN).view(1, 3, 1, 1) * 255.0
    std = torch.tensor(HE_STD).view(1, 3, 1, 1) * 255.0
    return (u8.float() - mean) / std


def synthetic_targets(batch, chans, size, seed=4321):
    """uint8 'mostly dark' targets mapped to [-0.9, 0.9] as src/dataset.py:573 does."""
    g = torch.Generator().manual_seed(seed)
    u = torch.empty((batch, chans, size, size)).exponential_(1.0 / 20.0, generator=g).clamp_(0, 255).floor_()
    return u / 255.0 * 1.8 - 0.9


# --------------------------------------------------------------------------------------------- encoder
def lora_qkv(x, sd, b):
    """QkvWithLoRA.forward, lora.py:29-33: qkv = W x + b; q += alpha (x A_q) B_q; v += alpha (x A_v) B_v."""
    D = x.shape[-1]
    qkv = F.linear(x, sd[b + "attn.qkv.qkv.weight"], sd[b + "attn.qkv.qkv.bias"])
    dq = LORA_ALPHA * (x @ sd[b + "attn.qkv.lora_q.A"] @ sd[b + "attn.qkv.lora_q.B"])
    dv = LORA_ALPHA * (x @ sd[b + "attn.qkv.lora_v.A"] @ sd[b + "attn.qkv.lora_v.B"])
    return torch.cat([qkv[..., :D] + dq, qkv[..., D:2 * D], qkv[..., 2 * D:] + dv], dim=-1)


def vit_forward(sd, x, cfg, collect=None):
    """x [B,3,S,S] fp32 -> tokens [B, 5+g^2, D] after the final LayerNorm."""
    v = "encoder.vit."
    D, nh = cfg.embed_dim, cfg.num_heads
    hd = D // nh
    B = x.shape[0]
    t = F.conv2d(x, sd[v + "patch_embed.proj.weight"], sd[v + "patch_embed.proj.bias"], stride=PATCH)
    t = t.flatten(2).transpose(1, 2)  # [B, g*g, D]
    t = t + sd[v + "pos_embed"]
    t = torch.cat([sd[v + "cls_token"].expand(B, -1, -1), sd[v + "reg_token"].expand(B, -1, -1), t], dim=1)
    N = t.shape[1]
    if collect is not None:
        collect["tokens0"] = t
    for i in range(cfg.depth):
        b = v + "blocks.%d." % i
        h = F.layer_norm(t, (D,), sd[b + "norm1.weight"], sd[b + "norm1.bias"], LN_EPS)
        qkv = lora_qkv(h, sd, b).reshape(B, N, 3, nh, hd).permute(2, 0, 3, 1, 4)
        q, k, val = qkv[0], qkv[1], qkv[2]
        att = (q @ k.transpose(-2, -1)) * (hd ** -0.5)
        att = att.softmax(dim=-1)
        o = (att @ val).transpose(1, 2).reshape(B, N, D)
        o = F.linear(o, sd[b + "attn.proj.weight"], sd[b + "attn.proj.bias"])
        t = t + sd[b + "ls1.gamma"] * o
        h = F.layer_norm(t, (D,), sd[b + "norm2.weight"], sd[b + "norm2.bias"], LN_EPS)
        h = F.linear(h, sd[b + "mlp.fc1.weight"], sd[b + "mlp.fc1.bias"])
        x1, x2 = h.chunk(2, dim=-1)
        h = F.silu(x1) * x2
        h = F.linear(h, sd[b + "mlp.fc2.weight"], sd[b + "mlp.fc2.bias"])
        t = t + sd[b + "ls2.gamma"] * h
        if collect is not None:
            collect["block%d" % i] = t
    return F.layer_norm(t, (D,), sd[v + "norm.weight"], sd[v + "norm.bias"], LN_EPS)


def encoder_forward(sd, x, cfg, collect=None):
    """tokens -> [B, D, S/16, S/16] feature map (drop prefix tokens, channel-last view, bicubic g -> S/16)."""
    tok = vit_forward(sd, x, cfg, collect)
    B = x.shape[0]
    g = cfg.grid
    f = tok[:, NUM_PREFIX:].permute(0, 2, 1).reshape(B, cfg.embed_dim, g, g)
    tgt = cfg.img_size / 16
    sf = (tgt / g, tgt / g)
    f = F.interpolate(f, scale_factor=sf, mode="bicubic")
    if collect is not None:
        collect["features"] = f
    return f


# --------------------------------------------------------------------------------------------- decoder
def _bn(x, sd, p, training):
    if training:
        sd[p + "num_batches_tracked"] += 1
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                        training=training, momentum=BN_MOMENTUM, eps=BN_EPS)


def _conv_bn_relu(x, sd, pc, pb, stride, training):
    return F.relu(_bn(F.conv2d(x, sd[pc + "weight"], None, stride=stride, padding=1), sd, pb, training))


def decoder_forward(sd, feats, images, training=False, collect=None):
    d = "decoder."
    details = [images]
    x = images
    for i in range(3):
        x = _conv_bn_relu(x, sd, d + "convstream.convs.%d.conv." % i, d + "convstream.convs.%d.bn." % i, 2, training)
        details.append(x)
    f = feats
    for i in range(4):
        up = F.interpolate(f, scale_factor=2, mode="bilinear", align_corners=False)
        f = torch.cat([details[3 - i], up], dim=1)
        f = _conv_bn_relu(f, sd, d + "fusion_blks.%d.conv.conv." % i, d + "fusion_blks.%d.conv.bn." % i, 1, training)
        if collect is not None:
            collect["fusion%d" % i] = f
    outs = []
    h = 0
    while (d + "segmentation_head_%d.1.weight" % h) in sd:
        p = d + "segmentation_head_%d." % h
        a = F.conv2d(f, sd[p + "0.psi.0.weight"], sd[p + "0.psi.0.bias"])
        a = F.relu(_bn(a, sd, p + "0.psi.1.", training))
        a = torch.sigmoid(F.conv2d(a, sd[p + "0.psi.3.weight"], sd[p + "0.psi.3.bias"]))
        y = F.conv2d(f * a, sd[p + "1.weight"], sd[p + "1.bias"], padding=1)
        outs.append(torch.tanh(y))
        h += 1
    return torch.cat(outs, dim=1)


def miphei_forward(sd, x, cfg, training=False, collect=None):
    """ViTMatte.forward, mipheivit.py:106-110: x [B,3,S,S] -> [B,C,S,S] in (-1,1)."""
    return decoder_forward(sd, encoder_forward(sd, x, cfg, collect), x, training, collect)


# --------------------------------------------------------------------------------------------- losses
def weighted_mse_loss(y_true, y_pred, marker_weights, lambda_factor=50.0):
    loss = (y_pred - y_true) ** 2
    loss = loss.mean(dim=(0, 2, 3)) * marker_weights
    return loss.mean() * lambda_factor


def mae_loss(y_true, y_pred, lambda_factor=1.0):
    return (y_pred - y_true).abs().mean() * lambda_factor


def mse_loss(y_true, y_pred, lambda_factor=1.0):
    return ((y_pred - y_true) ** 2).mean() * lambda_factor


def l1_l2_loss(y_pred, y_true, lambda_factor=1.0):
    return lambda_factor * ((y_pred - y_true).abs().mean() + ((y_pred - y_true) ** 2).mean()) / 2


# --------------------------------------------------------------------------------------------- optimiser
def lr_lambda(step, total_steps, warmup_steps=400):
    """pix2pix_lr_scheduler's LambdaLR factor (utils.py:217-230): linear warm-up, flat to half, linear to zero."""
    half = total_steps // 2
    if step < warmup_steps:
        return float(step) / float(max(1, warmup_steps))
    if step < half:
        return 1.0
    return max(0.0, float(total_steps - step) / float(max(1, total_steps - half)))


def clip_grad_norm(grads, max_norm=1.0):
    """torch.nn.utils.clip_grad_norm_ (L2, error_if_nonfinite False) as Lightning's clip_gradients calls it."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return [g * coef for g in grads], total


def adam_update(p, g, m, v, step, lr, beta1=0.5, beta2=0.999, eps=1e-7):
    """torch.optim.Adam single-tensor update (no weight decay, no amsgrad); step counts from 1."""
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


def train_step(sd, opt_state, x, y, cfg, marker_weights, base_lr, total_steps, lambda_factor=50.0, warmup_steps=400):
    """One optimisation step on the trainable subset; returns (loss, grad dict, grad norm). Mutates sd / opt_state."""
    keys = trainable_keys(sd)
    params = []
    for k in keys:
        sd[k] = sd[k].detach().requires_grad_(True)
        params.append(sd[k])
    pred = miphei_forward(sd, x, cfg, training=True)
    loss = weighted_mse_loss(y, pred, marker_weights, lambda_factor)
    grads = torch.autograd.grad(loss, params)
    raw = {k: g.detach().clone() for k, g in zip(keys, grads)}
    clipped, gnorm = clip_grad_norm([g.detach() for g in grads], 1.0)
    step = opt_state.setdefault("step", 0) + 1
    opt_state["step"] = step
    lr = base_lr * lr_lambda(step - 1, total_steps, warmup_steps)
    with torch.no_grad():
        for k, g in zip(keys, clipped):
            p = sd[k].detach()
            m = opt_state.setdefault("m." + k, torch.zeros_like(p))
            v = opt_state.setdefault("v." + k, torch.zeros_like(p))
            adam_update(p, g, m, v, step, lr)
            sd[k] = p
    return loss.detach(), raw, gnorm, pred.detach()


# --------------------------------------------------------------------------------------------- parity metrics
def pearson(a, b):
    a = a.double().flatten() - a.double().mean()
    b = b.double().flatten() - b.double().mean()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def cosine(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def per_channel_max_rel_err(got, ref):
    """max |got-ref| per output channel divided by that channel's max |ref| (north-star parity metric)."""
    C = ref.shape[1]
    g = got.double().transpose(0, 1).reshape(C, -1)
    r = ref.double().transpose(0, 1).reshape(C, -1)
    return ((g - r).abs().amax(dim=1) / (r.abs().amax(dim=1) + 1e-12)).tolist()
