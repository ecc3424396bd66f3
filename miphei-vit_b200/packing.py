"""Weight re-layout for the kernels (host-side plumbing, done when weights change — not per forward).

All functions take the reference-layout fp32 parameters (PyTorch conv / linear layouts under the reference's
state-dict keys) and return contiguous bf16 / fp32 tensors in the layouts include/miphei_b200.h documents.
"""
import torch


def _pad64(c):
    return (c + 63) // 64 * 64


def decoder_layout(model):
    """[(name, parameter, offset, numel)] of the trainable decoder parameters inside a flat fp32 buffer — module order,
    every segment 16-byte aligned — and the padded total.  Shared by trainer.FlatParams (the decoder segment comes first in
    its flat buffer) and decoder_train.DecoderTrain (gather tables), so both always agree on the offsets."""
    out, off = [], 0
    for n, p in model.named_parameters():
        if n.startswith("decoder.") and p.requires_grad:
            out.append((n, p, off, p.numel()))
            off += (p.numel() + 3) // 4 * 4
    return out, off


def pack_conv3x3(weight, splits, dtype=torch.bfloat16):
    """[Cout, sum(splits), 3, 3] -> bf16 [Cout, 9 * sum(pad64(s))]: per tap (ky, kx) each source's channels zero-padded
    to a multiple of 64, sources in concat order (mv_gemm_bf16 conv mode). `splits` may hold a narrower real channel count
    than the NHWC buffer carries (e.g. the 3 image channels stored in 8): only the real channels get weights."""
    cout = weight.shape[0]
    assert weight.shape[1] == sum(splits) and weight.shape[2:] == (3, 3)
    w = weight.detach().float().permute(0, 2, 3, 1).reshape(cout, 9, -1)  # [Cout, tap, Cin]
    parts = []
    off = 0
    for s in splits:
        blk = torch.zeros((cout, 9, _pad64(s)), dtype=torch.float32, device=weight.device)
        blk[:, :, :s] = w[:, :, off:off + s]
        parts.append(blk)
        off += s
    out = torch.cat(parts, dim=2).reshape(cout, -1)
    return (out if dtype is None else out.to(dtype)).contiguous()


def fold_bn_eval(bn_weight, bn_bias, running_mean, running_var, eps=1e-5, conv_bias=None):
    """BatchNorm2d in eval mode (and an optional preceding conv bias) as per-channel (scale, shift) fp32 vectors."""
    scale = bn_weight.detach().float() / torch.sqrt(running_var.detach().float() + eps)
    shift = bn_bias.detach().float() - running_mean.detach().float() * scale
    if conv_bias is not None:
        shift = shift + conv_bias.detach().float() * scale
    return scale.contiguous(), shift.contiguous()


def pack_heads(head_params):
    """head_params: list over heads of dict(psi0_w [16,C,1,1], psi0_b, bn=(w, b, rm, rv), psi3_w [1,16,1,1], psi3_b [1],
    conv_w [1,C,3,3], conv_b [1]) -> operands of MV_GEMM_HEAD_GATE / MV_GEMM_HEAD_CONV (eval-mode BatchNorm)."""
    w1, sc, sh, w2, b2, w3, b3 = [], [], [], [], [], [], []
    for hp in head_params:
        w1.append(hp["psi0_w"].detach().float().flatten(1))  # [16, C]
        s, t = fold_bn_eval(*hp["bn"], conv_bias=hp["psi0_b"])
        sc.append(s), sh.append(t)
        w2.append(hp["psi3_w"].detach().float().flatten())
        b2.append(hp["psi3_b"].detach().float().flatten())
        w3.append(hp["conv_w"].detach().float())
        b3.append(hp["conv_b"].detach().float().flatten())
    C = w1[0].shape[1]
    w1 = torch.cat(w1, 0)  # [16*heads, C]
    w1p = torch.zeros((w1.shape[0], 64), dtype=torch.float32, device=w1.device)
    w1p[:, :C] = w1
    w3 = torch.cat(w3, 0)  # [heads, C, 3, 3]
    return dict(
        gate_w=w1p.to(torch.bfloat16).contiguous(), gate_scale=torch.cat(sc).contiguous(),
        gate_shift=torch.cat(sh).contiguous(), gate_w2=torch.cat(w2).contiguous(), gate_b2=torch.cat(b2).contiguous(),
        conv_w=pack_conv3x3(w3, [C]), conv_b=torch.cat(b3).contiguous())
