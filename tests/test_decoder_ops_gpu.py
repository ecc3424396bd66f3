"""GPU parity of the decoder kernels (implicit-GEMM 3x3 conv, fused heads, resize glue) against torch fp32 ops."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _mods():
    from miphei_vit_b200 import ops, packing
    return ops, packing


def _rand(shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(shape, generator=g, device="cuda") * scale


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous().bfloat16()


def _rel(got, ref):
    return (got.float() - ref).abs().max().item() / (ref.abs().max().item() + 1e-12)


@pytest.mark.parametrize("B,H,C0,C1,Cout,stride", [
    (2, 32, 192, 1536, 256, 1), (2, 64, 96, 256, 128, 1), (1, 128, 48, 128, 64, 1), (1, 256, 8, 64, 32, 1),
    (2, 256, 8, 0, 48, 2), (2, 128, 48, 0, 96, 2), (2, 64, 96, 0, 192, 2), (3, 16, 192, 128, 256, 1),
])
def test_conv3x3_implicit_gemm(B, H, C0, C1, Cout, stride):
    ops, packing = _mods()
    real0 = 3 if C0 == 8 else C0
    x0 = _rand((B, C0, H, H), 1.0, 1)
    if C0 == 8:
        x0[:, 3:] = 0
    x1 = _rand((B, C1, H, H), 1.0, 2) if C1 else None
    w = _rand((Cout, real0 + C1, 3, 3), 0.05, 3)
    scale = 1 + _rand((Cout,), 0.1, 4)
    shift = _rand((Cout,), 0.1, 5)
    wp = packing.pack_conv3x3(w, [real0, C1] if C1 else [real0])
    out = ops.gemm(_nhwc(x0), wp, conv=dict(stride=stride, a2=_nhwc(x1) if C1 else None), scale=scale, shift=shift,
                   act=ops.ACT_RELU)
    xin = torch.cat([x0[:, :real0], x1], 1) if C1 else x0[:, :real0]
    ref = F.relu(F.conv2d(xin.bfloat16().float(), w.bfloat16().float(), stride=stride, padding=1) *
                 scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    Ho = H // stride
    got = out.view(B, Ho, Ho, Cout).permute(0, 3, 1, 2)
    assert _rel(got, ref) < 1.5e-2, _rel(got, ref)


@pytest.mark.parametrize("B,S,heads", [(2, 128, 16), (1, 256, 16), (2, 128, 3)])
def test_fused_heads(B, S, heads):
    ops, packing = _mods()
    C = 32
    f = F.relu(_rand((B, C, S, S), 1.0, 1))
    hps, refs = [], []
    fb = f.bfloat16().float()
    for h in range(heads):
        hp = dict(psi0_w=_rand((16, C, 1, 1), 0.2, 10 + h), psi0_b=_rand((16,), 0.1, 30 + h),
                  bn=(1 + _rand((16,), 0.1, 50 + h), _rand((16,), 0.1, 70 + h), _rand((16,), 0.1, 90 + h),
                      0.5 + torch.rand(16, device="cuda")),
                  psi3_w=_rand((1, 16, 1, 1), 0.5, 110 + h), psi3_b=_rand((1,), 0.1, 130 + h),
                  conv_w=_rand((1, C, 3, 3), 0.1, 150 + h), conv_b=_rand((1,), 0.1, 170 + h))
        hps.append(hp)
        a = F.conv2d(fb, hp["psi0_w"].bfloat16().float(), hp["psi0_b"])
        a = F.relu(F.batch_norm(a, hp["bn"][2], hp["bn"][3], hp["bn"][0], hp["bn"][1], False, 0.1, 1e-5))
        g = torch.sigmoid(F.conv2d(a, hp["psi3_w"], hp["psi3_b"]))
        refs.append(torch.tanh(F.conv2d(fb * g, hp["conv_w"].bfloat16().float(), hp["conv_b"], padding=1)))
    ref = torch.cat(refs, 1)
    pk = packing.pack_heads(hps)
    fn = _nhwc(f)
    gate = ops.gemm(fn.view(-1, C), pk["gate_w"][:, :C], mode=ops.GEMM_HEAD_GATE, scale=pk["gate_scale"],
                    shift=pk["gate_shift"], in2=pk["gate_w2"], resid=pk["gate_b2"])
    assert gate.shape == (B * S * S, heads)
    out = ops.gemm(fn, pk["conv_w"], mode=ops.GEMM_HEAD_CONV, conv=dict(stride=1), shift=pk["conv_b"], in2=gate,
                   out_dtype=torch.float32)
    assert out.shape == (B, heads, S, S)
    assert (out - ref).abs().max().item() < 2e-2, (out - ref).abs().max().item()
    out8 = ops.gemm(fn, pk["conv_w"], mode=ops.GEMM_HEAD_CONV, conv=dict(stride=1), shift=pk["conv_b"], in2=gate,
                    out_dtype=torch.uint8)
    ref8 = (((ref + 0.9) / 1.8).clamp(0, 1) * 255).to(torch.uint8)
    assert (out8.int() - ref8.int()).abs().max().item() <= 3


def test_prep_input_and_patch_embed():
    ops, _ = _mods()
    B, S, D = 2, 256, 256
    x = _rand((B, 3, S, S), 1.0, 1)
    img, pm = ops.prep_input(x)
    assert torch.equal(img[..., :3], x.permute(0, 2, 3, 1).bfloat16()) and (img[..., 3:] == 0).all()
    w = _rand((D, 3, 14, 14), 0.05, 2)
    ref = F.conv2d(x.bfloat16().float(), w.bfloat16().float(), stride=14).flatten(2).transpose(1, 2).reshape(-1, D)
    wp = torch.zeros((D, 592), device="cuda")
    wp[:, :588] = w.flatten(1)
    got = ops.gemm(pm, wp.bfloat16(), out_dtype=torch.float32)
    assert _rel(got, ref) < 1e-3


@pytest.mark.parametrize("B,g,t,D", [(2, 18, 16, 1536), (1, 9, 8, 128), (1, 36, 32, 256)])
def test_tokens_to_map_bicubic(B, g, t, D):
    ops, _ = _mods()
    N = g * g + 5
    tok = _rand((B * N, D), 1.0, 1).bfloat16()
    got = ops.tokens_to_map(tok, B, N, 5, g, t)
    f = tok.float().view(B, N, D)[:, 5:].permute(0, 2, 1).reshape(B, D, g, g)
    ref = F.interpolate(f, scale_factor=(t / g, t / g), mode="bicubic")
    assert ref.shape[-1] == t
    assert (got.float().permute(0, 3, 1, 2) - ref).abs().max().item() < 3e-2


def test_upsample2x_bilinear():
    ops, _ = _mods()
    x = _rand((2, 64, 16, 16), 1.0, 1)
    got = ops.upsample2x(_nhwc(x))
    ref = F.interpolate(x.bfloat16().float(), scale_factor=2, mode="bilinear", align_corners=False)
    assert (got.float().permute(0, 3, 1, 2) - ref).abs().max().item() < 2e-2


def test_fill_prefix():
    ops, _ = _mods()
    B, N, D = 3, 86, 128
    x = torch.zeros((B * N, D), device="cuda")
    pre = _rand((5, D), 1.0, 1)
    ops.fill_prefix(x, pre, B, N)
    xv = x.view(B, N, D)
    assert torch.equal(xv[:, :5], pre.expand(B, 5, D)) and (xv[:, 5:] == 0).all()
