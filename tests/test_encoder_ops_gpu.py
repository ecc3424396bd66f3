"""GPU parity of LayerNorm and attention kernels against torch fp32 references of the same ops."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from miphei_vit_b200 import ops
    return ops


def _rand(shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(shape, generator=g, device="cuda") * scale


@pytest.mark.parametrize("M,D", [(5264, 1536), (77, 128), (1000, 256), (329, 768), (10528, 1536), (2400, 1536), (9000, 256)])
def test_layernorm_fwd_bwd(M, D):
    ops = _ops()
    x = _rand((M, D), 2.0, 1) + 0.5
    w = 1 + _rand((D,), 0.1, 2)
    b = _rand((D,), 0.1, 3)
    y, mean, rstd = ops.layernorm_fwd(x, w, b, stats=True)
    ref = F.layer_norm(x, (D,), w, b, 1e-6)
    assert (y.float() - ref).abs().max().item() < 0.03
    assert (mean - x.mean(1)).abs().max().item() < 1e-5
    # strided output (LoRA K-extension buffer)
    buf = torch.zeros((M, D + 64), dtype=torch.bfloat16, device="cuda")
    ops.layernorm_fwd(x, w, b, out=buf[:, :D])
    assert torch.equal(buf[:, :D], y) and (buf[:, D:] == 0).all()
    # backward
    xr = x.clone().requires_grad_(True)
    dy = _rand((M, D), 1.0, 4)
    dres = _rand((M, D), 1.0, 5)
    F.layer_norm(xr, (D,), w, b, 1e-6).backward(dy)
    want = xr.grad + dres
    got, gotb = ops.layernorm_bwd(x, w, dy, dres=dres, want_bf16=True)
    assert (got - want).abs().max().item() < 1e-4 * want.abs().max().item() + 1e-5
    assert (gotb.float() - want).abs().max().item() < 0.02 * want.abs().max().item()
    got2 = ops.layernorm_bwd(x, w, dy.bfloat16())
    want2 = torch.autograd.grad(F.layer_norm(xr, (D,), w, b, 1e-6), xr, dy.bfloat16().float())[0]
    assert (got2 - want2).abs().max().item() < 1e-4 * want2.abs().max().item() + 1e-5


@pytest.mark.parametrize("B,N,H", [(2, 329, 24), (1, 86, 2), (3, 128, 4), (2, 16, 1), (1, 384, 3), (2, 368, 2), (5, 329, 5), (2, 1301, 3), (1, 700, 2), (3, 129, 2), (1, 5334, 1), (2, 200, 2), (16, 329, 24), (1, 32, 2), (1, 48, 1), (2, 64, 2), (1, 40, 1), (1, 100, 1)])
def test_attention_fwd(B, N, H):
    ops = _ops()
    D = H * 64
    qkv = _rand((B * N, 3 * D), 2.0, 7).bfloat16()
    out, lse = ops.attn_fwd(qkv, B, N, H, want_lse=True)
    q, k, v = qkv.float().view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * 0.125
    ref = (s.softmax(-1) @ v).transpose(1, 2).reshape(B * N, D)
    err = (out.float() - ref).abs().max().item()
    assert err < 0.03 * ref.abs().max().item(), err
    assert (lse - torch.logsumexp(s, -1)).abs().max().item() < 2e-2
