"""GPU parity of the backward / optimiser kernels against torch autograd and torch.optim on the same inputs."""
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


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


@pytest.mark.parametrize("B,N,H", [(2, 329, 3), (1, 86, 2), (2, 128, 1), (1, 384, 2), (3, 200, 2), (4, 329, 24), (3, 16, 2), (2, 48, 1),
                                   (8, 64, 24), (1, 1301, 2), (1, 700, 3), (40, 33, 24)])
def test_attention_backward(B, N, H):
    ops = _ops()
    D = H * 64
    qkv = _rand((B * N, 3 * D), 1.0, 7).bfloat16()
    out, lse = ops.attn_fwd(qkv, B, N, H, want_lse=True)
    dout = _rand((B * N, D), 1.0, 8).bfloat16()
    dqkv = ops.attn_bwd(qkv, out, dout, lse, B, N, H)
    x = qkv.float().requires_grad_(True)
    q, k, v = x.view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B * N, D)
    o.backward(dout.float())
    ref = x.grad
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        c = _cos(dqkv[:, sl], ref[:, sl])
        err = (dqkv[:, sl].float() - ref[:, sl]).abs().max().item() / (ref[:, sl].abs().max().item() + 1e-9)
        assert c > 0.9995 and err < 0.03, (name, c, err)


@pytest.mark.parametrize("M,N,K", [(16, 1536, 5264), (16, 4608, 10528), (128, 256, 1000), (40, 64, 333 * 8)])
def test_gemm_nn_atomic(M, N, K):
    ops = _ops()
    a = _rand((M, K), 1.0, 1).bfloat16()
    b = _rand((K, N), 0.1, 2).bfloat16()
    out = torch.zeros((M, N), dtype=torch.float32, device="cuda")
    ops.gemm(a, b, mode=ops.GEMM_NN_ATOMIC, out=out)
    ref = a.float() @ b.float()
    assert (out - ref).abs().max().item() < 2e-3 * ref.abs().max().item()


@pytest.mark.parametrize("M,D", [(1000, 256), (10528, 1536), (37, 128), (5000, 512), (2 * 329, 1024), (700, 320)])
def test_lora_grads(M, D):
    """mma.sync path for the widths it is instantiated for (D / 64 in {2, 4, 6, 8, 12, 16, 24}), split-K tcgen05 GEMMs otherwise"""
    ops = _ops()
    xn_ext = _rand((M, D + 64), 1.0, 1).bfloat16()
    dq_ext = _rand((M, 3 * D + 64), 1.0, 2).bfloat16()
    dAq, dAv = torch.zeros(D, 8, device="cuda"), torch.zeros(D, 8, device="cuda")
    dBq, dBv = torch.zeros(8, D, device="cuda"), torch.zeros(8, D, device="cuda")
    ops.lora_grads(xn_ext, dq_ext, D, 0.5, dAq, dAv, dBq, dBv)
    xn, T = xn_ext[:, :D].float(), xn_ext[:, D:D + 16].float()
    dqkv, dT = dq_ext[:, :3 * D].float(), dq_ext[:, 3 * D:3 * D + 16].float()
    for got, ref in ((dAq, xn.t() @ dT[:, :8]), (dAv, xn.t() @ dT[:, 8:]), (dBq, 0.5 * T[:, :8].t() @ dqkv[:, :D]),
                     (dBv, 0.5 * T[:, 8:].t() @ dqkv[:, 2 * D:])):
        assert (got - ref).abs().max().item() < 2e-3 * ref.abs().max().item()


@pytest.mark.parametrize("B,g,t,D", [(2, 18, 16, 1536), (1, 9, 8, 128)])
def test_tokens_to_map_backward(B, g, t, D):
    ops = _ops()
    N = g * g + 5
    dmap = _rand((B, t, t, D), 1.0, 3).bfloat16()
    got = ops.tokens_to_map_bwd(dmap, B, N, 5, g)
    f = torch.zeros((B, D, g, g), device="cuda", requires_grad=True)
    F.interpolate(f, scale_factor=(t / g, t / g), mode="bicubic").backward(dmap.float().permute(0, 3, 1, 2))
    ref = f.grad.permute(0, 2, 3, 1).reshape(B, g * g, D)
    gv = got.float().view(B, N, D)
    assert (gv[:, :5] == 0).all()
    assert (gv[:, 5:] - ref).abs().max().item() < 2e-2 * ref.abs().max().item()


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_loss_forward_backward(mode):
    ops = _ops()
    B, C, S = 3, 16, 64
    pred = torch.tanh(_rand((B, C, S, S), 1.0, 1))
    tgt = _rand((B, C, S, S), 0.5, 2).clamp(-0.9, 0.9)
    w = torch.linspace(1, 10, C, device="cuda")
    p = pred.clone().requires_grad_(True)
    if mode == 0:
        ref = (((p - tgt) ** 2).mean(dim=(0, 2, 3)) * w).mean() * 50.0
        lam = 50.0
    elif mode == 1:
        ref = F.l1_loss(p, tgt) * 3.0
        lam = 3.0
    else:
        ref = 3.0 * (F.l1_loss(p, tgt) + F.mse_loss(p, tgt)) / 2
        lam = 3.0
    ref.backward()
    loss, grad = ops.loss_fwd_bwd(pred, tgt, w if mode == 0 else None, mode=mode, lambda_factor=lam)
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item()) + 1e-7
    assert (grad - p.grad).abs().max().item() < 1e-6 * p.grad.abs().max().item() + 1e-12


def test_clip_and_adam_match_torch():
    ops = _ops()
    n = 100003
    p0 = _rand((n,), 1.0, 1)
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p_ref], lr=8e-4, betas=(0.5, 0.999), eps=1e-7)
    p = p0.clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        g = _rand((n,), 0.01 * step, 10 + step)
        p_ref.grad = g.clone()
        ref_norm = torch.nn.utils.clip_grad_norm_([p_ref], 1.0)
        opt.step()
        nc = ops.grad_norm(g, 1.0)
        assert abs(nc[0].item() - ref_norm.item()) < 1e-4 * ref_norm.item()
        ops.adam_clip_step(p, g, m, v, nc, step, 8e-4)
        assert (p - p_ref.detach()).abs().max().item() < 2e-6


@pytest.mark.parametrize("M", [64, 1000, 65536 + 37])
def test_gram32_and_closed_form_head_batchnorm(M):
    """mv_gram32 / mv_heads_bn_from_gram against torch: moments of f, then train-mode BatchNorm2d statistics of the
    1x1-conv gate units W1 f + b1 (AttentionBlock.psi[0..1], src/generators/unet.py:407-422)."""
    ops = _ops()
    f = (_rand((M, 32), 1.0, 3).relu() + 0.1).bfloat16()  # post-ReLU map: non-negative, non-zero mean
    gram = ops.gram32(f)
    fd = f.double()
    ref_g = fd.t() @ fd
    assert torch.allclose(gram[:32].double(), ref_g, rtol=2e-4, atol=1e-3 * M ** 0.5)
    assert torch.allclose(gram[32].double(), fd.sum(0), rtol=2e-4, atol=1e-3)
    assert float(gram[33:].abs().max()) == 0.0
    C = 48
    w1 = _rand((C, 32), 0.05, 4).bfloat16().float().contiguous()
    b1, gamma, beta = _rand((C,), 0.1, 5), _rand((C,), 0.2, 6) + 1.0, _rand((C,), 0.1, 7)
    bn = torch.nn.BatchNorm2d(C).cuda().train()
    with torch.no_grad():
        bn.weight.copy_(gamma)
        bn.bias.copy_(beta)
        bn.running_mean.normal_(0, 0.1)
        bn.running_var.uniform_(0.5, 1.5)
    rm, rv = bn.running_mean.clone(), bn.running_var.clone()
    a = (f.float() @ w1.t() + b1).t().reshape(1, C, M, 1)
    y_ref = bn(a)
    fin = ops.heads_bn_from_gram(gram, M, w1, b1, gamma, beta, rm, rv, momentum=bn.momentum, eps=bn.eps)
    y = (f.float() @ w1.t()) * fin[0] + fin[1]  # folded (scale, shift) act on W1 f
    assert torch.allclose(y.t().reshape(1, C, M, 1), y_ref, rtol=1e-3, atol=2e-3)
    assert torch.allclose(fin[2], a.mean((0, 2, 3)), rtol=1e-4, atol=1e-5)
    assert torch.allclose(fin[3], (a.var((0, 2, 3), unbiased=False) + bn.eps).rsqrt(), rtol=1e-3)
    assert torch.allclose(rm, bn.running_mean, rtol=1e-4, atol=1e-5)
    assert torch.allclose(rv, bn.running_var, rtol=1e-3, atol=1e-5)
