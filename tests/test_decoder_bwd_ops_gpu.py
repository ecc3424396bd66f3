"""GPU parity of the decoder training kernels (BatchNorm batch statistics / backward, conv dgrad + wgrad on the implicit
GEMM, resize adjoints, head-gradient stencil) against torch autograd of the same ops."""
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


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def _rel(got, ref):
    return (got.float() - ref).abs().max().item() / (ref.abs().max().item() + 1e-12)


@pytest.mark.parametrize("zdt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,Cin,Cout", [(2, 32, 64, 48), (2, 16, 192, 256), (1, 64, 8, 96)])
def test_conv_bn_relu_train_forward_backward(B, H, Cin, Cout, zdt):
    ops, packing = _mods()
    x = _rand((B, Cin, H, H), 1.0, 1).bfloat16().float()
    w = _rand((Cout, Cin, 3, 3), 0.05, 2).bfloat16().float()
    gamma, beta = 1 + _rand((Cout,), 0.1, 3), _rand((Cout,), 0.1, 4)
    rm, rv = _rand((Cout,), 0.1, 5), 0.5 + torch.rand(Cout, device="cuda")
    rm_ref, rv_ref = rm.clone(), rv.clone()
    # reference
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    zr = F.conv2d(xr, wr, padding=1)
    yr = F.relu(F.batch_norm(zr, rm_ref, rv_ref, gr, br, True, 0.1, 1e-5))
    dy = _rand(yr.shape, 1.0, 6).bfloat16().float()
    yr.backward(dy)
    # kernels
    xn = _nhwc(x)
    M = B * H * H
    stats = torch.zeros((2, Cout), device="cuda")
    z = ops.gemm(xn, packing.pack_conv3x3(w, [Cin]), conv=dict(stride=1), colstats=stats, out_dtype=zdt)
    fin = ops.bn_finalize(stats, M, gamma, beta, rm, rv)
    y = ops.bn_relu_apply(z, fin[0], fin[1])
    assert _rel(y.view(B, H, H, Cout).permute(0, 3, 1, 2), yr.detach()) < 2e-2
    assert (rm - rm_ref).abs().max().item() < 1e-3 and (rv - rv_ref).abs().max().item() < 2e-3
    dyn = _nhwc(dy).view(M, Cout)
    dz, sums = ops.bn_relu_bwd(dyn, y, z, fin[2], fin[3], gamma)
    # ReLU-mask flips where the bf16-stored z sits at rounding distance from 0 move single terms of these sums
    assert _cos(sums[0], br.grad) > 0.999 and _cos(sums[1], gr.grad) > 0.999
    # weight gradient: NN GEMM over pixels with the conv input read as tap-shifted TMA tiles
    dzT = ops.transpose_bf16(dz)
    Kp = 9 * 64 * ((Cin + 63) // 64)
    dwp = torch.zeros((Cout, Kp), device="cuda")
    ops.gemm(dzT[:, :M], xn, mode=ops.GEMM_NN_ATOMIC, conv=dict(stride=1), out=dwp)
    dw = dwp.view(Cout, 9, -1)[:, :, :Cin].reshape(Cout, 3, 3, Cin).permute(0, 3, 1, 2)
    assert _cos(dw, wr.grad) > 0.9995, _cos(dw, wr.grad)
    # data gradient: conv of dz with the flipped / transposed weights
    wd = packing.pack_conv3x3(w.flip(2, 3).permute(1, 0, 2, 3).contiguous(), [Cout])
    dx = ops.gemm(dz.view(B, H, H, Cout), wd, conv=dict(stride=1))
    assert _cos(dx.view(B, H, H, Cin).permute(0, 3, 1, 2), xr.grad) > 0.9995


@pytest.mark.parametrize("B,H,Cin,Cout", [(2, 64, 48, 96), (1, 32, 96, 192)])
def test_stride2_conv_gradients(B, H, Cin, Cout):
    ops, packing = _mods()
    x = _rand((B, Cin, H, H), 1.0, 1).bfloat16().float().requires_grad_(True)
    w = _rand((Cout, Cin, 3, 3), 0.05, 2).bfloat16().float().requires_grad_(True)
    z = F.conv2d(x, w, stride=2, padding=1)
    dz = _rand(z.shape, 1.0, 3).bfloat16().float()
    z.backward(dz)
    Ho = H // 2
    dzn = _nhwc(dz)
    u = ops.zero_insert2x(dzn)
    wd = packing.pack_conv3x3(w.detach().flip(2, 3).permute(1, 0, 2, 3).contiguous(), [Cout])
    dx = ops.gemm(u, wd, conv=dict(stride=1))
    assert _rel(dx.view(B, H, H, Cin).permute(0, 3, 1, 2), x.grad) < 2e-2
    dzT = ops.transpose_bf16(dzn.view(-1, Cout))
    dwp = torch.zeros((Cout, 9 * 64 * ((Cin + 63) // 64)), device="cuda")
    ops.gemm(dzT[:, :B * Ho * Ho], _nhwc(x.detach()), mode=ops.GEMM_NN_ATOMIC, conv=dict(stride=2), out=dwp)
    dw = dwp.view(Cout, 9, -1)[:, :, :Cin].reshape(Cout, 3, 3, Cin).permute(0, 3, 1, 2)
    assert _rel(dw, w.grad) < 2e-2


def test_wgrad_two_sources():
    ops, packing = _mods()
    B, H, C0, C1, Cout = 2, 32, 96, 256, 128
    x0, x1 = _rand((B, C0, H, H), 1.0, 1).bfloat16().float(), _rand((B, C1, H, H), 1.0, 2).bfloat16().float()
    w = _rand((Cout, C0 + C1, 3, 3), 0.05, 3).requires_grad_(True)
    z = F.conv2d(torch.cat([x0, x1], 1), w, padding=1)
    dz = _rand(z.shape, 1.0, 4).bfloat16().float()
    z.backward(dz)
    dzT = ops.transpose_bf16(_nhwc(dz).view(-1, Cout))
    p0, p1 = (C0 + 63) // 64 * 64, (C1 + 63) // 64 * 64
    dwp = torch.zeros((Cout, 9 * (p0 + p1)), device="cuda")
    ops.gemm(dzT[:, :B * H * H], _nhwc(x0), mode=ops.GEMM_NN_ATOMIC, conv=dict(stride=1, a2=_nhwc(x1)), out=dwp)
    v = dwp.view(Cout, 9, p0 + p1)
    dw = torch.cat([v[:, :, :C0], v[:, :, p0:p0 + C1]], 2).reshape(Cout, 3, 3, C0 + C1).permute(0, 3, 1, 2)
    assert _rel(dw, w.grad) < 2e-2


def test_upsample2x_backward_and_add():
    ops, _ = _mods()
    B, h, C = 2, 16, 64
    x = torch.zeros((B, C, h, h), device="cuda", requires_grad=True)
    dup = _rand((B, C + 32, 2 * h, 2 * h), 1.0, 1).bfloat16().float()
    F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False).backward(dup[:, 32:])
    dn = _nhwc(dup)  # [B, 2h, 2w, 32 + C]; the gradient of the upsampled source is a channel slice
    got = ops.upsample2x_bwd(dn[..., 32:])
    assert _rel(got.permute(0, 3, 1, 2), x.grad) < 2e-2
    a, b = _rand((100, 48), 1.0, 2).bfloat16(), _rand((100, 112), 1.0, 3).bfloat16()
    s = ops.add_bf16(a, b[:, 64:])
    assert _rel(s, a.float() + b[:, 64:].float()) < 1e-2


def test_transpose_with_ones_row():
    ops, _ = _mods()
    x = _rand((1000, 32), 1.0, 1).bfloat16()
    t = ops.transpose_bf16(x, ones_row=True)
    assert t.shape[0] == 40 and torch.equal(t[:32, :1000], x.t()) and (t[32, :1000] == 1).all() and (t[33:] == 0).all()


@pytest.mark.parametrize("M,C", [(1000, 32), (4099, 48), (2048, 64), (777, 256), (65536, 32), (300, 8), (512, 30)])
def test_transpose_shapes_formats_and_strides(M, C):
    """register-transpose fast path (C % 8 == 0, aligned) and the generic kernel: row tails, strided input rows, fp16 input
    converted to bf16 on the way, the row of ones"""
    ops, _ = _mods()
    x = _rand((M, C), 1.0, 1)
    for dt in (torch.bfloat16, torch.float16):
        xb = x.to(dt)
        t = ops.transpose_bf16(xb)
        assert t.dtype == torch.bfloat16 and torch.equal(t[:C, :M], xb.to(torch.bfloat16).t())
        wide = _rand((M, C + 16), 1.0, 2).to(dt)          # strided rows (a channel slice of a wider map)
        t2 = ops.transpose_bf16(wide[:, 8:8 + C])
        assert torch.equal(t2[:C, :M], wide[:, 8:8 + C].to(torch.bfloat16).t())
        if C % 8 == 0:
            t3 = ops.transpose_bf16(xb, ones_row=True)
            assert torch.equal(t3[:C, :M], xb.to(torch.bfloat16).t()) and bool((t3[C, :M] == 1).all()) and bool((t3[C + 1:] == 0).all())


def test_gate_mask_epilogue():
    ops, _ = _mods()
    M, C, heads = 700, 32, 16
    f = _rand((M, C), 1.0, 1).bfloat16()
    w1 = _rand((16 * heads, C), 0.2, 2).bfloat16()
    scale, shift = 1 + _rand((256,), 0.1, 3), _rand((256,), 0.3, 4)
    du = _rand((M, 16), 1.0, 5).bfloat16()
    e = ops.gemm(f, w1, scale=scale, shift=shift, act=ops.ACT_GATE_MASK, in2=du)
    a = (f.float() @ w1.float().t()) * scale + shift
    ref = torch.where(a > 0, du.float().repeat_interleave(16, dim=1), torch.zeros_like(a))
    mism = ((e.float() - ref).abs() > 1e-6).float().mean().item()
    assert mism < 2e-3  # sign flips only where |a| is at rounding level


@pytest.mark.parametrize("B,S,heads", [(2, 64, 16), (1, 128, 3)])
def test_heads_gradient_stencil(B, S, heads):
    """dt, du, db2, db3 of  pred_h = tanh(b3_h + sum_tap g_h(p+o) t_{h,tap}(p+o)),  g = sigmoid(u)."""
    ops, _ = _mods()
    M = B * S * S
    t = (_rand((M, 144), 0.3, 1)).bfloat16()
    u = _rand((M, 16), 1.0, 2)
    b3 = _rand((16,), 0.1, 3)
    tr = t.float().requires_grad_(True)
    ur = u.clone().requires_grad_(True)
    g = torch.sigmoid(ur)
    q = (tr.view(M, 9, 16) * g.view(M, 1, 16)).view(B, S, S, 9, 16)
    s = torch.zeros((B, S, S, 16), device="cuda")
    for tap in range(9):
        oy, ox = tap // 3 - 1, tap % 3 - 1
        shifted = torch.zeros_like(s)
        ys, ye = max(0, -oy), min(S, S - oy)
        xs, xe = max(0, -ox), min(S, S - ox)
        shifted[:, ys:ye, xs:xe] = q[:, ys + oy:ye + oy, xs + ox:xe + ox, tap]
        s = s + shifted
    pred = torch.tanh(s + b3).permute(0, 3, 1, 2)[:, :heads].contiguous()
    dpred = _rand(pred.shape, 1.0, 4)
    pred.backward(dpred)
    gate = g.detach().bfloat16()
    db3 = torch.zeros(16, device="cuda")
    ds = ops.heads_ds(dpred.contiguous(), pred.detach().contiguous(), db3)
    db2 = torch.zeros(16, device="cuda")
    dt, du = ops.heads_bwd_stencil(t, ds, gate, B, S, S, db2)
    mask = torch.zeros(16, device="cuda")
    mask[:heads] = 1
    assert _rel(dt.float().view(M, 9, 16) * mask, tr.grad.view(M, 9, 16) * mask) < 3e-2
    assert _rel(du.float() * mask, ur.grad * mask) < 3e-2
    assert _rel(db2 * mask, ur.grad.sum(0) * mask) < 3e-2
    ds_ref = (dpred * (1 - pred.detach() ** 2)).sum(dim=(0, 2, 3))
    assert _rel(db3[:heads], ds_ref) < 1e-3
