"""GPU parity of what the fp16 training decoder added: fp16 / mixed-format tensor-core operands (mv_gemm_args.ab_f16),
fp16 variants of the memory-bound glue kernels, and the table-driven re-layout / device-side optimiser schedule kernels —
each against a plain torch fp32 reference of the same op."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
H, B16 = torch.float16, torch.bfloat16


def _ops():
    from miphei_vit_b200 import ops
    return ops


def _rand(shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(shape, generator=g, device="cuda") * scale


def _rel(got, ref):
    return (got.float() - ref.float()).abs().max().item() / (ref.float().abs().max().item() + 1e-12)


@pytest.mark.parametrize("M,N,K", [(1000, 256, 64), (5264, 1536, 512), (2100, 144, 32), (700, 48, 192)])
def test_gemm_fp16_operands(M, N, K):
    """kind::f16 UMMA with fp16 A and B: the fp32 output must equal the fp32 product of the ROUNDED operands to
    accumulation-order accuracy (this is what separates a format mix-up from a rounding difference)."""
    ops = _ops()
    a = _rand((M, K), 1.0, 1).to(H)
    b = _rand((N, K), 0.05, 2).to(H)
    out = ops.gemm(a, b, out_dtype=torch.float32)
    assert _rel(out, a.float() @ b.float().t()) < 2e-5
    # weight-gradient form (reduction index slow in both operands, MN-major B): out[N, 64] = b^T[N, M'] . c[M', 64]
    Mp = M // 8 * 8
    at = a[:Mp].t().contiguous()                      # [K, Mp] used as the A operand [rows = K, reduction = Mp]
    c = _rand((Mp, 64), 0.05, 3).to(H)
    out2 = torch.zeros((K, 64), device="cuda")
    ops.gemm(at, c, mode=ops.GEMM_NN_ATOMIC, out=out2)
    assert _rel(out2, at.float() @ c.float()) < 2e-5


def test_mixed_operand_formats_are_rejected_on_the_host():
    """a kind::f16 descriptor with different A / B formats is an illegal instruction on sm_100a (measured: it kills the
    context) — the C ABI and the wrapper must refuse it before anything is launched"""
    from miphei_vit_b200 import lib
    ops = _ops()
    a = _rand((256, 64), 1.0, 1).to(H)
    b = _rand((128, 64), 1.0, 2).to(B16)
    with pytest.raises(AssertionError):
        ops.gemm(a, b)
    args = lib.GemmArgs()
    out = torch.empty((256, 128), dtype=B16, device="cuda")
    args.a, args.lda, args.b, args.ldb = a.data_ptr(), 64, b.data_ptr(), 64
    args.m, args.n, args.k, args.out, args.ldo, args.ab_f16 = 256, 128, 64, out.data_ptr(), 128, 1
    import ctypes
    assert lib.init(0).mv_gemm_bf16(ctypes.byref(args), None) == -1
    assert "share one 16-bit format" in lib.last_error()


def test_conv_fwd_and_wgrad_fp16_maps():
    """implicit-GEMM conv over fp16 NHWC maps with fp16 weights, and its weight gradient with a bf16 dz^T A operand."""
    from miphei_vit_b200 import packing
    ops = _ops()
    Bn, Hh, C0, C1, Cout = 2, 32, 96, 128, 64
    x0, x1 = _rand((Bn, C0, Hh, Hh), 1.0, 1), _rand((Bn, C1, Hh, Hh), 1.0, 2)
    w = _rand((Cout, C0 + C1, 3, 3), 0.05, 3)
    n0, n1 = (t.permute(0, 2, 3, 1).contiguous().to(H) for t in (x0, x1))
    wp = packing.pack_conv3x3(w, [C0, C1], dtype=H)
    stats = torch.zeros((2, Cout), device="cuda")
    z = torch.empty((Bn * Hh * Hh, Cout), device="cuda")
    ops.gemm(n0, wp, conv=dict(stride=1, a2=n1), colstats=stats, out=z)
    xin = torch.cat([n0.float(), n1.float()], 3).permute(0, 3, 1, 2)
    ref = F.conv2d(xin, w.to(H).float(), padding=1)
    got = z.view(Bn, Hh, Hh, Cout).permute(0, 3, 1, 2)
    assert _rel(got, ref) < 2e-5
    assert torch.allclose(stats[0], ref.sum((0, 2, 3)), rtol=1e-3, atol=1e-2)
    dz = _rand((Bn * Hh * Hh, Cout), 1e-3, 4).bfloat16()
    dzT = ops.transpose_bf16(dz)
    M = dz.shape[0]
    kp = 9 * (128 + 128)
    dwp = torch.zeros((Cout, kp), device="cuda")
    b0, b1 = ops.f16_to_bf16(n0), ops.f16_to_bf16(n1)   # bf16 twins: both operands of the weight-gradient GEMM are bf16
    assert torch.equal(b0, n0.to(B16)) and b0.dtype == B16
    xin = torch.cat([b0.float(), b1.float()], 3).permute(0, 3, 1, 2)
    ops.gemm(dzT[:, :M], b0, mode=ops.GEMM_NN_ATOMIC, conv=dict(stride=1, a2=b1), out=dwp)
    xr = xin.clone().requires_grad_(False)
    wr = w.clone().requires_grad_(True)
    F.conv2d(xr, wr, padding=1).backward(dz.float().view(Bn, Hh, Hh, Cout).permute(0, 3, 1, 2))
    v = dwp.view(Cout, 9, 256)
    gw = torch.cat([v[:, :, :C0], v[:, :, 128:128 + C1]], 2).reshape(Cout, 3, 3, C0 + C1).permute(0, 3, 1, 2)
    assert _rel(gw, wr.grad) < 1e-4


def test_elementwise_fp16_variants():
    ops = _ops()
    # LayerNorm with fp16 output
    x = _rand((300, 256), 2.0, 1)
    w, b = 1 + _rand((256,), 0.1, 2), _rand((256,), 0.1, 3)
    y = torch.empty((300, 256), dtype=H, device="cuda")
    ops.layernorm_fwd(x, w, b, out=y)
    assert _rel(y, F.layer_norm(x, (256,), w, b, 1e-6)) < 1e-3
    # bilinear x2 on fp16 maps
    m = _rand((2, 64, 9, 9), 1.0, 4)
    up = ops.upsample2x(m.permute(0, 2, 3, 1).contiguous().to(H))
    assert up.dtype == H
    ref = F.interpolate(m.to(H).float(), scale_factor=2, mode="bilinear", align_corners=False)
    assert _rel(up.permute(0, 3, 1, 2), ref) < 1e-3
    # bicubic token map resize on fp16 tokens
    g, t, D, Bn = 18, 16, 128, 2
    tok = _rand((Bn * (g * g + 5), D), 1.0, 5).to(H)
    fm = ops.tokens_to_map(tok, Bn, g * g + 5, 5, g, t)
    assert fm.dtype == H
    src = tok.float().view(Bn, g * g + 5, D)[:, 5:].permute(0, 2, 1).reshape(Bn, D, g, g)
    ref = F.interpolate(src, scale_factor=(t / g, t / g), mode="bicubic")
    assert _rel(fm.permute(0, 3, 1, 2), ref) < 2e-3
    # BatchNorm + ReLU apply with fp16 output; the backward mask reads either format
    z = _rand((4096, 64), 1.0, 6)
    sc, sh = 1 + _rand((64,), 0.1, 7), _rand((64,), 0.1, 8)
    yh = torch.empty((4096, 64), dtype=H, device="cuda")
    ops.bn_relu_apply(z, sc, sh, out=yh)
    assert _rel(yh, torch.relu(z * sc + sh)) < 1e-3
    yb = ops.bn_relu_apply(z, sc, sh)
    dy = _rand((4096, 64), 1e-3, 9).bfloat16()
    mean, rstd, gamma = z.mean(0), (z.var(0, unbiased=False) + 1e-5).rsqrt(), 1 + _rand((64,), 0.1, 10)
    dz_h, s_h = ops.bn_relu_bwd(dy, yh, z, mean, rstd, gamma)
    dz_b, s_b = ops.bn_relu_bwd(dy, yb, z, mean, rstd, gamma)
    assert torch.equal((yh > 0), (yb > 0)) and torch.allclose(s_h, s_b, rtol=1e-4, atol=1e-7)  # same mask from either format
    assert _rel(dz_h, dz_b.float()) < 1e-2  # (column sums are atomic: last-bit differences between runs)
    # image staging: fp16 NHWC image
    xi = _rand((2, 3, 128, 128), 1.0, 11)
    img = torch.empty((2, 128, 128, 8), dtype=H, device="cuda")
    ops.prep_input(xi, img=img, want_patches=False)
    assert torch.equal(img[..., :3], xi.permute(0, 2, 3, 1).to(H)) and float(img[..., 3:].abs().sum()) == 0.0
    # gram of an fp16 map
    f = (_rand((5000, 32), 1.0, 12).relu() + 0.1).to(H)
    gram = ops.gram32(f)
    fd = f.double()
    assert torch.allclose(gram[:32].double(), fd.t() @ fd, rtol=2e-4, atol=0.1)
    assert torch.allclose(gram[32].double(), fd.sum(0), rtol=2e-4, atol=1e-2)
    # transpose of an fp16 map: converted to bf16 on the way (it meets bf16 gradients in the moment GEMMs), row of ones
    fT = ops.transpose_bf16(f, ones_row=True)
    assert fT.dtype == B16 and torch.equal(fT[:32, :5000], f.t().to(B16)) and bool((fT[32, :5000] == 1).all())


def test_gather_cast_add_i64_memset():
    ops = _ops()
    src = _rand((100000,), 1.0, 1)
    idx = torch.randint(-2, 100000, (250001,), device="cuda", dtype=torch.int32)
    for dt in (torch.float32, B16, H):
        dst = torch.full((250001,), 7.0, dtype=dt, device="cuda")
        ops.gather_cast(src, idx, dst)
        ref = torch.where(idx >= 0, src[idx.clamp(min=0).long()], torch.zeros((), device="cuda")).to(dt)
        ref = torch.where(idx == -2, torch.full((), 7.0, dtype=dt, device="cuda"), ref)
        assert torch.equal(dst, ref)
    c = torch.arange(23, dtype=torch.int64, device="cuda")
    ops.add_i64(c, 3)
    assert torch.equal(c, torch.arange(23, dtype=torch.int64, device="cuda") + 3)
    t = _rand((12345,), 1.0, 2)
    ops.memset(t)
    assert float(t.abs().sum()) == 0.0


def test_device_side_schedule_and_adam_match_torch_lambda_lr():
    """mv_adam_schedule + mv_adam_clip_step_dev against torch.optim.Adam + LambdaLR(pix2pix factor) + clip_grad_norm_."""
    from miphei_vit_b200.trainer import lr_lambda
    ops = _ops()
    n, total, warm, base = 50001, 12, 3, 8e-4
    p0 = _rand((n,), 1.0, 1)
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p_ref], lr=base, betas=(0.5, 0.999), eps=1e-7)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: lr_lambda(s, total, warm))
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    hyper = torch.zeros(4, device="cuda")
    for it in range(total):
        g = _rand((n,), 0.01 * (it + 1), 10 + it)
        p_ref.grad = g.clone()
        torch.nn.utils.clip_grad_norm_([p_ref], 1.0)
        lr_now = opt.param_groups[0]["lr"]
        opt.step()
        sched.step()
        nc = ops.grad_norm(g, 1.0)
        ops.adam_schedule(step, base, total, warm, 0.5, 0.999, hyper)
        ops.adam_clip_step_dev(p, g, m, v, nc, hyper)
        assert abs(hyper[2].item() - lr_now) < 1e-9 + 1e-6 * lr_now and int(step.item()) == it + 1
        assert (p - p_ref.detach()).abs().max().item() < 5e-6


def test_heads_bwd_algebra_matches_torch_formulas():
    """mv_heads_bwd_algebra against the closed-form expressions written out in torch (double precision)."""
    ops = _ops()
    n_units, n = 48, 131072.0
    E, FF = _rand((40, 256), 1e-3, 1), _rand((40, 32), 10.0, 2).abs()
    E[33:] = 0
    FF[33:] = 0
    W1, b1 = _rand((256, 32), 0.05, 3), _rand((256,), 0.1, 4)
    gam, w2 = 1 + _rand((256,), 0.1, 5), _rand((256,), 0.05, 6)
    fin = torch.stack([1 + _rand((256,), 0.1, 7), _rand((256,), 0.1, 8), _rand((256,), 0.1, 9), 1 + _rand((256,), 0.1, 10).abs()])
    fin = fin.contiguous()
    W1[n_units:] = 0
    dW1, dg, db, dw2 = (torch.zeros(s, device="cuda") for s in ((256, 32), (256,), (256,), (256,)))
    ca_t = torch.zeros((32, 256), dtype=B16, device="cuda")
    mx = torch.zeros((32, 64), dtype=H, device="cuda")
    ks, mxs = torch.zeros(32, device="cuda"), torch.zeros(32, device="cuda")
    ops.heads_bwd_algebra(E, FF, W1, b1, gam, w2, fin, n, n_units, dW1, dg, db, dw2, ca_t, mx, mxs, ks)
    d = lambda t: t.double()  # noqa: E731
    sl = slice(0, n_units)
    scale_g, shift_g, mean, rstd = (d(fin[i])[sl] for i in range(4))
    W, bb, gg, ww = d(W1)[sl], d(b1)[sl], d(gam)[sl], d(w2)[sl]
    EF, E1, F2, F1 = d(E[:32]).t()[sl], d(E[32])[sl], d(FF[:32]), d(FF[32])
    A = (W * EF).sum(1)
    S1 = ww * E1
    S2 = ww * rstd * (A + (bb - mean) * E1)
    XF = rstd[:, None] * (W @ F2 + (bb - mean)[:, None] * F1[None, :])
    gr = gg * rstd
    rdW1 = gr[:, None] * (ww[:, None] * EF - (S1 / n)[:, None] * F1[None, :] - (S2 / n)[:, None] * XF)
    Ca = (gr * ww)[:, None] * W
    K0 = ((gr * S1 / n)[:, None] * W).sum(0)
    k2 = gr * S2 / n
    Mx = W.t() @ ((k2 * rstd)[:, None] * W)
    K1 = ((k2 * rstd * (bb - mean))[:, None] * W).sum(0)
    assert _rel(dW1[sl], rdW1) < 1e-4 and _rel(dg[sl], S2) < 1e-4 and _rel(db[sl], S1) < 1e-5
    assert _rel(dw2[sl], scale_g * A + shift_g * E1) < 1e-4
    assert _rel(ca_t[:, sl].float().t(), Ca) < 5e-3 and float(ca_t[:, n_units:].abs().sum()) == 0.0
    # gradient-sized values (1e-10 here) would underflow fp16: stored times a power of two, reciprocal in mx_scale
    assert float(mxs.min()) == float(mxs.max()) and 1.0 <= float(mx[:, :32].abs().max()) < 2.0
    assert _rel(mx[:, :32].float() * mxs[0], (-Mx).t()) < 2e-3 and float(mx[:, 32:].abs().sum()) == 0.0
    assert _rel(ks, -(K0 + K1)) < 1e-4


def test_lora_refresh_matches_reference_layout():
    ops = _ops()
    D, L, alpha = 128, 3, 1.0
    flat = _rand((L * 32 * D,), 1.0, 1)
    acat = [torch.zeros((16, D), dtype=B16, device="cuda") for _ in range(L)]
    wext = [torch.zeros((3 * D, D + 64), dtype=B16, device="cuda") for _ in range(L)]
    wbwd = [torch.zeros((D, 3 * D + 64), dtype=B16, device="cuda") for _ in range(L)]
    bcat = [torch.zeros((16, 3 * D), dtype=B16, device="cuda") for _ in range(L)]
    ptrs = torch.tensor([[a.data_ptr(), w.data_ptr(), wb.data_ptr(), bc.data_ptr()] for a, w, wb, bc in zip(acat, wext, wbwd, bcat)],
                        dtype=torch.int64, device="cuda")
    ops.lora_refresh(flat, ptrs, L, D, alpha, D + 64, 3 * D + 64)
    v = flat.view(L, 4, 8 * D)
    for i in range(L):
        Aq, Bq, Av, Bv = v[i, 0].view(D, 8), v[i, 1].view(8, D), v[i, 2].view(D, 8), v[i, 3].view(8, D)
        assert torch.equal(acat[i][:8], Aq.t().to(B16)) and torch.equal(acat[i][8:], Av.t().to(B16))
        assert torch.equal(wext[i][:D, D:D + 8], (alpha * Bq).t().to(B16))
        assert torch.equal(wext[i][2 * D:, D + 8:D + 16], (alpha * Bv).t().to(B16))
        assert float(wext[i][:, :D].abs().sum()) == 0.0 and float(wext[i][D:2 * D].abs().sum()) == 0.0
        assert torch.equal(wbwd[i][:, 3 * D:3 * D + 8], Aq.to(B16)) and torch.equal(wbwd[i][:, 3 * D + 8:3 * D + 16], Av.to(B16))
        assert torch.equal(bcat[i][:8, :D], (alpha * Bq).to(B16)) and torch.equal(bcat[i][8:, 2 * D:], (alpha * Bv).to(B16))
        assert float(bcat[i][:8, D:].abs().sum()) == 0.0 and float(bcat[i][8:, :2 * D].abs().sum()) == 0.0
