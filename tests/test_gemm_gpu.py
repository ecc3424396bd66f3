"""GPU parity of the tcgen05 GEMM (mv_gemm_bf16) against a plain torch fp32 reference of the same op."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from miphei_vit_b200 import ops
    return ops


def _rand(shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale)


def _close(got, ref, tol):
    got = got.float()
    err = (got - ref).abs().max().item()
    den = ref.abs().max().item() + 1e-12
    assert err / den < tol, "max abs err %g (ref max %g, rel %g)" % (err, den, err / den)


@pytest.mark.parametrize("M,N,K,bn", [
    (128, 256, 64, 256), (128, 256, 128, 256), (256, 512, 256, 256), (128, 128, 64, 128),
    (5264, 4608, 1536, 0), (5264, 1536, 1536, 0), (5264, 1536, 4096, 0), (329, 1536, 1536, 128),
    (1000, 48, 64, 0), (777, 96, 448, 0), (640, 192, 896, 0), (300, 32, 640, 0), (300, 16, 1536, 0),
    (2000, 64, 1600, 0), (129, 256, 1600, 256), (5264, 4608, 1552, 192), (700, 384, 128, 192), (3000, 1728, 256, 0),
    (2100, 144, 32, 0),
])
def test_gemm_plain(M, N, K, bn):
    ops = _ops()
    a = _rand((M, K), 1.0, 1).bfloat16()
    b = _rand((N, K), 0.05, 2).bfloat16()
    out = ops.gemm(a, b, block_n=bn)
    ref = a.float() @ b.float().t()
    _close(out, ref, 1e-2)
    outf = ops.gemm(a, b, block_n=bn, out_dtype=torch.float32)
    _close(outf, ref, 2e-5)


def test_gemm_epilogue_linear():
    ops = _ops()
    M, N, K = 700, 1536, 512
    a = _rand((M, K), 1.0, 1).bfloat16()
    b = _rand((N, K), 0.05, 2).bfloat16()
    scale = _rand((N,), 1.0, 3)
    shift = _rand((N,), 1.0, 4)
    resid = _rand((M, N), 1.0, 5)
    aux = torch.zeros((M, N), dtype=torch.bfloat16, device="cuda")
    out = ops.gemm(a, b, scale=scale, shift=shift, resid=resid, out_dtype=torch.float32, aux=aux)
    ref = (a.float() @ b.float().t()) * scale + shift + resid
    _close(out, ref, 2e-5)
    _close(aux, ref, 1e-2)
    out2 = ops.gemm(a, b, scale=scale, shift=shift, act=ops.ACT_RELU)
    _close(out2, torch.relu((a.float() @ b.float().t()) * scale + shift), 1e-2)
    # in-place residual update (out aliases resid)
    r2 = resid.clone()
    ops.gemm(a, b, shift=shift, resid=r2, out=r2)
    _close(r2, (a.float() @ b.float().t()) + shift + resid, 2e-5)


def test_gemm_row_remap():
    """patch-embedding form: rows regrouped into token rows with a per-group offset, residual indexed modulo."""
    ops = _ops()
    B, P, N, K = 3, 324, 1536, 640
    a = _rand((B * P, K), 1.0, 1).bfloat16()
    b = _rand((N, K), 0.05, 2).bfloat16()
    shift = _rand((N,), 1.0, 4)
    pos = _rand((P, N), 1.0, 5)
    out = torch.full((B * 329, N), 7.0, dtype=torch.float32, device="cuda")
    ops.gemm(a, b, shift=shift, resid=pos, out=out, rows_per_group=P, group_stride=329, row_offset=5, resid_row_mod=True)
    ref = ((a.float() @ b.float().t()) + shift).view(B, P, N) + pos
    got = out.view(B, 329, N)
    _close(got[:, 5:], ref, 2e-5)
    assert (got[:, :5] == 7.0).all()


def test_gemm_swiglu_fwd_bwd():
    ops = _ops()
    M, H, K = 600, 4096, 1536
    a = _rand((M, K), 1.0, 1).bfloat16()
    w = _rand((2 * H, K), 0.03, 2).bfloat16()
    bias = _rand((2 * H,), 0.5, 3)
    h = torch.zeros((M, 2 * H), dtype=torch.bfloat16, device="cuda")
    u = ops.gemm(a, w, mode=ops.GEMM_SWIGLU, shift=bias, aux=h)
    pre = a.float() @ w.float().t() + bias
    ref = torch.nn.functional.silu(pre[:, :H]) * pre[:, H:]
    _close(u, ref, 1e-2)
    _close(h, pre, 1e-2)
    # backward epilogue: dH from dU = dY @ W2 with saved pre-activations
    dy = _rand((M, 1536), 1.0, 6).bfloat16()
    w2t = _rand((H, 1536), 0.03, 7).bfloat16()
    dh = ops.gemm(dy, w2t, mode=ops.GEMM_SWIGLU_BWD, in2=h)
    du = dy.float() @ w2t.float().t()
    hg, hv = h.float()[:, :H], h.float()[:, H:]
    sg = torch.sigmoid(hg)
    ref_dg = du * hv * (sg * (1 + hg * (1 - sg)))
    ref_dv = du * hg * sg
    _close(dh[:, :H], ref_dg, 1e-2)
    _close(dh[:, H:], ref_dv, 1e-2)


def test_gemm_strided_operands():
    ops = _ops()
    M, N, K = 500, 256, 1536
    abuf = _rand((M, 1600), 1.0, 1).bfloat16()
    b = _rand((N, 1600), 0.05, 2).bfloat16()
    out = ops.gemm(abuf[:, :K], b[:, :K])
    _close(out, abuf[:, :K].float() @ b[:, :K].float().t(), 1e-2)
    obuf = torch.zeros((M, 512), dtype=torch.bfloat16, device="cuda")
    ops.gemm(abuf, b, out=obuf[:, 256:])
    _close(obuf[:, 256:], abuf.float() @ b.float().t(), 1e-2)
    assert (obuf[:, :256] == 0).all()


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (512, 512, 256), (5264, 4608, 1552), (5264, 1536, 4096), (300, 768, 128),
                                   (10528, 1536, 1536)])
def test_gemm_cta_pair(M, N, K):
    """cta_group::2 tiles (two CTAs share a 256-row tile, B halves read from the peer's smem) vs the 1-CTA kernel."""
    ops = _ops()
    a = _rand((M, K), 1.0, 1).bfloat16()
    b = _rand((N, K), 0.05, 2).bfloat16()
    shift = _rand((N,), 1.0, 3)
    resid = _rand((M, N), 1.0, 4)
    ref = a.float() @ b.float().t() + shift + resid
    one = ops.gemm(a, b, shift=shift, resid=resid, out_dtype=torch.float32, block_n=256, pair=1)
    two = ops.gemm(a, b, shift=shift, resid=resid, out_dtype=torch.float32, block_n=256, pair=2)
    _close(two, ref, 2e-5)
    assert torch.equal(one, two)  # same accumulation order per output element
    _close(ops.gemm(a, b, block_n=256, pair=2), a.float() @ b.float().t(), 1e-2)
    if N % 192 == 0:
        _close(ops.gemm(a, b, block_n=192, pair=2), a.float() @ b.float().t(), 1e-2)
        _close(ops.gemm(a, b, shift=shift, resid=resid, out_dtype=torch.float32, block_n=192, pair=2), ref, 2e-5)


def test_gemm_cta_pair_swiglu():
    ops = _ops()
    M, H, K = 1500, 4096, 1536
    a = _rand((M, K), 1.0, 1).bfloat16()
    w = _rand((2 * H, K), 0.03, 2).bfloat16()
    bias = _rand((2 * H,), 0.5, 3)
    h1 = torch.zeros((M, 2 * H), dtype=torch.bfloat16, device="cuda")
    h2 = torch.zeros_like(h1)
    u1 = ops.gemm(a, w, mode=ops.GEMM_SWIGLU, shift=bias, aux=h1, pair=1)
    u2 = ops.gemm(a, w, mode=ops.GEMM_SWIGLU, shift=bias, aux=h2, pair=2)
    assert torch.equal(u1, u2) and torch.equal(h1, h2)
    dy = _rand((M, 1536), 1.0, 6).bfloat16()
    w2t = _rand((H, 1536), 0.03, 7).bfloat16()
    d1 = ops.gemm(dy, w2t, mode=ops.GEMM_SWIGLU_BWD, in2=h1, pair=1)
    d2 = ops.gemm(dy, w2t, mode=ops.GEMM_SWIGLU_BWD, in2=h1, pair=2)
    assert torch.equal(d1, d2)


@pytest.mark.parametrize("M,N,K", [(5264, 1536, 1536), (5264, 4608, 1536), (5264, 1536, 4096), (10528, 1536, 1536), (2000, 768, 512),
                                   (1300, 1536, 256)])
def test_gemm_stream_k_matches_whole_tile_schedule(M, N, K):
    """the last partial wave of tiles split along K over all SMs (partials reduced at L2, last arriver runs the epilogue)
    gives the same result as the whole-tile schedule — fp32 reassociation only — launch after launch (the workspace is
    restored to zero by the kernel itself), for every fused epilogue that uses it"""
    ops = _ops()
    a = _rand((M, K), 1.0, 1).bfloat16()
    b = _rand((N, K), 0.05, 2).bfloat16()
    scale, shift, resid = _rand((N,), 1.0, 3), _rand((N,), 1.0, 4), _rand((M, N), 1.0, 5)
    ref = (a.float() @ b.float().t()) * scale + shift + resid
    base = ops.gemm(a, b, scale=scale, shift=shift, resid=resid, out_dtype=torch.float32, stream_k=False)
    for _ in range(3):
        out = ops.gemm(a, b, scale=scale, shift=shift, resid=resid, out_dtype=torch.float32, stream_k="force")
        _close(out, ref, 2e-5)
        assert (out - base).abs().max().item() <= 1e-4 * ref.abs().max().item()
    outb = ops.gemm(a, b, shift=shift, stream_k="force")
    _close(outb, a.float() @ b.float().t() + shift, 1e-2)
    ws = ops._sk_workspace(a.device)
    assert ws is not None and int(ws.count_nonzero()) == 0          # zeros restored, counters back to 0


def test_gemm_stream_k_swiglu_fwd_bwd():
    ops = _ops()
    M, D, H = 5264, 1536, 1024
    x = _rand((M, D), 1.0, 1).bfloat16()
    w1 = _rand((2 * H, D), 0.03, 2).bfloat16()
    b1 = _rand((2 * H,), 0.1, 3)
    aux0 = torch.empty((M, 2 * H), dtype=torch.bfloat16, device="cuda")
    aux1 = torch.empty_like(aux0)
    u0 = ops.gemm(x, w1, mode=ops.GEMM_SWIGLU, shift=b1, aux=aux0, stream_k=False)
    u1 = ops.gemm(x, w1, mode=ops.GEMM_SWIGLU, shift=b1, aux=aux1, stream_k="force")
    h = x.float() @ w1.float().t() + b1
    ref = torch.nn.functional.silu(h[:, :H]) * h[:, H:]
    _close(u1, ref, 1.5e-2)
    assert (u1.float() - u0.float()).abs().max().item() <= 2e-2 * ref.abs().max().item()
    assert (aux1.float() - aux0.float()).abs().max().item() <= 2e-2 * h.abs().max().item()
    w2t = _rand((H, D), 0.03, 5).bfloat16()          # dU = dX . W2^T: [M, D] x [H, D]^T
    dx = _rand((M, D), 1.0, 6).bfloat16()
    g0 = ops.gemm(dx, w2t, mode=ops.GEMM_SWIGLU_BWD, in2=aux0, stream_k=False)
    g1 = ops.gemm(dx, w2t, mode=ops.GEMM_SWIGLU_BWD, in2=aux0, stream_k="force")
    assert (g1.float() - g0.float()).abs().max().item() <= 2e-2 * g0.float().abs().max().item()
    assert int(ops._sk_workspace(x.device).count_nonzero()) == 0


@pytest.mark.parametrize("M,N,K", [(5264, 1536, 1536), (5264, 1536, 4096), (10528, 1536, 1536), (1040, 512, 128), (2049, 768, 640)])
def test_gemm_fp32_residual_tma_epilogue(M, N, K):
    """attn.proj / fc2 form: x_out = (a @ w^T) * ls + ls*b + x, fp32 residual stream.  On CTA pairs the residual tiles come in
    and the results go out as TMA bulk copies through swizzled shared memory (no per-thread global access); the single-CTA
    kernel keeps the register epilogue — both must match the fp32 reference, out of place and in place, M tails included."""
    ops = _ops()
    a = _rand((M, K), 1.0, 1).bfloat16()
    w = _rand((N, K), 0.05, 2).bfloat16()
    scale, shift, resid = _rand((N,), 0.3, 3), _rand((N,), 0.3, 4), _rand((M, N), 1.0, 5)
    ref = (a.float() @ w.float().t()) * scale + shift + resid
    old = ops.gemm(a, w, scale=scale, shift=shift, resid=resid, out_dtype=torch.float32, block_n=256, pair=1)
    new = ops.gemm(a, w, scale=scale, shift=shift, resid=resid, out_dtype=torch.float32, block_n=256, pair=2)
    _close(new, ref, 2e-5)
    assert (new - old).abs().max().item() <= 2e-6 * ref.abs().max().item()
    x = resid.clone()                                   # in place: out aliases the residual (what the engine does)
    for _ in range(2):
        ops.gemm(a, w, scale=scale, shift=shift, resid=x, out=x, pair=2)
    _close(x, ref + (ref - resid), 4e-5)
    nos = ops.gemm(a, w, resid=resid, out_dtype=torch.float32, pair=2)   # no scale / shift
    _close(nos, a.float() @ w.float().t() + resid, 2e-5)
    big = torch.full((M + 64, N + 32), 7.0, device="cuda")               # strided output / residual views, guard band untouched
    view = big[32:32 + M, 32:32 + N]
    view.copy_(resid)
    ops.gemm(a, w, scale=scale, shift=shift, resid=view, out=view, pair=2)
    _close(view, ref, 2e-5)
    assert float((big[:32] - 7).abs().max()) == 0 and float((big[32 + M:] - 7).abs().max()) == 0
    assert float((big[:, :32] - 7).abs().max()) == 0


@pytest.mark.parametrize("M,N,K", [(5264, 4608, 1536), (10528, 1536, 1536), (1040, 512, 128), (2049, 776, 640)])
def test_gemm_bf16_out_tma_epilogue(M, N, K):
    """QKV / dX form on CTA pairs: act(acc * scale + shift) -> bf16 through swizzled shared memory and TMA stores; must agree
    with the register epilogue of the single-CTA kernel bit for bit (same fp32 math, same rounding), tails and strides included"""
    ops = _ops()
    a = _rand((M, K), 1.0, 1).bfloat16()
    w = _rand((N, K), 0.05, 2).bfloat16()
    scale, shift = _rand((N,), 0.5, 3), _rand((N,), 0.5, 4)
    ref = (a.float() @ w.float().t()) * scale + shift
    for act in (ops.ACT_NONE, ops.ACT_RELU):
        r = torch.relu(ref) if act == ops.ACT_RELU else ref
        old = ops.gemm(a, w, scale=scale, shift=shift, act=act, block_n=256, pair=1)
        new = ops.gemm(a, w, scale=scale, shift=shift, act=act, block_n=256, pair=2)
        _close(new, r, 1e-2)
        assert torch.equal(old, new)
    big = torch.full((M + 8, N + 16), 3.0, dtype=torch.bfloat16, device="cuda")   # strided output view, guard band untouched
    view = big[8:, 8:8 + N]
    ops.gemm(a, w, shift=shift, out=view, pair=2)
    _close(view, a.float() @ w.float().t() + shift, 1e-2)
    assert float((big[:8].float() - 3).abs().max()) == 0 and float((big[:, :8].float() - 3).abs().max()) == 0
    assert float((big[:, 8 + N:].float() - 3).abs().max()) == 0


@pytest.mark.parametrize("M,N,K", [(8192, 256, 32), (5000, 144, 32), (4100, 256, 64), (6000, 200, 96), (70000, 144, 32), (1048576, 144, 32)])
def test_gemm_skinny_k_tma_epilogue(M, N, K):
    """One or two K blocks over many rows (the decoder's 32-channel maps): CTA pairs, eight epilogue warps, TMA stores.
    Must agree bit for bit with the register epilogue of the 128-wide tiles (pair=1 keeps that schedule), N tails included."""
    ops = _ops()
    a = _rand((M, K), 1.0, 1).bfloat16()
    w = _rand((N, K), 0.2, 2).bfloat16()
    scale, shift = 1 + _rand((N,), 0.1, 3), _rand((N,), 0.3, 4)
    ref = (a.float() @ w.float().t()) * scale + shift
    for act in (ops.ACT_NONE, ops.ACT_RELU):
        r = torch.relu(ref) if act == ops.ACT_RELU else ref
        old = ops.gemm(a, w, scale=scale, shift=shift, act=act, pair=1)
        new = ops.gemm(a, w, scale=scale, shift=shift, act=act)
        _close(new, r, 1e-2)
        assert torch.equal(old, new)
    plain = ops.gemm(a, w)
    assert torch.equal(plain, ops.gemm(a, w, pair=1))
    if N % 16 == 0:  # GATE_MASK: out[m, n] = du[m, n / 16] where the unit is active
        du = _rand((M, N // 16), 1.0, 5).bfloat16()
        dup = torch.zeros((M, 16), dtype=torch.bfloat16, device="cuda")
        dup[:, :N // 16] = du
        old = ops.gemm(a, w, scale=scale, shift=shift, act=ops.ACT_GATE_MASK, in2=dup, pair=1)
        new = ops.gemm(a, w, scale=scale, shift=shift, act=ops.ACT_GATE_MASK, in2=dup)
        assert torch.equal(old, new)
        want = torch.where(ref > 0, du.float().repeat_interleave(16, dim=1), torch.zeros_like(ref))
        assert ((new.float() - want).abs() > 1e-6).float().mean().item() < 2e-3  # sign flips at rounding level only
    big = torch.full((M + 8, N + 24), 3.0, dtype=torch.bfloat16, device="cuda")   # strided output view, guard band untouched
    view = big[8:, 8:8 + N]
    ops.gemm(a, w, shift=shift, out=view)
    _close(view, a.float() @ w.float().t() + shift, 1e-2)
    assert float((big[:8].float() - 3).abs().max()) == 0 and float((big[:, :8].float() - 3).abs().max()) == 0
    assert float((big[:, 8 + N:].float() - 3).abs().max()) == 0


@pytest.mark.parametrize("M,N,K", [(10528, 1536, 1536), (10528, 1536, 4096), (5264, 4608, 1536), (4000, 768, 256)])
def test_gemm_tail_retiling(M, N, K):
    """optional schedule of the big CTA-pair GEMMs (measured slower, off by default): full waves of 256 x 256 tiles, then the
    partial last wave re-launched as 128 x 128 single-CTA tiles over its rectangle (no reduction). Same results as the
    single-CTA kernel, fp32-residual and bf16 forms, in place too."""
    ops = _ops()
    a = _rand((M, K), 1.0, 1).bfloat16()
    w = _rand((N, K), 0.05, 2).bfloat16()
    scale, shift, resid = _rand((N,), 0.3, 3), _rand((N,), 0.3, 4), _rand((M, N), 1.0, 5)
    ref = (a.float() @ w.float().t()) * scale + shift + resid
    old = ops.gemm(a, w, scale=scale, shift=shift, resid=resid, out_dtype=torch.float32, block_n=256, pair=1)
    new = ops.gemm(a, w, scale=scale, shift=shift, resid=resid, out_dtype=torch.float32, tail=True)
    _close(new, ref, 2e-5)
    assert (new - old).abs().max().item() <= 2e-6 * ref.abs().max().item()
    x = resid.clone()
    ops.gemm(a, w, scale=scale, shift=shift, resid=x, out=x, tail=True)
    assert torch.equal(x, new)
    oldb = ops.gemm(a, w, scale=scale, shift=shift, act=ops.ACT_RELU, block_n=256, pair=1)
    newb = ops.gemm(a, w, scale=scale, shift=shift, act=ops.ACT_RELU, tail=True)
    assert torch.equal(oldb, newb)
