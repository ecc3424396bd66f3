"""Tensor-level wrappers over the C ABI: they take torch CUDA tensors, pass raw pointers / strides / the current
stream to libmiphei_b200.so and return torch tensors.  No arithmetic happens in Python or in torch here."""
import ctypes

import torch

from . import lib as _lib

_H16 = (torch.bfloat16, torch.float16)
GEMM_LINEAR, GEMM_SWIGLU, GEMM_SWIGLU_BWD, GEMM_HEAD_GATE, GEMM_HEAD_CONV, GEMM_NN_ATOMIC = 0, 1, 2, 3, 4, 5
ACT_NONE, ACT_RELU, ACT_GATE_MASK = 0, 1, 2


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def require_cuda(device, what):
    """Every entry into the kernels goes through here: no CUDA device, no computation (there is no CPU path)."""
    if device.type != "cuda":
        raise _lib.MipheiB200Error("%s needs a CUDA device (got %s); miphei_b200 has no CPU path" % (what, device))


def _lib_for(t):
    require_cuda(t.device, "miphei_b200 ops")
    return _lib.init(t.device.index if t.device.index is not None else torch.cuda.current_device())


def _rowmajor(t, name):
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError("%s must be 2-D with unit inner stride, got shape %s stride %s" % (name, tuple(t.shape), t.stride()))
    return t


_SK_BYTES = 20 * 1024 * 1024 + 4096
_sk_workspaces = {}


def _sk_workspace(device):
    """stream-K workspace of the CURRENT stream on this device (zeroed once; the kernels keep it zero between launches).
    Keyed by stream: GEMMs that can run concurrently never share one."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _sk_workspaces.get(key)
    if ws is None:
        if torch.cuda.is_current_stream_capturing():
            return None  # never allocate inside a graph capture: warm-up runs create the workspace of a capture stream
        ws = torch.zeros(_SK_BYTES, dtype=torch.uint8, device=device)
        _sk_workspaces[key] = ws
    return ws


def gemm(a, b, *, mode=GEMM_LINEAR, out=None, out_dtype=torch.bfloat16, out_rows=None, scale=None, shift=None,
         resid=None, act=ACT_NONE, aux=None, in2=None, rows_per_group=0, group_stride=0, row_offset=0,
         resid_row_mod=False, block_n=0, conv=None, out_kind=None, colstats=None, no_out=False, pair=0,
         kskip=None, stream_k=False, tail=False):
    """out = epilogue(a[M,K] @ b[N,K]^T) on the tcgen05 GEMM (mv_gemm_bf16). See include/miphei_b200.h.
    stream_k: False (default) | True (when the library judges it worthwhile) | "force" — measured slower on every encoder
    shape (the L2 reduction of the partial tiles costs more than the idle last wave), kept for small grids."""
    lib = _lib_for(a)
    # 16-bit operands: bf16 (default) or fp16 (training-mode decoder maps / weights); the formats are independent
    assert a.dtype in _H16 and b.dtype in _H16
    if mode == GEMM_NN_ATOMIC and conv is not None:
        N = Kb = None
    elif mode == GEMM_NN_ATOMIC:
        _rowmajor(b, "b")
        Kb, N = b.shape  # B is [K, N] row-major
    else:
        _rowmajor(b, "b")
        N, Kb = b.shape
    a2 = None
    if conv is not None and mode == GEMM_NN_ATOMIC:
        # weight gradient: a = dz^T [Cout, pixels]; b (and conv["a2"]) = NHWC input maps of the conv
        stride = conv.get("stride", 1)
        a2 = conv.get("a2")
        _rowmajor(a, "a")
        assert b.dim() == 4 and b.is_contiguous()
        Bc, Hin, Win, C0 = b.shape
        C1 = a2.shape[3] if a2 is not None else 0
        Ho, Wo = Hin // stride, Win // stride
        M, K = a.shape
        N, Kb = 9 * 64 * ((C0 + 63) // 64 + (C1 + 63) // 64), Bc * Ho * Wo
    elif conv is not None:
        # a (and optionally conv["a2"]) are contiguous NHWC maps [B, Hin, Win, C]; conv = dict(stride=1|2, a2=None)
        stride = conv.get("stride", 1)
        a2 = conv.get("a2")
        assert a.dim() == 4 and a.is_contiguous() and (a2 is None or (a2.is_contiguous() and a2.shape[:3] == a.shape[:3]))
        Bc, Hin, Win, C0 = a.shape
        C1 = a2.shape[3] if a2 is not None else 0
        Ho, Wo = Hin // stride, Win // stride
        M, K = Bc * Ho * Wo, 9 * 64 * ((C0 + 63) // 64 + (C1 + 63) // 64)
    else:
        _rowmajor(a, "a")
        M, K = a.shape
    assert K == Kb, ((M, K), b.shape)
    if mode == GEMM_SWIGLU:
        out_cols = N // 2
    elif mode == GEMM_SWIGLU_BWD:
        out_cols = 2 * N
    else:
        out_cols = N
    if mode == GEMM_HEAD_GATE:
        out_cols = N // 16
    if no_out:
        assert colstats is not None and out is None
    elif mode == GEMM_HEAD_CONV:
        if out is None:
            out = torch.empty((Bc, N, Ho, Wo), dtype=out_dtype, device=a.device)
        assert out.is_contiguous()
    else:
        if out is None:
            out = torch.empty((out_rows if out_rows is not None else M, out_cols), dtype=out_dtype, device=a.device)
        _rowmajor(out, "out")
    args = _lib.GemmArgs()
    args.a, args.lda = a.data_ptr(), (a.stride(0) if a.dim() == 2 else 0)
    if conv is not None:
        args.conv = 1
        args.a2 = a2.data_ptr() if a2 is not None else None
        args.conv_batch, args.conv_h, args.conv_w, args.conv_stride = Bc, Ho, Wo, stride
        args.conv_c0, args.conv_c1 = C0, C1
    args.b, args.ldb = b.data_ptr(), (b.stride(0) if b.dim() == 2 else 0)
    args.m, args.n, args.k = M, N, K
    args.mode, args.act = mode, act
    if colstats is not None:
        assert colstats.dtype == torch.float32 and colstats.is_contiguous() and colstats.numel() == 2 * N
        args.colstats = colstats.data_ptr()
    if no_out:
        args.out_f32, args.out, args.ldo = 0, None, 0
    else:
        args.out_f32 = 1 if out.dtype == torch.float32 else (2 if out.dtype == torch.uint8 else 0)
        args.out, args.ldo = out.data_ptr(), (out.stride(0) if mode != GEMM_HEAD_CONV else 0)
    if aux is not None:
        args.aux, args.ldaux = aux.data_ptr(), aux.stride(0)
    if scale is not None:
        assert scale.dtype == torch.float32 and scale.is_contiguous()
        args.scale = scale.data_ptr()
    if shift is not None:
        assert shift.dtype == torch.float32 and shift.is_contiguous()
        args.shift = shift.data_ptr()
    if resid is not None:
        assert resid.dtype == torch.float32
        args.resid, args.ldr = resid.data_ptr(), (resid.stride(0) if resid.dim() == 2 else 0)
    if in2 is not None:
        assert in2.dtype == (torch.float32 if mode == GEMM_HEAD_GATE else torch.bfloat16)
        args.in2, args.ldin2 = in2.data_ptr(), (in2.stride(0) if in2.dim() == 2 else 0)
    args.rows_per_group, args.group_stride, args.row_offset = rows_per_group, group_stride, row_offset
    args.resid_row_mod = 1 if resid_row_mod else 0
    args.block_n = block_n
    args.reserved2 = pair  # 0 auto, 1 never, 2 always: CTA-pair (cta_group::2) tiles
    a2t = conv.get("a2") if conv is not None else None
    # both tensor-core operands must share one format (a mixed descriptor is an illegal instruction on sm_100a)
    assert a.dtype == b.dtype and (a2t is None or a2t.dtype == a.dtype), (a.dtype, b.dtype)
    args.ab_f16 = 3 if a.dtype == torch.float16 else 0
    if mode in (GEMM_LINEAR, GEMM_SWIGLU, GEMM_SWIGLU_BWD) and conv is None and M >= 1024 and stream_k:
        ws = _sk_workspace(a.device)
        if ws is not None:
            args.workspace, args.workspace_bytes = ws.data_ptr(), ws.numel()
            args.reserved3 = 1 if stream_k == "force" else 0
    if tail:
        args.reserved3 = 3
    if kskip is not None:  # K range with all-zero B columns: never loaded
        args.kskip_begin, args.kskip_end = int(kskip[0]), int(kskip[1])
    _lib.check(lib.mv_gemm_bf16(ctypes.byref(args), _stream()), "mv_gemm_bf16")
    return out


def layernorm_fwd(x, w, b, *, out=None, eps=1e-6, stats=False):
    """y = LayerNorm(x) (mv_layernorm_fwd): x fp32 [M, D] -> bf16 [M, D] (out may be a wider, strided buffer)."""
    lib = _lib_for(x)
    _rowmajor(x, "x")
    assert x.dtype == torch.float32 and w.dtype == torch.float32 and b.dtype == torch.float32
    M, D = x.shape
    if out is None:
        out = torch.empty((M, D), dtype=torch.bfloat16, device=x.device)
    mean = rstd = None
    if stats:
        mean = torch.empty(M, dtype=torch.float32, device=x.device)
        rstd = torch.empty(M, dtype=torch.float32, device=x.device)
    assert out.dtype in _H16
    _lib.check(lib.mv_layernorm_fwd(_ptr(x), x.stride(0), _ptr(w), _ptr(b), _ptr(out), out.stride(0),
                                    1 if out.dtype == torch.float16 else 0, _ptr(mean), _ptr(rstd), M, D, eps, _stream()),
               "mv_layernorm_fwd")
    return (out, mean, rstd) if stats else out


def layernorm_bwd(x, w, dy, *, dres=None, eps=1e-6, want_bf16=False, out=None, out_bf16=None):
    """dx = dres + dLN(dy) (mv_layernorm_bwd); returns fp32 dx (and a bf16 copy when want_bf16)."""
    lib = _lib_for(x)
    M, D = x.shape
    dx = out if out is not None else torch.empty((M, D), dtype=torch.float32, device=x.device)
    dxb = out_bf16 if out_bf16 is not None else (
        torch.empty((M, D), dtype=torch.bfloat16, device=x.device) if want_bf16 else None)
    want_bf16 = dxb is not None
    _lib.check(lib.mv_layernorm_bwd(_ptr(x), x.stride(0), _ptr(w), _ptr(dy), dy.stride(0),
                                    1 if dy.dtype == torch.float32 else 0, _ptr(dres),
                                    dres.stride(0) if dres is not None else 0, _ptr(dx), dx.stride(0), _ptr(dxb),
                                    dxb.stride(0) if dxb is not None else 0, M, D, eps, _stream()), "mv_layernorm_bwd")
    return (dx, dxb) if want_bf16 else dx


def attn_fwd(qkv, batch, n_tok, heads, *, out=None, want_lse=False, scale=None, lse=None):
    """softmax(q k^T * scale) v per (image, head) from fused qkv rows (mv_attn_fwd)."""
    lib = _lib_for(qkv)
    _rowmajor(qkv, "qkv")
    assert qkv.dtype == torch.bfloat16 and qkv.shape == (batch * n_tok, 3 * heads * 64)
    if out is None:
        out = torch.empty((batch * n_tok, heads * 64), dtype=torch.bfloat16, device=qkv.device)
    if lse is None and want_lse:
        lse = torch.empty((batch, heads, n_tok), dtype=torch.float32, device=qkv.device)
    if scale is None:
        scale = 64 ** -0.5
    _lib.check(lib.mv_attn_fwd(_ptr(qkv), qkv.stride(0), _ptr(out), out.stride(0), _ptr(lse), batch, n_tok, heads,
                               float(scale), _stream()), "mv_attn_fwd")
    return (out, lse) if want_lse else out


def prep_input(x, want_image=True, want_patches=True, img=None, pm=None):
    """x fp32 NCHW -> (NHWC bf16 image padded to 8 channels, patch matrix [B*g*g, 592]) (mv_prep_input)."""
    lib = _lib_for(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] == 3 and x.shape[2] == x.shape[3]
    B, _, S, _ = x.shape
    g = S // 14
    if img is None and want_image:
        img = torch.empty((B, S, S, 8), dtype=torch.bfloat16, device=x.device)
    if pm is None and want_patches:
        pm = torch.empty((B * g * g, 592), dtype=torch.bfloat16, device=x.device)
    f16 = 1 if (img is not None and img.dtype == torch.float16) else 0
    _lib.check(lib.mv_prep_input(_ptr(x), _ptr(img), f16, _ptr(pm), B, S, 592, _stream()), "mv_prep_input")
    return img, pm


HOPTIMUS_MEAN = (0.707223, 0.578729, 0.703617)  # src/dataset.py:601
HOPTIMUS_STD = (0.211883, 0.230117, 0.177517)


def prep_input_u8(tiles, img=None, pm=None, mean=HOPTIMUS_MEAN, std=HOPTIMUS_STD, want_image=True, want_patches=True):
    """raw uint8 NHWC tiles [B, S, S, 3] -> (NHWC bf16 image, patch matrix), normalised on the device (mv_prep_input_u8)."""
    lib = _lib_for(tiles)
    assert tiles.dtype == torch.uint8 and tiles.is_contiguous() and tiles.dim() == 4 and tiles.shape[3] == 3
    B, S = tiles.shape[0], tiles.shape[1]
    g = S // 14
    if img is None and want_image:
        img = torch.empty((B, S, S, 8), dtype=torch.bfloat16, device=tiles.device)
    if pm is None and want_patches:
        pm = torch.empty((B * g * g, 592), dtype=torch.bfloat16, device=tiles.device)
    sc = (ctypes.c_float * 3)(*[1.0 / (255.0 * s) for s in std])
    bi = (ctypes.c_float * 3)(*[-m / s for m, s in zip(mean, std)])
    f16 = 1 if (img is not None and img.dtype == torch.float16) else 0
    _lib.check(lib.mv_prep_input_u8(_ptr(tiles), sc, bi, _ptr(img), f16, _ptr(pm), B, S, 592, _stream()), "mv_prep_input_u8")
    return img, pm


def fill_prefix(x_res, prefix, batch, n_tok):
    lib = _lib_for(x_res)
    assert x_res.dtype == torch.float32 and prefix.dtype == torch.float32 and prefix.is_contiguous()
    _lib.check(lib.mv_fill_prefix(_ptr(x_res), x_res.stride(0), _ptr(prefix), batch, n_tok, prefix.shape[0],
                                  prefix.shape[1], _stream()), "mv_fill_prefix")


def tokens_to_map(tokens, batch, n_tok, prefix, grid, target, out=None):
    """final-norm tokens bf16 [B*n_tok, D] -> NHWC bf16 [B, target, target, D], bicubic (mv_tokens_to_map)."""
    lib = _lib_for(tokens)
    D = tokens.shape[1]
    if out is None:
        out = torch.empty((batch, target, target, D), dtype=tokens.dtype, device=tokens.device)
    assert tokens.dtype in _H16 and out.dtype == tokens.dtype
    _lib.check(lib.mv_tokens_to_map(_ptr(tokens), tokens.stride(0), _ptr(out), batch, n_tok, prefix, grid, target, D,
                                    1 if tokens.dtype == torch.float16 else 0, _stream()), "mv_tokens_to_map")
    return out


def upsample2x(x, out=None):
    """bilinear x2, NHWC bf16 (mv_upsample2x)."""
    lib = _lib_for(x)
    assert x.dtype in _H16 and x.is_contiguous() and x.dim() == 4
    B, h, w, C = x.shape
    if out is None:
        out = torch.empty((B, 2 * h, 2 * w, C), dtype=x.dtype, device=x.device)
    assert out.dtype == x.dtype
    _lib.check(lib.mv_upsample2x(_ptr(x), _ptr(out), B, h, w, C, 1 if x.dtype == torch.float16 else 0, _stream()),
               "mv_upsample2x")
    return out


def tokens_to_map_bwd(dmap, batch, n_tok, prefix, grid, out=None):
    """adjoint of tokens_to_map: d_map NHWC bf16 -> d_tokens bf16 [B*n_tok, D] (mv_tokens_to_map_bwd)."""
    lib = _lib_for(dmap)
    assert dmap.dtype == torch.bfloat16 and dmap.is_contiguous()
    B, t, _, D = dmap.shape
    if out is None:
        out = torch.empty((batch * n_tok, D), dtype=torch.bfloat16, device=dmap.device)
    _lib.check(lib.mv_tokens_to_map_bwd(_ptr(dmap), _ptr(out), out.stride(0), batch, n_tok, prefix, grid, t, D, _stream()),
               "mv_tokens_to_map_bwd")
    return out


def attn_bwd(qkv, out, dout, lse, batch, n_tok, heads, *, dqkv=None, dsum=None, scale=None):
    """dq | dk | dv from dout and the saved forward tensors (mv_attn_bwd)."""
    lib = _lib_for(qkv)
    D = heads * 64
    if dqkv is None:
        dqkv = torch.empty((batch * n_tok, 3 * D), dtype=torch.bfloat16, device=qkv.device)
    if dsum is None:
        dsum = torch.empty((batch, heads, n_tok), dtype=torch.float32, device=qkv.device)
    if scale is None:
        scale = 64 ** -0.5
    _lib.check(lib.mv_attn_bwd(_ptr(qkv), qkv.stride(0), _ptr(out), out.stride(0), _ptr(dout), dout.stride(0), _ptr(lse),
                               _ptr(dsum), _ptr(dqkv), dqkv.stride(0), batch, n_tok, heads, float(scale), _stream()),
               "mv_attn_bwd")
    return dqkv


def lora_grads(xn_ext, dqkv_ext, D, alpha, dAq, dAv, dBq, dBv, workspace=None):
    lib = _lib_for(xn_ext)
    M = xn_ext.shape[0]
    need = int(lib.mv_lora_grads_workspace_bytes(M, D))
    if workspace is None:
        workspace = torch.empty(need, dtype=torch.uint8, device=xn_ext.device)
    for t in (dAq, dAv, dBq, dBv):
        assert t.dtype == torch.float32 and t.is_contiguous()
    _lib.check(lib.mv_lora_grads(_ptr(xn_ext), xn_ext.stride(0), _ptr(dqkv_ext), dqkv_ext.stride(0), M, D, float(alpha),
                                 _ptr(dAq), _ptr(dAv), _ptr(dBq), _ptr(dBv), _ptr(workspace), workspace.numel(), _stream()),
               "mv_lora_grads")
    return workspace


LOSS_WMSE, LOSS_MAE, LOSS_L1L2 = 0, 1, 2


def loss_fwd_bwd(pred, target, weights=None, mode=LOSS_WMSE, lambda_factor=1.0, grad=None, want_grad=True,
                 grad_scale=1.0, workspace=None, loss=None):
    """loss (fp32 scalar tensor) and d loss / d pred in one pass (mv_loss_fwd_bwd); pred/target NCHW fp32."""
    lib = _lib_for(pred)
    assert pred.dtype == torch.float32 and target.dtype == torch.float32 and pred.is_contiguous() and target.is_contiguous()
    B, C, H, W = pred.shape
    if want_grad and grad is None:
        grad = torch.empty_like(pred)
    nws = int(lib.mv_loss_workspace_floats(B, C, H * W))
    if workspace is None:
        workspace = torch.empty(nws, dtype=torch.float32, device=pred.device)
    if loss is None:
        loss = torch.empty(1, dtype=torch.float32, device=pred.device)
    _lib.check(lib.mv_loss_fwd_bwd(_ptr(pred), _ptr(target), _ptr(grad if want_grad else None), _ptr(weights), B, C, H * W,
                                   mode, float(lambda_factor), float(grad_scale), _ptr(loss), _ptr(workspace),
                                   workspace.numel(), _stream()), "mv_loss_fwd_bwd")
    return loss, grad


def grad_norm(flat_grads, max_norm, norm_out=None, workspace=None):
    lib = _lib_for(flat_grads)
    if norm_out is None:
        norm_out = torch.empty(2, dtype=torch.float32, device=flat_grads.device)
    if workspace is None:
        workspace = torch.empty(1024, dtype=torch.float32, device=flat_grads.device)
    _lib.check(lib.mv_grad_norm(_ptr(flat_grads), flat_grads.numel(), float(max_norm), _ptr(norm_out), _ptr(workspace),
                                _stream()), "mv_grad_norm")
    return norm_out


def adam_clip_step(params, grads, exp_avg, exp_avg_sq, norm_coef, step, lr, beta1=0.5, beta2=0.999, eps=1e-7, grad_mul=1.0):
    lib = _lib_for(params)
    _lib.check(lib.mv_adam_clip_step(_ptr(params), _ptr(grads), _ptr(exp_avg), _ptr(exp_avg_sq), params.numel(),
                                     _ptr(norm_coef), float(grad_mul), float(lr), float(beta1), float(beta2), float(eps),
                                     int(step), _stream()), "mv_adam_clip_step")


def bn_finalize(colstats, count, gamma, beta, running_mean, running_var, pre_bias=None, momentum=0.1, eps=1e-5, out=None):
    """batch statistics -> (scale, shift, mean, rstd) + running-stat update (mv_bn_finalize)."""
    lib = _lib_for(colstats)
    C = gamma.numel()
    if out is None:
        out = torch.empty((4, C), dtype=torch.float32, device=colstats.device)
    for t in (colstats, gamma, beta, running_mean, running_var, pre_bias, out):  # raw pointers below: fp32, dense, same device
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and t.device == colstats.device)
    _lib.check(lib.mv_bn_finalize(_ptr(colstats), float(count), _ptr(gamma), _ptr(beta), _ptr(pre_bias), _ptr(running_mean),
                                  _ptr(running_var), float(momentum), float(eps), C, _ptr(out[0]), _ptr(out[1]), _ptr(out[2]),
                                  _ptr(out[3]), _stream()), "mv_bn_finalize")
    return out


def gram32(f, out=None, zero=True):
    """out (fp32 [40, 32], zeroed here): rows 0..31 = f^T f, row 32 = 1^T f for f bf16 [M, 32] (mv_gram32)."""
    lib = _lib_for(f)
    _rowmajor(f, "f")
    assert f.dtype in _H16 and f.shape[1] == 32
    if out is None:
        out = torch.zeros((40, 32), dtype=torch.float32, device=f.device)
    elif zero:
        out.zero_()
    _lib.check(lib.mv_gram32(_ptr(f), f.stride(0), f.shape[0], 1 if f.dtype == torch.float16 else 0, _ptr(out), _stream()),
               "mv_gram32")
    return out


def heads_bn_from_gram(gram, count, w1, b1, gamma, beta, running_mean, running_var, momentum=0.1, eps=1e-5, out=None,
                       w1_fmt=0):
    """closed-form BatchNorm batch statistics of W1 f + b1 from the moments of f (mv_heads_bn_from_gram)."""
    lib = _lib_for(gram)
    C = gamma.numel()
    assert w1.dtype == torch.float32 and w1.is_contiguous() and w1.shape == (C, 32)
    for t in (gram, b1, gamma, beta, running_mean, running_var):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and t.device == gram.device)
    if out is None:
        out = torch.empty((4, C), dtype=torch.float32, device=gram.device)
    _lib.check(lib.mv_heads_bn_from_gram(_ptr(gram), float(count), _ptr(w1), _ptr(b1), _ptr(gamma), _ptr(beta),
                                         _ptr(running_mean), _ptr(running_var), float(momentum), float(eps), C,
                                         _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(out[3]), int(w1_fmt), _stream()),
               "mv_heads_bn_from_gram")
    return out


def heads_bwd_algebra(E, FF, w1, b1, gamma, w2, fin, count, n_units, dw1, dgamma, dbeta, dw2, ca_t, mx_n, mx_scale, kshift):
    """closed-form gate-MLP / BatchNorm backward of the 16 heads (mv_heads_bwd_algebra)."""
    lib = _lib_for(E)
    for t in (E, FF, w1, b1, gamma, w2, fin, dw1, dgamma, dbeta, dw2, mx_scale, kshift):
        assert t.dtype == torch.float32 and t.is_contiguous()
    assert E.shape == (40, 256) and FF.shape == (40, 32) and fin.shape == (4, 256) and w1.shape == (256, 32)
    assert ca_t.dtype == torch.bfloat16 and ca_t.shape == (32, 256) and ca_t.is_contiguous()
    assert mx_n.dtype in _H16 and mx_n.shape == (32, 64) and mx_n.is_contiguous()
    _lib.check(lib.mv_heads_bwd_algebra(_ptr(E), _ptr(FF), _ptr(w1), _ptr(b1), _ptr(gamma), _ptr(w2), _ptr(fin), float(count),
                                        int(n_units), _ptr(dw1), _ptr(dgamma), _ptr(dbeta), _ptr(dw2), _ptr(ca_t), _ptr(mx_n),
                                        1 if mx_n.dtype == torch.float16 else 0, _ptr(mx_scale), _ptr(kshift), _stream()),
               "mv_heads_bwd_algebra")


def lora_refresh(lora_flat, ptrs, depth, D, alpha, ldw, ldb):
    lib = _lib_for(lora_flat)
    assert lora_flat.dtype == torch.float32 and lora_flat.is_contiguous() and lora_flat.numel() >= depth * 32 * D
    assert ptrs.dtype == torch.int64 and ptrs.is_contiguous() and ptrs.numel() == depth * 4
    _lib.check(lib.mv_lora_refresh(_ptr(lora_flat), _ptr(ptrs), depth, D, float(alpha), int(ldw), int(ldb), _stream()),
               "mv_lora_refresh")


def gather_cast(src, idx, dst):
    """dst[i] = convert(src[idx[i]]) (idx -1 -> 0, -2 -> untouched); dst fp32 / bf16 / fp16 (mv_gather_cast)."""
    lib = _lib_for(src)
    assert src.dtype == torch.float32 and src.is_contiguous() and idx.dtype == torch.int32 and idx.is_contiguous()
    assert dst.is_contiguous() and dst.numel() == idx.numel()
    mode = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[dst.dtype]
    _lib.check(lib.mv_gather_cast(_ptr(src), _ptr(idx), _ptr(dst), idx.numel(), mode, _stream()), "mv_gather_cast")


def add_i64(t, inc=1):
    lib = _lib_for(t)
    assert t.dtype == torch.int64 and t.is_contiguous()
    _lib.check(lib.mv_add_i64(_ptr(t), t.numel(), int(inc), _stream()), "mv_add_i64")


def memset(t, value=0):
    lib = _lib_for(t)
    assert t.is_contiguous()
    _lib.check(lib.mv_memset_async(_ptr(t), int(value), t.numel() * t.element_size(), _stream()), "mv_memset_async")


def adam_schedule(step, base_lr, total_steps, warmup_steps, beta1, beta2, hyper):
    lib = _lib_for(step)
    assert step.dtype == torch.int64 and hyper.dtype == torch.float32 and hyper.numel() >= 4
    _lib.check(lib.mv_adam_schedule(_ptr(step), float(base_lr), int(total_steps), int(warmup_steps), float(beta1),
                                    float(beta2), _ptr(hyper), _stream()), "mv_adam_schedule")


def adam_clip_step_dev(params, grads, exp_avg, exp_avg_sq, norm_coef, hyper, beta1=0.5, beta2=0.999, eps=1e-7, grad_mul=1.0):
    lib = _lib_for(params)
    _lib.check(lib.mv_adam_clip_step_dev(_ptr(params), _ptr(grads), _ptr(exp_avg), _ptr(exp_avg_sq), params.numel(),
                                         _ptr(norm_coef), float(grad_mul), _ptr(hyper), float(beta1), float(beta2),
                                         float(eps), _stream()), "mv_adam_clip_step_dev")


def bn_relu_apply(z, scale, shift, out=None):
    lib = _lib_for(z)
    M, C = z.shape
    assert z.is_contiguous()
    if out is None:
        out = torch.empty((M, C), dtype=torch.bfloat16, device=z.device)
    assert out.dtype in _H16 and out.is_contiguous()
    _lib.check(lib.mv_bn_relu_apply(_ptr(z), 1 if z.dtype == torch.float32 else 0, _ptr(scale), _ptr(shift), _ptr(out),
                                    1 if out.dtype == torch.float16 else 0, M, C, _stream()), "mv_bn_relu_apply")
    return out


def bn_relu_bwd(dy, y, z, mean, rstd, gamma, sums=None, dz=None):
    """dz = BN'(dy * [y>0]); returns (dz, sums) with sums[0] = dbeta, sums[1] = dgamma (mv_bn_relu_bwd)."""
    lib = _lib_for(dy)
    M, C = z.shape
    assert y.is_contiguous() and z.is_contiguous() and dy.stride(1) == 1
    if sums is None:
        sums = torch.empty((2, C), dtype=torch.float32, device=z.device)
    if dz is None:
        dz = torch.empty((M, C), dtype=torch.bfloat16, device=z.device)
    _lib.check(lib.mv_bn_relu_bwd(_ptr(dy), dy.stride(0), _ptr(y), _ptr(z), 1 if z.dtype == torch.float32 else 0, _ptr(mean),
                                  _ptr(rstd), _ptr(gamma), _ptr(sums), _ptr(dz), M, C, _stream()), "mv_bn_relu_bwd")
    return dz, sums


def transpose_bf16(x, ones_row=False, out=None):
    """[M, C] (strided rows; bf16, or fp16 converted on the way) -> bf16 [C (+1 ones row, padded to 8 rows), ld >= M]
    K-major copy (mv_transpose_bf16)."""
    lib = _lib_for(x)
    M, C = x.shape
    ld = (M + 7) // 8 * 8
    rows = C + (8 if ones_row else 0)
    if out is None:
        out = torch.zeros((rows, ld), dtype=torch.bfloat16, device=x.device) if ones_row else \
            torch.empty((rows, ld), dtype=torch.bfloat16, device=x.device)
    assert x.dtype in _H16 and out.dtype == torch.bfloat16
    _lib.check(lib.mv_transpose_bf16(_ptr(x), x.stride(0), _ptr(out), out.stride(0), M, C, 1 if ones_row else 0,
                                     1 if x.dtype == torch.float16 else 0, _stream()), "mv_transpose_bf16")
    return out


def f16_to_bf16(x, out=None):
    """bf16 twin of a contiguous fp16 tensor (mv_f16_to_bf16)."""
    lib = _lib_for(x)
    assert x.dtype == torch.float16 and x.is_contiguous() and x.numel() % 8 == 0
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    assert out.dtype == torch.bfloat16 and out.is_contiguous() and out.numel() == x.numel()
    _lib.check(lib.mv_f16_to_bf16(_ptr(x), _ptr(out), x.numel(), _stream()), "mv_f16_to_bf16")
    return out


def upsample2x_bwd(dup, out=None):
    """dup: [B, 2h, 2w, C] view (channel slice allowed) -> [B, h, w, C] (mv_upsample2x_bwd)."""
    lib = _lib_for(dup)
    B, H2, W2, C = dup.shape
    assert dup.stride(3) == 1 and dup.stride(1) == W2 * dup.stride(2) and dup.stride(0) == H2 * dup.stride(1)
    if out is None:
        out = torch.empty((B, H2 // 2, W2 // 2, C), dtype=torch.bfloat16, device=dup.device)
    _lib.check(lib.mv_upsample2x_bwd(_ptr(dup), dup.stride(2), _ptr(out), B, H2 // 2, W2 // 2, C, _stream()),
               "mv_upsample2x_bwd")
    return out


def zero_insert2x(dz, out=None):
    lib = _lib_for(dz)
    B, h, w, C = dz.shape
    assert dz.is_contiguous()
    if out is None:
        out = torch.empty((B, 2 * h, 2 * w, C), dtype=torch.bfloat16, device=dz.device)
    _lib.check(lib.mv_zero_insert2x(_ptr(dz), _ptr(out), B, h, w, C, _stream()), "mv_zero_insert2x")
    return out


def add_bf16(a, b, out=None):
    """out = a + b for [M, C] bf16 matrices with arbitrary row pitch (mv_add_bf16)."""
    lib = _lib_for(a)
    M, C = a.shape
    if out is None:
        out = torch.empty((M, C), dtype=torch.bfloat16, device=a.device)
    _lib.check(lib.mv_add_bf16(_ptr(a), a.stride(0), _ptr(b), b.stride(0), _ptr(out), M, C, _stream()), "mv_add_bf16")
    return out


def heads_ds(dpred, pred, dbias, out=None):
    lib = _lib_for(dpred)
    B, Hh, H, W = pred.shape
    assert dpred.is_contiguous() and pred.is_contiguous() and dpred.dtype == torch.float32 and pred.dtype == torch.float32
    if out is None:
        out = torch.empty((B * H * W, 16), dtype=torch.bfloat16, device=pred.device)
    _lib.check(lib.mv_heads_ds(_ptr(dpred), _ptr(pred), _ptr(out), _ptr(dbias), B, Hh, H * W, _stream()), "mv_heads_ds")
    return out


def heads_bwd_stencil(t, ds, gate, B, H, W, db2, dt=None, du=None):
    lib = _lib_for(t)
    M = B * H * W
    assert t.shape == (M, 144) and t.is_contiguous() and ds.shape == (M, 16) and gate.shape == (M, 16) and gate.is_contiguous()
    if dt is None:
        dt = torch.empty((M, 144), dtype=torch.bfloat16, device=t.device)
    if du is None:
        du = torch.empty((M, 16), dtype=torch.bfloat16, device=t.device)
    _lib.check(lib.mv_heads_bwd_stencil(_ptr(t), _ptr(ds), _ptr(gate), _ptr(dt), _ptr(du), _ptr(db2), B, H, W, _stream()),
               "mv_heads_bwd_stencil")
    return dt, du


def cell_means(pred, target, nuclei, cap=1024, return_counts=False):
    """Per-nucleus mean intensities (mv_cell_means + mv_cell_means_pack): pred / target fp32 [B, C, H, W] (target may be
    None), nuclei integer labels [B, H, W] or [B, 1, H, W], 0 = background.  Returns (pred_means [n, C], target_means
    [n, C] or None, cell_ids [n] int64, n_unique [B] int32) with the rows of image 0 first and ids ascending per image —
    the layout of MeanCellExtrator.extract_mean (src/utils.py:49-121).  One host sync (the row count)."""
    lib = _lib_for(pred)
    assert pred.dtype == torch.float32 and pred.is_contiguous() and pred.dim() == 4
    B, C, H, W = pred.shape
    if target is not None:
        assert target.dtype == torch.float32 and target.is_contiguous() and target.shape == pred.shape
    nuclei = nuclei.reshape(B, H * W)
    if nuclei.dtype not in (torch.int32, torch.int64):
        nuclei = nuclei.long()
    nuclei = nuclei.contiguous()
    dev = pred.device
    while True:
        mp = torch.empty((B, cap, C), dtype=torch.float32, device=dev)
        mt = torch.empty((B, cap, C), dtype=torch.float32, device=dev) if target is not None else None
        ids = torch.empty((B, cap), dtype=torch.int64, device=dev)
        cnt = torch.empty((B, cap), dtype=torch.float32, device=dev)
        nu = torch.zeros(B + 1, dtype=torch.int32, device=dev)  # [B] counts + overflow flag
        wsb = int(lib.mv_cell_means_workspace_bytes(B, C, cap))  # 0 while the tables fit in shared memory
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev) if wsb else None
        _lib.check(lib.mv_cell_means(_ptr(pred), _ptr(target), _ptr(nuclei), nuclei.element_size(), B, C, H * W, cap, _ptr(mp),
                                     _ptr(mt), _ptr(ids), _ptr(cnt), _ptr(nu), ctypes.c_void_p(nu.data_ptr() + 4 * B), _ptr(ws),
                                     wsb, _stream()), "mv_cell_means")
        host = nu.cpu()
        if int(host[B]) == 0:
            break
        if cap >= H * W:
            raise _lib.MipheiB200Error("mv_cell_means: more distinct labels than pixels?")
        cap *= 2  # an image holds more nuclei than rows: retry with a larger table (global-memory tables beyond ~1.3 k)
    n = int(host[:B].sum())
    op = torch.empty((n, C), dtype=torch.float32, device=dev)
    ot = torch.empty((n, C), dtype=torch.float32, device=dev) if target is not None else None
    oi = torch.empty((n,), dtype=torch.int64, device=dev)
    oc = torch.empty((n,), dtype=torch.float32, device=dev)
    if n > 0:
        _lib.check(lib.mv_cell_means_pack(_ptr(mp), _ptr(mt), _ptr(ids), _ptr(cnt), _ptr(nu), B, C, cap, _ptr(op), _ptr(ot),
                                          _ptr(oi), _ptr(oc), _stream()), "mv_cell_means_pack")
    if return_counts:
        return op, ot, oi, nu[:B], oc
    return op, ot, oi, nu[:B]


def cell_means_bwd(dmeans, ids, counts, n_unique, nuclei, B, C, H, W):
    """d map [B, C, H, W] (fp32) from d means [n, C]: mv_cell_means_bwd."""
    lib = _lib_for(dmeans)
    dmeans = dmeans.float().contiguous()
    nuclei = nuclei.reshape(B, H * W)
    if nuclei.dtype not in (torch.int32, torch.int64):
        nuclei = nuclei.long()
    nuclei = nuclei.contiguous()
    out = torch.empty((B, C, H, W), dtype=torch.float32, device=dmeans.device)
    _lib.check(lib.mv_cell_means_bwd(_ptr(dmeans), _ptr(ids), _ptr(counts), _ptr(n_unique), _ptr(nuclei), nuclei.element_size(),
                                     B, C, H * W, _ptr(out), _stream()), "mv_cell_means_bwd")
    return out


# ---------------------------------------------------------------------------------------------- whole-slide plumbing
def thumb_std_hist(thumb):
    """uint8 [H, W, C] (C = 1..4) -> (np.uint8(thumb.std(-1)) as uint8 [H, W], histogram int32 [256]) (mv_thumb_std_hist)."""
    lib = _lib_for(thumb)
    assert thumb.dtype == torch.uint8 and thumb.dim() == 3 and thumb.is_contiguous() and 1 <= thumb.shape[2] <= 4
    H, W, C = thumb.shape
    std = torch.empty((H, W), dtype=torch.uint8, device=thumb.device)
    hist = torch.empty(256, dtype=torch.int32, device=thumb.device)
    _lib.check(lib.mv_thumb_std_hist(_ptr(thumb), H * W, C, _ptr(std), _ptr(hist), _stream()), "mv_thumb_std_hist")
    return std, hist


def otsu_threshold(hist, n_pix, out=None):
    """the cv2 THRESH_OTSU threshold of an 8-bit image from its histogram -> device int32 [1] (mv_otsu_threshold)."""
    lib = _lib_for(hist)
    assert hist.dtype == torch.int32 and hist.numel() == 256 and hist.is_contiguous()
    if out is None:
        out = torch.empty(1, dtype=torch.int32, device=hist.device)
    _lib.check(lib.mv_otsu_threshold(_ptr(hist), int(n_pix), _ptr(out), _stream()), "mv_otsu_threshold")
    return out


def tile_tissue(mask_u8, boxes, thresh=None, fixed_thresh=0):
    """counts int32 [n] of pixels > threshold inside each clipped box (x0, y0, x1, y1) of a uint8 [H, W] map (mv_tile_tissue)."""
    lib = _lib_for(mask_u8)
    assert mask_u8.dtype == torch.uint8 and mask_u8.dim() == 2 and mask_u8.is_contiguous()
    assert boxes.dtype == torch.int32 and boxes.dim() == 2 and boxes.shape[1] == 4 and boxes.is_contiguous()
    counts = torch.empty(boxes.shape[0], dtype=torch.int32, device=mask_u8.device)
    _lib.check(lib.mv_tile_tissue(_ptr(mask_u8), mask_u8.shape[1], _ptr(thresh), int(fixed_thresh), _ptr(boxes), boxes.shape[0],
                                  _ptr(counts), _stream()), "mv_tile_tissue")
    return counts


def stitch_tiles(tiles, xy, crop, keep, canvas_ptr, canvas_h, canvas_w, sequential=False):
    """insert the [crop, crop + keep)^2 window of every uint8 tile [B, C, S, S] at canvas position xy[b] (mv_stitch_tiles);
    canvas_ptr: device-addressable uint8 [C, canvas_h, canvas_w] (device tensor data_ptr or mapped pinned host memory)."""
    lib = _lib_for(tiles)
    assert tiles.dtype == torch.uint8 and tiles.dim() == 4 and tiles.is_contiguous() and tiles.shape[2] == tiles.shape[3]
    assert xy.dtype == torch.int32 and xy.shape == (tiles.shape[0], 2) and xy.is_contiguous() and xy.device == tiles.device
    B, C, S, _ = tiles.shape
    _lib.check(lib.mv_stitch_tiles(_ptr(tiles), _ptr(xy), B, C, S, int(crop), int(keep), ctypes.c_void_p(int(canvas_ptr)),
                                   int(canvas_h), int(canvas_w), 1 if sequential else 0, _stream()), "mv_stitch_tiles")
