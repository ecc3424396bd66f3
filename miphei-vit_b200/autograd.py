"""Training-mode execution: the encoder forward/backward run on the hand-written kernels inside one autograd.Function
(activations saved in a per-batch-size tape), so `loss.backward()` of the reference's training_step
(src/models.py:134-135) reaches the LoRA matrices; the frozen ViT weights get no gradient (dX-only backward).

The decoder in TRAINING mode (BatchNorm batch statistics + its backward) currently runs on PyTorch's CUDA ops — an
interim library path that DESIGN.md lists as not yet hand-written; the eval-mode decoder is fully hand-written.
"""
import torch
import torch.nn.functional as F

from . import ops

NUM_PREFIX = 5


class _TrainTape:
    """Per-block saved activations for one batch size (persistent buffers: stable TMA descriptors, no allocator churn)."""

    def __init__(self, eng, B):
        dev = eng.device
        D, H, N = eng.D, eng.H, eng.N
        M = B * N
        bf, f32 = torch.bfloat16, torch.float32
        e = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)  # noqa: E731
        self.B, self.M = B, M
        L = eng.depth
        self.x = [e((M, D), f32) for _ in range(L + 1)]      # residual stream at every block input (+ final)
        self.xmid = [e((M, D), f32) for _ in range(L)]       # after the attention branch
        self.xn_ext = [torch.zeros((M, D + 64), dtype=bf, device=dev) for _ in range(L)]
        self.qkv = [e((M, 3 * D), bf) for _ in range(L)]
        self.o = [e((M, D), bf) for _ in range(L)]
        self.lse = [e((B, eng.heads, N), f32) for _ in range(L)]
        self.h = [e((M, 2 * H), bf) for _ in range(L)]
        # transients shared by all blocks
        self.xn2 = e((M, D), bf)
        self.u = e((M, H), bf)
        self.dh = e((M, 2 * H), bf)
        self.dxn = e((M, D), bf)
        self.do = e((M, D), bf)
        self.dqkv_ext = torch.zeros((M, 3 * D + 64), dtype=bf, device=dev)
        self.dsum = e((B, eng.heads, N), f32)
        self.dx = [e((M, D), f32) for _ in range(2)]
        self.dxb = [e((M, D), bf) for _ in range(2)]
        self.dtok = e((M, D), bf)
        self.lora_ws = torch.empty(int(ops._lib.load().mv_lora_grads_workspace_bytes(M, D)), dtype=torch.uint8, device=dev)


def _pack_backward_weights(eng):
    """Transposed (K-major for the dX GEMMs) copies of the frozen weights, LayerScale folded in."""
    D = eng.D
    vit = eng.model.encoder.vit
    with torch.no_grad():
        for pb, blk in zip(eng.blocks, vit.blocks):
            q = blk.attn.qkv
            g1, g2 = blk.ls1.gamma.detach().float(), blk.ls2.gamma.detach().float()
            pb["w2_bwd"] = (blk.mlp.fc2.weight.detach().float() * g2[:, None]).t().contiguous().to(torch.bfloat16)
            pb["w1_bwd"] = blk.mlp.fc1.weight.detach().t().contiguous().to(torch.bfloat16)
            pb["wproj_bwd"] = (blk.attn.proj.weight.detach().float() * g1[:, None]).t().contiguous().to(torch.bfloat16)
            wb = torch.zeros((D, 3 * D + 64), dtype=torch.bfloat16, device=eng.device)
            wb[:, :3 * D] = q.qkv.weight.detach().t()
            pb["wqkv_bwd_ext"] = wb
            pb["bcat"] = torch.zeros((16, 3 * D), dtype=torch.bfloat16, device=eng.device)
    eng._bwd_packed = True


def _refresh_lora_backward(eng):
    D = eng.D
    with torch.no_grad():
        for pb in eng.blocks:
            lq, lv = pb["lora"]
            pb["wqkv_bwd_ext"][:, 3 * D:3 * D + 8] = lq.A.detach()
            pb["wqkv_bwd_ext"][:, 3 * D + 8:3 * D + 16] = lv.A.detach()
            pb["bcat"][:8, :D] = lq.alpha * lq.B.detach()
            pb["bcat"][8:, 2 * D:] = lv.alpha * lv.B.detach()


def encoder_forward_train(eng, x, tape):
    """Same kernel sequence as the eval forward, writing every block's activations into the tape."""
    B, N, D, g = tape.B, eng.N, eng.D, eng.g
    ws = eng._workspace(B)
    ws.x_in.copy_(x)
    ops.prep_input(ws.x_in, img=ws.img8, pm=ws.pm)
    x0 = tape.x[0]
    ops.fill_prefix(x0, eng.prefix, B, N)
    ops.gemm(ws.pm, eng.pe_w, shift=eng.pe_b, resid=eng.pos, out=x0, rows_per_group=g * g, group_stride=N,
             row_offset=NUM_PREFIX, resid_row_mod=True)
    for i, pb in enumerate(eng.blocks):
        xe = tape.xn_ext[i]
        ops.layernorm_fwd(tape.x[i], pb["n1w"], pb["n1b"], out=xe[:, :D])
        ops.gemm(xe[:, :D], pb["acat"], out=xe[:, D:D + 16])
        ops.gemm(xe[:, :D + 16], pb["wqkv_ext"][:, :D + 16], shift=pb["bqkv"], out=tape.qkv[i])
        ops.attn_fwd(tape.qkv[i], B, N, eng.heads, out=tape.o[i], lse=tape.lse[i])
        ops.gemm(tape.o[i], pb["wproj"], scale=pb["g1"], shift=pb["g1b"], resid=tape.x[i], out=tape.xmid[i])
        ops.layernorm_fwd(tape.xmid[i], pb["n2w"], pb["n2b"], out=tape.xn2)
        ops.gemm(tape.xn2, pb["w1"], mode=ops.GEMM_SWIGLU, shift=pb["b1"], out=tape.u, aux=tape.h[i])
        ops.gemm(tape.u, pb["w2"], scale=pb["g2"], shift=pb["g2b"], resid=tape.xmid[i], out=tape.x[i + 1])
    ops.layernorm_fwd(tape.x[-1], eng.nw, eng.nb, out=ws.tok)
    ops.tokens_to_map(ws.tok, B, N, NUM_PREFIX, g, eng.S // 16, out=ws.fmap)
    return ws.fmap


def encoder_backward(eng, tape, dmap, on_start=None, sink=False):
    """dmap: NHWC bf16 gradient of the resized feature map. Returns per-block (dA_q, dB_q, dA_v, dB_v)."""
    B, N, D = tape.B, eng.N, eng.D
    if on_start is not None:
        on_start()
    ops.tokens_to_map_bwd(dmap, B, N, NUM_PREFIX, eng.g, out=tape.dtok)
    cur = 0
    ops.layernorm_bwd(tape.x[-1], eng.nw, tape.dtok, out=tape.dx[cur], out_bf16=tape.dxb[cur])
    grads = [None] * eng.depth
    dq = tape.dqkv_ext
    for i in range(eng.depth - 1, -1, -1):
        pb = eng.blocks[i]
        dx, dxb = tape.dx[cur], tape.dxb[cur]
        nxt = cur ^ 1
        ops.gemm(dxb, pb["w2_bwd"], mode=ops.GEMM_SWIGLU_BWD, in2=tape.h[i], out=tape.dh)
        ops.gemm(tape.dh, pb["w1_bwd"], out=tape.dxn)
        ops.layernorm_bwd(tape.xmid[i], pb["n2w"], tape.dxn, dres=dx, out=tape.dx[nxt], out_bf16=tape.dxb[nxt])
        ops.gemm(tape.dxb[nxt], pb["wproj_bwd"], out=tape.do)
        ops.attn_bwd(tape.qkv[i], tape.o[i], tape.do, tape.lse[i], B, N, eng.heads, dqkv=dq, dsum=tape.dsum)
        # dT = [dQ | dK | dV] . [alpha B_q ; 0 ; alpha B_v]^T: the dK third multiplies zeros and is never loaded
        ops.gemm(dq[:, :3 * D], pb["bcat"], out=dq[:, 3 * D:3 * D + 16], kskip=(D, 2 * D))
        lq, lv = pb["lora"]
        if sink:  # trainer mode: the kernels write straight into the flat gradient buffer (p.grad are views of it)
            gAq, gBq, gAv, gBv = lq.A.grad, lq.B.grad, lv.A.grad, lv.B.grad
        else:
            gAq, gBq = torch.empty_like(lq.A), torch.empty_like(lq.B)
            gAv, gBv = torch.empty_like(lv.A), torch.empty_like(lv.B)
        ops.lora_grads(tape.xn_ext[i], dq, D, lq.alpha, gAq, gAv, gBq, gBv, workspace=tape.lora_ws)
        grads[i] = (gAq, gBq, gAv, gBv)
        if i > 0:
            ops.gemm(dq[:, :3 * D + 16], pb["wqkv_bwd_ext"][:, :3 * D + 16], out=tape.dxn)
            ops.layernorm_bwd(tape.x[i], pb["n1w"], tape.dxn, dres=tape.dx[nxt], out=tape.dx[cur], out_bf16=tape.dxb[cur])
        # after this block `cur` holds dx w.r.t. the block input again
    return grads


class _MipheiFn(torch.autograd.Function):
    """Whole generator in training mode: encoder + decoder forward on the kernels, one backward for every trainable
    parameter (LoRA A/B of each block + the decoder)."""

    @staticmethod
    def forward(ctx, eng, x, *params):
        tape = eng._train_tape(x.shape[0])
        fmap = encoder_forward_train(eng, x, tape)
        ws = eng._workspace(x.shape[0])
        pred = eng.decoder_train.forward(fmap, ws.img8)
        ctx.eng, ctx.tape, ctx.params = eng, tape, params
        return pred

    @staticmethod
    def backward(ctx, dpred):
        eng, tape, params = ctx.eng, ctx.tape, ctx.params
        grads, dfmap = eng.decoder_train.backward(dpred.float())
        sink = eng.direct_grad_sink  # trainer mode: write straight into the flat gradient buffer
        if sink:
            # ~150 decoder gradients -> their views of the flat gradient buffer in one multi-tensor copy
            dst = [p.grad for p in grads]
            src = [g.reshape(p.grad.shape) if g.shape != p.grad.shape else g for p, g in grads.items()]
            src = [g if g.dtype == d.dtype else g.to(d.dtype) for g, d in zip(src, dst)]
            torch._foreach_copy_(dst, src)
            grads = {}
        sink_lora = sink and all(q.grad is not None and q.grad.is_contiguous() for pb in eng.blocks[:1] for l in pb["lora"]
                                 for q in (l.A, l.B))
        lgrads = encoder_backward(eng, tape, dfmap, on_start=eng.on_encoder_backward_start, sink=sink_lora)
        if not sink_lora:
            for pb, (gAq, gBq, gAv, gBv) in zip(eng.blocks, lgrads):
                lq, lv = pb["lora"]
                for p, g in ((lq.A, gAq), (lq.B, gBq), (lv.A, gAv), (lv.B, gBv)):
                    if sink:
                        p.grad.copy_(g)
                    else:
                        grads[p] = g
        return (None, None) + tuple(grads.get(p) for p in params)


def decoder_forward_torch(dec, feat, images):
    """Detail_Capture.forward on PyTorch CUDA ops — kept ONLY as a debugging cross-check (tools/), never on the product path."""
    with torch.autocast("cuda", enabled=False):
        details = [images]
        x = images
        for m in dec.convstream.convs:
            x = F.relu(m.bn(m.conv(x)))
            details.append(x)
        f = feat
        for i, m in enumerate(dec.fusion_blks):
            up = F.interpolate(f, scale_factor=2, mode="bilinear", align_corners=False)
            f = torch.cat([details[3 - i], up], dim=1)
            f = F.relu(m.conv.bn(m.conv.conv(f)))
        outs = []
        for h in range(dec.num_heads):
            head = getattr(dec, "segmentation_head_%d" % h)
            gate = head[0].psi(f)
            outs.append(torch.tanh(head[1](f * gate)))
        return torch.cat(outs, dim=1)


def miphei_train_forward(eng, x):
    if not eng.model.training:
        raise NotImplementedError("gradients through the eval-mode (running-statistics BatchNorm) generator are not "
                                  "implemented; call model.train() for training or torch.no_grad() for inference")
    if not getattr(eng, "_bwd_packed", False):
        _pack_backward_weights(eng)
        eng._lora_bwd_versions = None
    ver = tuple(p._version for pb in eng.blocks for l in pb["lora"] for p in (l.A, l.B))
    if eng._lora_bwd_versions != ver:
        _refresh_lora_backward(eng)
        eng._lora_bwd_versions = ver
    if eng.decoder_train is None:
        from .decoder_train import DecoderTrain
        eng.decoder_train = DecoderTrain(eng)
        eng._dec_train_versions = None
    if eng._dec_train_versions != eng._lora_versions or eng._dec_train_versions is None:
        eng.decoder_train.pack()
        eng._dec_train_versions = eng._lora_versions
    xf = x.float().contiguous()
    if torch.is_grad_enabled():
        pred = _MipheiFn.apply(eng, xf, *eng._trainables)
    else:  # train-mode forward without autograd (BatchNorm batch statistics, running stats updated)
        tape = eng._train_tape(xf.shape[0])
        fmap = encoder_forward_train(eng, xf, tape)
        pred = eng.decoder_train.forward(fmap, eng._workspace(xf.shape[0]).img8)
    out_dtype = eng._out_dtype(x)
    return pred if pred.dtype == out_dtype else pred.to(out_dtype)
