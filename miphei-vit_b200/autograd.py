"""Training-mode execution: encoder and decoder forward / backward run on the hand-written kernels inside ONE
autograd.Function, so `loss.backward()` of the reference's training_step (src/models.py:134-135) reaches the LoRA matrices
and the decoder; the frozen ViT weights get no gradient (dX-only backward).

Activations live in a per-batch-size tape of persistent buffers (stable addresses: cached TMA descriptors, CUDA-graph
capture by trainer.Trainer).  The tape is single-use: every forward stamps a generation number, and a backward whose
generation is stale (a second forward at the same batch size ran in between, or backward is called twice) raises instead
of returning gradients of the wrong activations.
"""
import torch

from . import ops

NUM_PREFIX = 5


class _TrainTape:
    """Saved activations for one batch size. Inputs and encoder->decoder tensors are private to the tape (eval-mode
    inference at the same batch size uses the engine's own workspaces and never touches them)."""

    def __init__(self, eng, B):
        dev = eng.device
        D, H, N, S = eng.D, eng.H, eng.N, eng.S
        M = B * N
        bf, f32, f16 = torch.bfloat16, torch.float32, torch.float16
        e = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)  # noqa: E731
        self.B, self.M = B, M
        self.generation = 0
        L = eng.depth
        self.x_in = e((B, 3, S, S), f32)
        self.img = e((B, S, S, 8), f16)                      # decoder image input: fp16 (see decoder_train.py)
        self.pm = e((B * eng.g * eng.g, 592), bf)
        self.tok = e((M, D), f16)                            # final-norm tokens and the resized map feed the fp16 decoder
        self.fmap = e((B, S // 16, S // 16, D), f16)
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


def encoder_forward_train(eng, tape):
    """tape.x_in (fp32 NCHW) -> tape.fmap; the eval kernel sequence, writing every block's activations into the tape."""
    B, N, D, g = tape.B, eng.N, eng.D, eng.g
    ops.prep_input(tape.x_in, img=tape.img, pm=tape.pm)
    x0 = tape.x[0]
    ops.fill_prefix(x0, eng.prefix, B, N)
    ops.gemm(tape.pm, eng.pe_w, shift=eng.pe_b, resid=eng.pos, out=x0, rows_per_group=g * g, group_stride=N,
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
    ops.layernorm_fwd(tape.x[-1], eng.nw, eng.nb, out=tape.tok)
    ops.tokens_to_map(tape.tok, B, N, NUM_PREFIX, g, eng.S // 16, out=tape.fmap)
    return tape.fmap


def encoder_backward_head(eng, tape, dmap):
    """d(feature map) -> gradient of the residual stream after the last block (final LayerNorm + bicubic adjoint)."""
    ops.tokens_to_map_bwd(dmap, tape.B, eng.N, NUM_PREFIX, eng.g, out=tape.dtok)
    ops.layernorm_bwd(tape.x[-1], eng.nw, tape.dtok, out=tape.dx[0], out_bf16=tape.dxb[0])


def encoder_backward_blocks(eng, tape, hi, lo, lora_grads):
    """Blocks hi-1 .. lo (descending). The residual gradient enters and leaves in tape.dx[0] / dxb[0].
    lora_grads(i) -> (dA_q, dB_q, dA_v, dB_v) fp32 destination tensors of block i."""
    B, N, D = tape.B, eng.N, eng.D
    dq = tape.dqkv_ext
    for i in range(hi - 1, lo - 1, -1):
        pb = eng.blocks[i]
        dx, dxb = tape.dx[0], tape.dxb[0]
        ops.gemm(dxb, pb["w2_bwd"], mode=ops.GEMM_SWIGLU_BWD, in2=tape.h[i], out=tape.dh)
        ops.gemm(tape.dh, pb["w1_bwd"], out=tape.dxn)
        ops.layernorm_bwd(tape.xmid[i], pb["n2w"], tape.dxn, dres=dx, out=tape.dx[1], out_bf16=tape.dxb[1])
        ops.gemm(tape.dxb[1], pb["wproj_bwd"], out=tape.do)
        ops.attn_bwd(tape.qkv[i], tape.o[i], tape.do, tape.lse[i], B, N, eng.heads, dqkv=dq, dsum=tape.dsum)
        # dT = [dQ | dK | dV] . [alpha B_q ; 0 ; alpha B_v]^T: the dK third multiplies zeros and is never loaded
        ops.gemm(dq[:, :3 * D], pb["bcat"], out=dq[:, 3 * D:3 * D + 16], kskip=(D, 2 * D))
        gAq, gBq, gAv, gBv = lora_grads(i)
        ops.lora_grads(tape.xn_ext[i], dq, D, pb["lora"][0].alpha, gAq, gAv, gBq, gBv, workspace=tape.lora_ws)
        if i > 0:
            ops.gemm(dq[:, :3 * D + 16], pb["wqkv_bwd_ext"][:, :3 * D + 16], out=tape.dxn)
            ops.layernorm_bwd(tape.x[i], pb["n1w"], tape.dxn, dres=tape.dx[1], out=tape.dx[0], out_bf16=tape.dxb[0])


def prepare_training(eng):
    """Weight layouts of the training path (transposed frozen weights once; LoRA / decoder operands whenever the trainable
    parameters changed). Called before every train-mode forward."""
    if not getattr(eng, "_bwd_packed", False):
        _pack_backward_weights(eng)
        eng._lora_bwd_versions = None
    if eng.decoder_train is None:
        from .decoder_train import DecoderTrain
        eng.decoder_train = DecoderTrain(eng)
        eng._dec_train_versions = None
    ver = eng._trainable_version()
    if eng._lora_bwd_versions != ver:
        eng.refresh_lora_operands()
        eng._lora_bwd_versions = ver
    if eng._dec_train_versions != ver:
        eng.decoder_train.pack()
        eng._dec_train_versions = ver


class _MipheiFn(torch.autograd.Function):
    """Whole generator in training mode: encoder + decoder forward on the kernels, one backward for every trainable
    parameter (LoRA A/B of each block + the decoder)."""

    @staticmethod
    def forward(ctx, eng, x, *params):
        tape = eng._train_tape(x.shape[0])
        tape.generation += 1
        tape.x_in.copy_(x)
        fmap = encoder_forward_train(eng, tape)
        pred = eng.decoder_train.forward(fmap, tape.img)
        ctx.eng, ctx.tape, ctx.params = eng, tape, params
        ctx.gen = (tape.generation, eng.decoder_train.generation)
        # the prediction lives in a persistent buffer; hand autograd its own copy unless the trainer consumes it at once
        return pred if eng.direct_grad_sink else pred.clone()

    @staticmethod
    def backward(ctx, dpred):
        eng, tape, params = ctx.eng, ctx.tape, ctx.params
        dt = eng.decoder_train
        if dt is None or ctx.gen != (tape.generation, dt.generation):
            raise RuntimeError(
                "MIPHEI-ViT B200 generator: backward() of a stale forward — the activation tape of batch size %d was "
                "overwritten by a later train-mode forward (or this graph was already back-propagated). One forward, one "
                "backward per batch size; see INTEGRATION.md." % tape.B)
        tape.generation += 1  # single use: a second backward through the same graph raises
        dfmap = dt.backward(dpred.float().contiguous())
        sink = eng.direct_grad_sink  # trainer mode: the kernels write straight into the flat gradient buffer (p.grad views)
        if eng.on_encoder_backward_start is not None:
            eng.on_encoder_backward_start()
        got = {}
        if sink:
            dst = lambda i: tuple(t.grad for l in eng.blocks[i]["lora"] for t in (l.A, l.B))  # noqa: E731
        else:
            def dst(i):
                lq, lv = eng.blocks[i]["lora"]
                g = tuple(torch.empty_like(t, dtype=torch.float32) for t in (lq.A, lq.B, lv.A, lv.B))
                for t, gg in zip((lq.A, lq.B, lv.A, lv.B), g):
                    got[t] = gg
                return g
        encoder_backward_head(eng, tape, dfmap)
        encoder_backward_blocks(eng, tape, eng.depth, 0, dst)
        if sink:
            return (None, None) + (None,) * len(params)
        got.update(dt.grad_views())
        return (None, None) + tuple(got[p].to(p.dtype).clone() if p in got else None for p in params)


def miphei_train_forward(eng, x):
    if not eng.model.training:
        raise NotImplementedError("gradients through the eval-mode (running-statistics BatchNorm) generator are not "
                                  "implemented; call model.train() for training or torch.no_grad() for inference")
    prepare_training(eng)
    xf = x.float().contiguous()
    if torch.is_grad_enabled():
        pred = _MipheiFn.apply(eng, xf, *eng._trainables)
    else:  # train-mode forward without autograd (BatchNorm batch statistics, running stats updated)
        tape = eng._train_tape(xf.shape[0])
        tape.generation += 1
        tape.x_in.copy_(xf)
        fmap = encoder_forward_train(eng, tape)
        pred = eng.decoder_train.forward(fmap, tape.img).clone()
    eng.bump_weights()  # BatchNorm running statistics moved (kernel writes bypass autograd's version counters)
    out_dtype = eng._out_dtype(x)
    return pred if pred.dtype == out_dtype else pred.to(out_dtype)
