"""Data-parallel trainer for the MIPHEI-ViT generator: the reference's manual-optimisation step
(ModelModule.training_step, src/models.py:87-139: forward, loss, backward, clip_gradients(1.0), Adam(0.5, 0.999, 1e-7),
per-step LambdaLR from pix2pix_lr_scheduler, src/utils.py:217-230) with

  * every trainable parameter (LoRA A/B of the 40 blocks + the decoder, 6.7 M values) living in ONE flat fp32 buffer,
    decoder segment first, so clip + Adam is one fused kernel and the gradient exchange is two contiguous buckets;
  * one process per GPU (torch.distributed / NCCL): the decoder bucket is all-reduced on a side stream as soon as the
    decoder backward has finished — it overlaps the whole encoder backward — the LoRA bucket at the end;
  * per-replica BatchNorm statistics (the reference's semantics at the per-GPU batch; it has no multi-GPU rule).
"""
import math

import torch
import torch.distributed as dist

from . import ops


def lr_lambda(step, total_steps, warmup_steps=400):
    """pix2pix_lr_scheduler(total, 400, total // 2) as configure_optimizers builds it (src/models.py:363-369)."""
    half = total_steps // 2
    if step < warmup_steps:
        return step / warmup_steps
    if step < half:
        return 1.0
    return max(0.0, 1.0 - (step - half) / (total_steps - half))


def shard_tiles(n_tiles, rank, world):
    """Round-robin tile assignment of a whole-slide sweep (independent units, no collective)."""
    return range(rank, n_tiles, world)


class FlatParams:
    """Every trainable parameter as a view of ONE flat fp32 buffer, decoder segment first, with a matching flat gradient
    buffer (p.grad views) and two contiguous all-reduce buckets. Device-agnostic (tested on CPU with gloo)."""

    def __init__(self, model, device=None):
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        dec = [(n, p) for n, p in named if n.startswith("decoder.")]
        lora = [(n, p) for n, p in named if not n.startswith("decoder.")]
        self.order = dec + lora
        dev = device if device is not None else named[0][1].device
        pad = lambda n: (n + 3) // 4 * 4  # noqa: E731  (16-byte aligned segments)
        self.n_dec = sum(pad(p.numel()) for _, p in dec)
        self.total = self.n_dec + sum(pad(p.numel()) for _, p in lora)
        self.flat = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.gflat = torch.zeros(self.total, dtype=torch.float32, device=dev)
        from .packing import decoder_layout
        lay, n_dec = decoder_layout(model)
        assert n_dec == self.n_dec and [n for n, _, _, _ in lay] == [n for n, _ in dec]
        off = 0
        with torch.no_grad():
            for _, p in self.order:
                n = p.numel()
                self.flat[off:off + n].copy_(p.detach().float().flatten())
                p.data = self.flat[off:off + n].view(p.shape)
                p.grad = self.gflat[off:off + n].view(p.shape)
                off += pad(n)

    def bucket(self, i):
        return self.gflat[:self.n_dec] if i == 0 else self.gflat[self.n_dec:]

    def allreduce_bucket(self, i, async_op=True):
        """Average one bucket over the ranks (NCCL supports AVG natively; gloo gets SUM + scale)."""
        b = self.bucket(i)
        if dist.get_backend() == "nccl":
            return dist.all_reduce(b, op=dist.ReduceOp.AVG, async_op=async_op)
        w = dist.all_reduce(b, op=dist.ReduceOp.SUM, async_op=False)
        b.div_(dist.get_world_size())
        return w


class Trainer:
    """step(x, y) = the reference's manual-optimisation training_step on the fused kernels.

    Fast path (default): the step does not go through autograd at all — forward, loss, backward, clip and Adam are a fixed
    kernel sequence over persistent buffers, captured once per batch size in CUDA graphs and replayed (no framework op and
    no launch gap inside a step; learning rate and step count live on the device, mv_adam_schedule).  With several ranks the
    sequence is cut into four graphs around the NCCL calls: [forward + decoder backward] -> all-reduce(decoder bucket)
    overlapping [encoder backward, upper half] -> all-reduce(LoRA, upper blocks) overlapping [encoder backward, lower half]
    -> all-reduce(LoRA, lower blocks) -> [clip + Adam].  step_autograd() is the same step through `model(x)` /
    `pred.backward()` (the drop-in route a LightningModule takes)."""

    def __init__(self, model, marker_weights=None, base_lr=None, batch_size=None, total_steps=10000, warmup_steps=400,
                 lambda_factor=50.0, loss_mode=ops.LOSS_WMSE, betas=(0.5, 0.999), eps=1e-7, max_norm=1.0, use_graph=True):
        self.model = model
        self.eng = model.engine
        dev = next(model.parameters()).device
        self.device = dev
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.base_lr = base_lr if base_lr is not None else 2e-4 * math.sqrt(batch_size or 16)  # src/train.py:163
        self.total_steps, self.warmup_steps = total_steps, warmup_steps
        self.lambda_factor, self.loss_mode = lambda_factor, loss_mode
        self.betas, self.eps, self.max_norm = betas, eps, max_norm
        self.marker_weights = marker_weights.to(dev).float().contiguous() if marker_weights is not None else None
        self.step_count = 0
        self.fp = FlatParams(model, dev)
        self.order, self.n_dec = self.fp.order, self.fp.n_dec
        self.flat, self.gflat = self.fp.flat, self.fp.gflat
        total = self.fp.total
        self.m = torch.zeros(total, dtype=torch.float32, device=dev)
        self.v = torch.zeros(total, dtype=torch.float32, device=dev)
        self.norm = torch.zeros(2, dtype=torch.float32, device=dev)
        self.norm_ws = torch.zeros(1024, dtype=torch.float32, device=dev)
        self.loss_buf = torch.zeros(1, dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)   # optimiser steps taken (device copy of step_count)
        self.hyper = torch.zeros(4, dtype=torch.float32, device=dev)    # written by mv_adam_schedule
        self.comm_stream = torch.cuda.Stream(device=dev) if self.world > 1 else None
        self._dec_work = None
        self.use_graph = use_graph
        self._graphs = {}
        self._io = {}
        self.eng.on_encoder_backward_start = self._decoder_grads_ready
        self.eng.invalidate()
        self.eng.direct_grad_sink = True
        self.eng.flat_params = (self.flat, self.gflat, self.n_dec)
        # the LoRA segment must be laid out block by block as (A_q, B_q, A_v, B_v): mv_lora_refresh reads it that way
        names = [n for n, _ in self.order if not n.startswith("decoder.")]
        for i in range(len(names) // 4):
            want = ["encoder.vit.blocks.%d.attn.qkv.%s" % (i, t) for t in ("lora_q.A", "lora_q.B", "lora_v.A", "lora_v.B")]
            assert names[4 * i:4 * i + 4] == want, names[4 * i:4 * i + 4]

    # called by the encoder's autograd node when it starts its backward: every decoder gradient is final
    def _decoder_grads_ready(self):
        if self.world == 1:
            return
        self.comm_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm_stream):
            self._dec_work = self.fp.allreduce_bucket(0, async_op=True)

    def current_lr(self):
        return self.base_lr * lr_lambda(self.step_count, self.total_steps, self.warmup_steps)

    # ------------------------------------------------------------------ the step through autograd (drop-in route)
    def step_autograd(self, x, y):
        """One optimisation step on device tensors x [B,3,S,S], y [B,C,S,S] through model(x) / backward()."""
        self.model.train()
        pred = self.model(x)
        p32 = pred.detach().float().contiguous()
        loss, dpred = ops.loss_fwd_bwd(p32, y, self.marker_weights, mode=self.loss_mode, lambda_factor=self.lambda_factor,
                                       loss=self.loss_buf)
        pred.backward(dpred.to(pred.dtype))
        if self.world > 1:
            if self._dec_work is not None:
                self._dec_work.wait()
                torch.cuda.current_stream().wait_stream(self.comm_stream)
                self._dec_work = None
            self.fp.allreduce_bucket(1, async_op=False)
        self._optimizer_stage()
        self._after_step()
        return loss

    def _optimizer_stage(self):
        ops.grad_norm(self.gflat, self.max_norm, norm_out=self.norm, workspace=self.norm_ws)
        ops.adam_schedule(self.step_dev, self.base_lr, self.total_steps, self.warmup_steps, self.betas[0], self.betas[1],
                          self.hyper)
        ops.adam_clip_step_dev(self.flat, self.gflat, self.m, self.v, self.norm, self.hyper, self.betas[0], self.betas[1],
                               self.eps)

    def _after_step(self):
        self.step_count += 1
        # parameters changed in place through the flat buffer, behind autograd's version counters
        self.eng.bump_weights()

    # ------------------------------------------------------------------ the fast path
    def _buffers(self, B):
        io = self._io.get(B)
        if io is None:
            eng = self.eng
            shp = (B, eng.heads_out, eng.S, eng.S)
            io = dict(y=torch.empty(shp, dtype=torch.float32, device=self.device),
                      dpred=torch.empty(shp, dtype=torch.float32, device=self.device),
                      loss_ws=torch.empty(int(ops._lib.load().mv_loss_workspace_floats(B, eng.heads_out, eng.S * eng.S)),
                                          dtype=torch.float32, device=self.device))
            self._io[B] = io
        return io

    def _stage_forward(self, tape, io):
        """operand refresh from the flat parameters, forward, loss + its gradient, decoder backward."""
        from .autograd import encoder_forward_train
        eng = self.eng
        eng.refresh_lora_operands()
        dt = eng.decoder_train
        dt.pack()
        tape.generation += 1
        fmap = encoder_forward_train(eng, tape)
        pred = dt.forward(fmap, tape.img)
        ops.loss_fwd_bwd(pred, io["y"], self.marker_weights, mode=self.loss_mode, lambda_factor=self.lambda_factor,
                         grad=io["dpred"], loss=self.loss_buf, workspace=io["loss_ws"])
        io["dfmap"] = dt.backward(io["dpred"])

    def _stage_encoder_bwd(self, tape, io, hi, lo, head):
        from .autograd import encoder_backward_blocks, encoder_backward_head
        eng = self.eng
        if head:
            encoder_backward_head(eng, tape, io["dfmap"])
        encoder_backward_blocks(eng, tape, hi, lo, lambda i: tuple(t.grad for l in eng.blocks[i]["lora"] for t in (l.A, l.B)))

    def _lora_slice(self, lo, hi):
        per = 32 * self.eng.D
        return self.gflat[self.n_dec + lo * per:self.n_dec + hi * per]

    def _allreduce(self, t):
        if dist.get_backend() == "nccl":
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=True)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t.div_(self.world)
        return None

    def _stages(self, tape, io):
        """[(callable, collective slice issued after it or None)]: one stage on a single GPU, four around the collectives."""
        L = self.eng.depth
        if self.world == 1:
            def whole():
                self._stage_forward(tape, io)
                self._stage_encoder_bwd(tape, io, L, 0, True)
                self._optimizer_stage()
            return [(whole, None)]
        half = L // 2
        return [(lambda: self._stage_forward(tape, io), self.gflat[:self.n_dec]),
                (lambda: self._stage_encoder_bwd(tape, io, L, half, True), self._lora_slice(half, L)),
                (lambda: self._stage_encoder_bwd(tape, io, half, 0, False), self._lora_slice(0, half)),
                (self._optimizer_stage, None)]

    def step(self, x, y):
        """One optimisation step; x [B,3,S,S] fp32, y [B,C,S,S] fp32 — device tensors or PINNED host tensors (copied with
        non-blocking H2D copies).  Returns the loss as a 1-element device tensor (a persistent buffer: read it before the
        next step); no host synchronisation."""
        from .autograd import prepare_training
        eng = self.eng
        self.model.train()
        if not eng._packed or not eng._bwd_packed or eng.decoder_train is None:
            eng._ensure_packed(train=True)
            prepare_training(eng)
        eng._check_input_shape(x)
        B = x.shape[0]
        tape = eng._train_tape(B)
        io = self._buffers(B)
        tape.x_in.copy_(x, non_blocking=True)
        io["y"].copy_(y, non_blocking=True)
        stages = self._stages(tape, io)
        graphs = self._graphs.get(B) if self.use_graph else None
        if self.use_graph and graphs is None:
            # first step at this batch size runs eagerly (buffer allocation, descriptor creation, kernel attributes);
            # the second one is captured
            if io.get("warm"):
                graphs = [None] * len(stages)
                self._graphs[B] = graphs
            io["warm"] = True
        cur = torch.cuda.current_stream() if self.world > 1 else None
        works = []
        for k, (fn, coll) in enumerate(stages):
            if k == len(stages) - 1 and works:  # optimiser stage: every bucket reduced
                for w in works:
                    if w is not None:
                        w.wait()
                cur.wait_stream(self.comm_stream)
            if graphs is None:
                fn()
            else:
                if graphs[k] is None:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        fn()
                    graphs[k] = g  # capture records, the replay below executes
                graphs[k].replay()
            if coll is not None:
                self.comm_stream.wait_stream(cur)
                with torch.cuda.stream(self.comm_stream):
                    works.append(self._allreduce(coll))
        self._after_step()
        return self.loss_buf
