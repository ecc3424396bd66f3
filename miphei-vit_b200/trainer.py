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
    def __init__(self, model, marker_weights=None, base_lr=None, batch_size=None, total_steps=10000, warmup_steps=400,
                 lambda_factor=50.0, loss_mode=ops.LOSS_WMSE, betas=(0.5, 0.999), eps=1e-7, max_norm=1.0):
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
        self.comm_stream = torch.cuda.Stream(device=dev) if self.world > 1 else None
        self._dec_work = None
        self.eng.on_encoder_backward_start = self._decoder_grads_ready
        self.eng.invalidate()
        self.eng.direct_grad_sink = True

    # called by the encoder's autograd node when it starts its backward: every decoder gradient is final
    def _decoder_grads_ready(self):
        if self.world == 1:
            return
        self.comm_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm_stream):
            self._dec_work = self.fp.allreduce_bucket(0, async_op=True)

    def current_lr(self):
        return self.base_lr * lr_lambda(self.step_count, self.total_steps, self.warmup_steps)

    def step(self, x, y):
        """One optimisation step on device tensors x [B,3,S,S], y [B,C,S,S]; returns the loss tensor (no host sync)."""
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
        ops.grad_norm(self.gflat, self.max_norm, norm_out=self.norm, workspace=self.norm_ws)
        self.step_count += 1
        lr = self.base_lr * lr_lambda(self.step_count - 1, self.total_steps, self.warmup_steps)
        ops.adam_clip_step(self.flat, self.gflat, self.m, self.v, self.norm, self.step_count, lr, self.betas[0],
                           self.betas[1], self.eps)
        # parameters changed in place through the flat buffer, behind autograd's version counters
        self.eng.bump_weights()
        self.eng._lora_bwd_versions = None
        return loss


def bench_train(model, args, rank, world, dev):
    """BASELINE configs[2]: training step (fwd + bwd + loss + clip + Adam), batch 32 per GPU, weak scaling."""
    if getattr(args, "no_train", False):
        return None
    import torch

    B = getattr(args, "train_batch", 32)
    S = 256
    g = torch.Generator(device="cpu").manual_seed(4321 + rank)
    x = torch.randn((B, 3, S, S), generator=g).to(dev)
    y = (torch.empty((B, 16, S, S)).exponential_(1.0 / 20.0, generator=g).clamp_(0, 255).floor_() / 255.0 * 1.8 - 0.9).to(dev)
    w = torch.linspace(1.0, 10.6, 16)
    tr = Trainer(model, marker_weights=w, batch_size=B, total_steps=10000)
    steps = max(3, min(args.steps, 10))
    for _ in range(3):
        tr.step(x, y)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = tr.step(x, y)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    model.eval()
    return {"metric": "tiles_per_sec_train_256px_16ch", "value": world * B / ms * 1e3, "unit": "tiles/s",
            "ms_per_step": ms, "steps": steps, "batch_per_gpu": B, "loss": float(loss.item()),
            "tflops": world * B * 1633.87 / ms, "scaling": "weak",
            "note": "fwd+bwd+WeightedMSE+clip+Adam, every kernel hand-written (encoder, decoder with train-mode BatchNorm, "
                    "loss, optimiser); NCCL AVG all-reduce of 26.8 MB in 2 buckets overlapped with the encoder backward"}
