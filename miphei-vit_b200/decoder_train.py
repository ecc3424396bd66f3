"""Training-mode decoder (Detail_Capture, src/generators/mipheivit.py:166-220) on the hand-written kernels: forward with
BatchNorm batch statistics, and the full backward (conv dgrad / wgrad on the implicit GEMM, BN+ReLU backward, bilinear
adjoints, the 16 gated heads).

Precision.  Feature maps and forward conv operands are **fp16** (NHWC), raw conv outputs fp32, gradients bf16, every
reduction (statistics, weight gradients) fp32.  The reference trains under fp16 autocast (configs/config.yaml:23); with
bf16 maps the LoRA gradients — a small residual of a DC-dominated decoder gradient — lose cosine 3e-3 against the fp32
oracle through ReLU-mask / statistic perturbations of the FORWARD values (tests/tools/emulate_bf16_grads.py), with fp16
maps 3e-4.  The tensor core takes either 16-bit format, but both operands of an instruction must share it (a mixed
descriptor is an illegal instruction on sm_100a — measured), so the weight-gradient GEMMs, whose other operand is a bf16
gradient, read bf16 twins of the activation maps (mv_f16_to_bf16: +4 B per element, ~0.3 % of a step).

No framework ops inside a step.  All trainable decoder parameters are read from ONE flat fp32 buffer (the trainer's, or a
private staging copy on the plain autograd path) through index tables built once: `pack()` is three table-driven launches
(mv_gather_cast) that produce the fp16 forward operands, the bf16 data-gradient operands (taps flipped, transposed) and the
fp32 head constants; the backward leaves every parameter gradient in one fp32 accumulator arena and ONE more gather puts
them into parameter layout inside the flat gradient buffer.  The whole step is therefore capturable in a CUDA graph.

Heads backward in closed form.  With f the 32-channel full-resolution map, a = W1 f + b1, ahat = BN(a), r = relu(ahat),
u = w2 . r + b2, g = sigmoid(u), t = W3 f (per tap), pred = tanh(b3 + sum_tap g t), the kernels materialise only
dt [M,144], du [M,16] and e [M,256] (e = du where ahat > 0 else 0).  Everything else — the BatchNorm batch-statistic
terms included — follows from the small matrices f^T e, f^T f, f^T dt (fp32, 40x256 / 40x32 / 40x144) computed by the
split-K tensor-core GEMM and folded by mv_heads_bwd_algebra, so no [M, 256] gradient is ever written twice.
"""
import torch

from . import ops, packing

SKIP_CH = [192, 96, 48, 3]       # real channels of the skip input of fusion block i (D3, D2, D1, image)
FUS_OUT = [256, 128, 64, 32]
CS_OUT = [48, 96, 192]
ADT = torch.float16              # activation / forward-operand storage format of the training decoder


def _pad64(c):
    return (c + 63) // 64 * 64


def _flip_t(w):
    """conv weight [Cout, Cin, 3, 3] -> data-gradient conv weight [Cin, Cout, 3, 3] (taps flipped)."""
    return w.flip(2, 3).permute(1, 0, 2, 3).contiguous()


def _unpack_wgrad(v, cout, splits_real):
    """packed [Cout, 9 * sum(pad64)] -> [Cout, sum(real), 3, 3] (applied to POSITION tensors when the tables are built)."""
    v = v.view(cout, 9, -1)
    parts, off = [], 0
    for c in splits_real:
        parts.append(v[:, :, off:off + c])
        off += _pad64(c)
    w = torch.cat(parts, 2) if len(parts) > 1 else parts[0]
    return w.reshape(cout, 3, 3, -1).permute(0, 3, 1, 2).contiguous()


class _Arena:
    """Named segments of one contiguous device buffer (64-element aligned). With `indexed`, every element carries the
    index of the flat-parameter value it is gathered from (-1 = zero)."""

    def __init__(self, dtype, device, indexed):
        self.dtype, self.device, self.indexed = dtype, device, indexed
        self.items, self.size = [], 0

    def add(self, key, shape=None, pos=None):
        if pos is not None:
            shape = tuple(pos.shape)
        n = 1
        for s in shape:
            n *= s
        self.items.append((key, self.size, tuple(shape), pos))
        off = self.size
        self.size = (self.size + n + 63) // 64 * 64
        return off

    def finalize(self):
        self.data = torch.zeros(max(self.size, 64), dtype=self.dtype, device=self.device)
        self.v, self.off = {}, {}
        idx = torch.full((max(self.size, 64),), -1, dtype=torch.int64) if self.indexed else None
        for key, off, shape, pos in self.items:
            n = 1
            for s in shape:
                n *= s
            self.v[key] = self.data[off:off + n].view(shape)
            self.off[key] = off
            if self.indexed:
                idx[off:off + n] = pos.detach().double().cpu().flatten().round().long() - 1
        if self.indexed:
            self.idx = idx.to(torch.int32).to(self.device)
        return self


class DecoderTrain:
    def __init__(self, eng):
        self.eng = eng
        self.dec = eng.model.decoder
        self._bufs = {}
        dev = eng.device
        D = eng.D
        dec = self.dec
        self.layout, self.n_dec = packing.decoder_layout(eng.model)
        # position tensors: element i of the flat parameter buffer is represented by the value i + 1 (0 = padding), so the
        # ordinary packing functions, run once on positions instead of values, yield the gather tables
        pos = {}
        for name, p, off, n in self.layout:
            pos[name] = torch.arange(off + 1, off + n + 1, dtype=torch.float64).view(p.shape)
        P = lambda mod_prefix, leaf: pos["decoder." + mod_prefix + leaf]  # noqa: E731
        wf = _Arena(ADT, dev, True)             # forward operands (fp16)
        wb = _Arena(torch.bfloat16, dev, True)  # data-gradient operands (bf16: they meet bf16 gradients)
        wc = _Arena(torch.float32, dev, True)   # fp32 constants: BatchNorm affine, head MLP
        acc = _Arena(torch.float32, dev, False)  # per-step accumulators / parameter-gradient sources (zeroed per step)
        self.cs, self.fu = [], []
        bns = []
        for i, m in enumerate(dec.convstream.convs):
            pre = "convstream.convs.%d." % i
            w = P(pre, "conv.weight")
            cout, cin = w.shape[0], w.shape[1]
            wf.add("cs%d.w" % i, pos=packing.pack_conv3x3(w, [cin], dtype=None))
            if i > 0:
                wb.add("cs%d.wd" % i, pos=packing.pack_conv3x3(_flip_t(w), [cout], dtype=None))
            wc.add("cs%d.gamma" % i, pos=P(pre, "bn.weight"))
            wc.add("cs%d.beta" % i, pos=P(pre, "bn.bias"))
            self.cs.append(dict(key="cs%d" % i, bn=m.bn, cin=cin, cout=cout, splits=[cin], names=(pre + "conv.weight", pre + "bn.weight", pre + "bn.bias")))
            bns.append(m.bn)
        up_ch = [D, 256, 128, 64]
        for i, m in enumerate(dec.fusion_blks):
            pre = "fusion_blks.%d.conv." % i
            w = P(pre, "conv.weight")
            c0, c1, cout = SKIP_CH[i], up_ch[i], w.shape[0]
            wf.add("fu%d.w" % i, pos=packing.pack_conv3x3(w, [c0, c1], dtype=None))
            wt = _flip_t(w)  # [Cin_total, Cout, 3, 3]
            wb.add("fu%d.wd" % i, pos=packing.pack_conv3x3(wt if i < 3 else wt[c0:], [cout], dtype=None))  # block 3: the image needs no gradient
            wc.add("fu%d.gamma" % i, pos=P(pre, "bn.weight"))
            wc.add("fu%d.beta" % i, pos=P(pre, "bn.bias"))
            self.fu.append(dict(key="fu%d" % i, bn=m.conv.bn, c0=c0, c1=c1, cout=cout, splits=[c0, c1], names=(pre + "conv.weight", pre + "bn.weight", pre + "bn.bias")))
            bns.append(m.conv.bn)
        Hh = eng.heads_out
        heads = [getattr(dec, "segmentation_head_%d" % h) for h in range(Hh)]
        self.heads, self.n1 = heads, 16 * Hh
        z = lambda *s: torch.zeros(s, dtype=torch.float64)  # noqa: E731
        W1, b1, gam, bet, w2, b2, b3, W3 = z(256, 32), z(256), z(256), z(256), z(256), z(16), z(16), z(16, 32, 3, 3)
        for h in range(Hh):
            pre = "segmentation_head_%d." % h
            sl = slice(16 * h, 16 * h + 16)
            W1[sl] = P(pre, "0.psi.0.weight").flatten(1)
            b1[sl] = P(pre, "0.psi.0.bias")
            gam[sl] = P(pre, "0.psi.1.weight")
            bet[sl] = P(pre, "0.psi.1.bias")
            w2[sl] = P(pre, "0.psi.3.weight").flatten()
            b2[h] = P(pre, "0.psi.3.bias")[0]
            W3[h] = P(pre, "1.weight")[0]
            b3[h] = P(pre, "1.bias")[0]
            bns.append(heads[h][0].psi[1])
        for k, t in (("W1", W1), ("b1", b1), ("gam", gam), ("bet", bet), ("w2", w2), ("b2", b2), ("b3", b3)):
            wc.add("hd." + k, pos=t)
        w1p = z(256, 64)
        w1p[:, :32] = W1
        wf.add("hd.gate_w", pos=w1p)                                         # B of the gate GEMM (K = 32 of pitch 64)
        wf.add("hd.conv_w", pos=packing.pack_conv3x3(W3, [32], dtype=None))  # [16, 576] for HEAD_CONV
        w3t = W3.permute(2, 3, 0, 1).reshape(144, 32)                        # row tap*16 + h, col c
        w3tp = z(144, 64)
        w3tp[:, :32] = w3t
        wf.add("hd.w3t", pos=w3tp)                                           # B of T = f W3t^T   (K = 32)
        w3tT = z(32, 192)
        w3tT[:, :144] = w3t.t()
        wb.add("hd.w3tT", pos=w3tT)                                          # B of df1 = dt W3t  (K = 144)
        # ---- accumulators (zeroed by one memset per step) and the parameter-gradient gather table
        for L in self.cs + self.fu:
            k, cout = L["key"], L["cout"]
            kp = 9 * sum(_pad64(s if s != 3 else 8) for s in L["splits"])
            L["kp"] = kp
            acc.add(k + ".stats", (2, cout))
            acc.add(k + ".sums", (2, cout))
            acc.add(k + ".dwp", (cout, kp))
        for k, s in (("hd.FF", (40, 32)), ("hd.db3", (16,)), ("hd.db2", (16,)), ("hd.G3", (40, 144)), ("hd.E", (40, 256)),
                     ("hd.dW1", (256, 32)), ("hd.S2", (256,)), ("hd.S1", (256,)), ("hd.dw2", (256,))):
            acc.add(k, s)
        self.wf, self.wb, self.wc, self.acc = wf.finalize(), wb.finalize(), wc.finalize(), acc.finalize()
        apos = lambda key: torch.arange(1, self.acc.v[key].numel() + 1, dtype=torch.float64).view(self.acc.v[key].shape) \
            + self.acc.off[key]  # noqa: E731
        gidx = torch.full((self.n_dec,), -1, dtype=torch.int64)
        offs = {name: (off, n) for name, _, off, n in self.layout}

        def put(name, src_pos):
            off, n = offs["decoder." + name]
            assert src_pos.numel() == n, (name, src_pos.shape, n)
            gidx[off:off + n] = src_pos.flatten().round().long() - 1

        for L in self.cs + self.fu:
            wn, gn, bn_ = L["names"]
            put(wn, _unpack_wgrad(apos(L["key"] + ".dwp"), L["cout"], L["splits"]))
            sums = apos(L["key"] + ".sums")
            put(gn, sums[1])   # d gamma
            put(bn_, sums[0])  # d beta
        dW1, S2, S1, dw2, db2, db3 = (apos("hd." + k) for k in ("dW1", "S2", "S1", "dw2", "db2", "db3"))
        dW3 = apos("hd.G3")[:32].reshape(32, 9, 16).permute(2, 0, 1).reshape(16, 32, 3, 3)
        for h in range(Hh):
            pre = "segmentation_head_%d." % h
            sl = slice(16 * h, 16 * h + 16)
            put(pre + "0.psi.0.weight", dW1[sl])
            # psi[0].bias: analytically zero gradient (the BatchNorm removes the mean) -> stays -1
            put(pre + "0.psi.1.weight", S2[sl])
            put(pre + "0.psi.1.bias", S1[sl])
            put(pre + "0.psi.3.weight", dw2[sl])
            put(pre + "0.psi.3.bias", db2[h:h + 1])
            put(pre + "1.weight", dW3[h:h + 1])
            put(pre + "1.bias", db3[h:h + 1])
        self.gidx = gidx.to(torch.int32).to(dev)
        # ---- BatchNorm buffers: one int64 vector for the 23 num_batches_tracked counters, two fp32 vectors for the
        # running statistics of the 16 head BatchNorms (one finalize launch); the module's buffers become views of them
        self.bns = bns
        self.nbt = torch.zeros(len(bns), dtype=torch.int64, device=dev)
        self.rm_cat = torch.zeros(256, device=dev)
        self.rv_cat = torch.ones(256, device=dev)
        with torch.no_grad():
            for i, bn in enumerate(bns):
                self.nbt[i] = bn.num_batches_tracked
                bn._buffers["num_batches_tracked"] = self.nbt[i]
                if bn.running_mean.dtype != torch.float32 or bn.running_var.dtype != torch.float32:
                    raise ops._lib.MipheiB200Error("training needs fp32 BatchNorm running statistics (model.float())")
            for h, hd in enumerate(heads):
                bn = hd[0].psi[1]
                self.rm_cat[16 * h:16 * h + 16] = bn.running_mean
                self.rv_cat[16 * h:16 * h + 16] = bn.running_var
                bn._buffers["running_mean"] = self.rm_cat[16 * h:16 * h + 16]
                bn._buffers["running_var"] = self.rv_cat[16 * h:16 * h + 16]
        self.ca_t = torch.zeros((32, 256), dtype=torch.bfloat16, device=dev)
        self.mx_n = torch.zeros((32, 64), dtype=ADT, device=dev)
        self.kshift = torch.zeros(32, dtype=torch.float32, device=dev)
        self.mx_scale = torch.ones(32, dtype=torch.float32, device=dev)
        # plain autograd path: private flat staging copy of the parameters and a private flat gradient buffer
        self._own_src = None
        self._own_grad = None
        self.generation = 0

    # ------------------------------------------------------------------ flat parameter / gradient buffers
    def _flat(self):
        fp = getattr(self.eng, "flat_params", None)
        if fp is not None:  # trainer mode: the parameters ARE views of the flat buffer, decoder segment first
            flat, gflat, n_dec = fp
            assert n_dec == self.n_dec
            return flat[:n_dec], gflat[:n_dec], False
        if self._own_src is None:
            self._own_src = torch.zeros(self.n_dec, dtype=torch.float32, device=self.eng.device)
            self._own_grad = torch.zeros(self.n_dec, dtype=torch.float32, device=self.eng.device)
            self._src_views = [self._own_src[off:off + n].view(p.shape) for _, p, off, n in self.layout]
        return self._own_src, self._own_grad, True

    def pack(self):
        """Kernel operand layouts from the current parameter values: three table-driven launches."""
        src, _, own = self._flat()
        if own:
            with torch.no_grad():
                for v, (_, p, _, _) in zip(self._src_views, self.layout):
                    v.copy_(p.detach())
        ops.gather_cast(src, self.wf.idx, self.wf.data)
        ops.gather_cast(src, self.wb.idx, self.wb.data)
        ops.gather_cast(src, self.wc.idx, self.wc.data)

    def grad_views(self):
        """parameter -> gradient view of the private flat gradient buffer (plain autograd path)."""
        _, g, _ = self._flat()
        return {p: g[off:off + n].view(p.shape) for _, p, off, n in self.layout}

    # ------------------------------------------------------------------ buffers
    def _buf(self, key, shape, dtype, zero=False):
        k = (key, tuple(shape), dtype)
        t = self._bufs.get(k)
        if t is None:
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.eng.device)
            self._bufs[k] = t
        return t

    # ------------------------------------------------------------------ forward (train mode)
    def _conv_bn_relu(self, L, src0, src1, stride, Bn, Ho):
        key, cout = L["key"], L["cout"]
        M = Bn * Ho * Ho
        z = self._buf(key + ".z", (M, cout), torch.float32)  # fp32: no 16-bit-rounding ReLU-mask flips
        y = self._buf(key + ".y", (M, cout), ADT)
        stats = self.acc.v[key + ".stats"]
        ops.gemm(src0, self.wf.v[key + ".w"], conv=dict(stride=stride, a2=src1), colstats=stats, out=z)
        bn = L["bn"]
        fin = self._buf(key + ".fin", (4, cout), torch.float32)
        ops.bn_finalize(stats, M, self.wc.v[key + ".gamma"], self.wc.v[key + ".beta"], bn.running_mean, bn.running_var,
                        momentum=bn.momentum, eps=bn.eps, out=fin)
        ops.bn_relu_apply(z, fin[0], fin[1], out=y)
        L["z"], L["y"], L["fin"], L["M"], L["Ho"] = z, y, fin, M, Ho
        return y.view(Bn, Ho, Ho, cout)

    def _twin(self, key, t):
        """bf16 copy of an fp16 map (B operand of the weight-gradient GEMM, next to the bf16 dz^T)."""
        return ops.f16_to_bf16(t, out=self._buf(key, tuple(t.shape), torch.bfloat16))

    def forward(self, fmap, img):
        """fmap NHWC fp16 [B, t, t, D], img NHWC fp16 [B, S, S, 8] -> pred fp32 NCHW [B, heads, S, S] (persistent buffer)."""
        eng = self.eng
        assert fmap.dtype == ADT and img.dtype == ADT
        Bn, S = img.shape[0], img.shape[1]
        self.generation += 1
        ops.memset(self.acc.data)
        ops.add_i64(self.nbt, 1)
        d = [img]
        db = [self._twin("img.b", img)]
        for i, L in enumerate(self.cs):
            L["srcs"] = [db[i]]
            d.append(self._conv_bn_relu(L, d[i], None, 2, Bn, S >> (i + 1)))
            db.append(self._twin("cs%d.yb" % i, d[-1]))
        f = fmap
        for i, L in enumerate(self.fu):
            h2 = f.shape[1] * 2
            up = self._buf("fu%d.up" % i, (Bn, h2, h2, f.shape[3]), ADT)
            ops.upsample2x(f, out=up)
            L["srcs"] = [db[3 - i], self._twin("fu%d.upb" % i, up)]
            f = self._conv_bn_relu(L, d[3 - i], up, 1, Bn, h2)
        M = Bn * S * S
        f2 = f.view(M, 32)
        self.f, self.f2, self.Bn, self.S, self.M = f, f2, Bn, S, M
        # heads: batch statistics of the 256 gate units in closed form from the moments of f (E[f], E[f f^T]): one
        # 64 B/pixel pass, then the two fused inference kernels with the batch-statistic fold
        wc, wf = self.wc.v, self.wf.v
        FF = self.acc.v["hd.FF"]
        ops.gram32(f2, out=FF, zero=False)                              # rows 0..31 = f^T f, row 32 = 1^T f
        fin = self._buf("hd.fin", (4, 256), torch.float32)
        ops.heads_bn_from_gram(FF, M, wc["hd.W1"], wc["hd.b1"], wc["hd.gam"], wc["hd.bet"], self.rm_cat, self.rv_cat, out=fin,
                               w1_fmt=2 if ADT == torch.float16 else 1)
        self.hfin = fin
        gate = self._buf("hd.gate", (M, 16), torch.bfloat16, zero=True)  # columns >= heads stay zero
        ops.gemm(f2, wf["hd.gate_w"][:self.n1, :32], mode=ops.GEMM_HEAD_GATE, scale=fin[0], shift=fin[1], in2=wc["hd.w2"],
                 resid=wc["hd.b2"], out=gate[:, :eng.heads_out])
        self.gate = gate
        pred = self._buf("pred", (Bn, eng.heads_out, S, S), torch.float32)
        ops.gemm(f, wf["hd.conv_w"][:eng.heads_out], mode=ops.GEMM_HEAD_CONV, conv=dict(stride=1), shift=wc["hd.b3"], in2=gate,
                 out=pred)
        self.pred = pred
        return pred

    # ------------------------------------------------------------------ backward
    def _bn_conv_backward(self, L, dy, stride):
        """dy: [M, Cout] gradient of the layer output (post ReLU). Returns dz [M, Cout] bf16."""
        key, cout, M = L["key"], L["cout"], L["M"]
        dz = self._buf(key + ".dz", (M, cout), torch.bfloat16)
        ops.bn_relu_bwd(dy, L["y"], L["z"], L["fin"][2], L["fin"][3], self.wc.v[key + ".gamma"], sums=self.acc.v[key + ".sums"],
                        dz=dz)
        ld = (M + 7) // 8 * 8
        dzT = self._buf(key + ".dzT", (cout, ld), torch.bfloat16)
        ops.transpose_bf16(dz, out=dzT)
        srcs = L["srcs"]
        ops.gemm(dzT[:, :M], srcs[0], mode=ops.GEMM_NN_ATOMIC, conv=dict(stride=stride, a2=srcs[1] if len(srcs) > 1 else None),
                 out=self.acc.v[key + ".dwp"])
        return dz

    def backward(self, dpred):
        """dpred fp32 NCHW -> d fmap NHWC bf16; every parameter gradient lands in the flat gradient buffer."""
        eng = self.eng
        Bn, S, M = self.Bn, self.S, self.M
        bf = torch.bfloat16
        f2 = self.f2
        wc, wf, wb, acc = self.wc.v, self.wf.v, self.wb.v, self.acc.v
        # ---------------- heads
        ds = self._buf("hd.ds", (M, 16), bf)
        ops.heads_ds(dpred, self.pred, acc["hd.db3"], out=ds)
        T = self._buf("hd.T", (M, 144), bf)
        ops.gemm(f2, wf["hd.w3t"][:, :32], out=T)
        dt = self._buf("hd.dt", (M, 144), bf)
        du = self._buf("hd.du", (M, 16), bf)
        ops.heads_bwd_stencil(T, ds, self.gate, Bn, S, S, acc["hd.db2"], dt=dt, du=du)
        ld = (M + 7) // 8 * 8
        fT = self._buf("hd.fT", (40, ld), bf, zero=True)   # bf16 (converted by the transpose): it meets bf16 dt / e
        ops.transpose_bf16(f2, ones_row=True, out=fT)
        fTm = fT[:, :M]
        ops.gemm(fTm, dt, mode=ops.GEMM_NN_ATOMIC, out=acc["hd.G3"])             # f^T dt
        df1 = self._buf("hd.df1", (M, 32), torch.float32)
        ops.gemm(dt, wb["hd.w3tT"][:, :144], out=df1)                            # dt W3t
        fin = self.hfin
        e = self._buf("hd.e", (M, 256), bf)
        ops.gemm(f2, wf["hd.gate_w"][:, :32], scale=fin[0], shift=fin[1], act=ops.ACT_GATE_MASK, in2=du, out=e)
        ops.gemm(fTm, e, mode=ops.GEMM_NN_ATOMIC, out=acc["hd.E"])               # rows 0..31 = f^T e, row 32 = 1^T e
        # closed-form BatchNorm / gate / 1x1-conv gradients ([256, 32]-sized algebra, one CTA)
        ops.heads_bwd_algebra(acc["hd.E"], acc["hd.FF"], wc["hd.W1"], wc["hd.b1"], wc["hd.gam"], wc["hd.w2"], fin, M, self.n1,
                              acc["hd.dW1"], acc["hd.S2"], acc["hd.S1"], acc["hd.dw2"], self.ca_t, self.mx_n, self.mx_scale, self.kshift)
        # d f = dt W3t + e Ca - f Mx - (K0 + K1)
        dfa = self._buf("hd.dfa", (M, 32), torch.float32)
        ops.gemm(e, self.ca_t, resid=df1, shift=self.kshift, out=dfa)
        dy = self._buf("hd.dy3", (M, 32), bf)
        ops.gemm(f2, self.mx_n[:, :32], scale=self.mx_scale, resid=dfa, out=dy)
        # ---------------- fusion blocks
        gskip = {}
        dfmap = None
        for i in range(3, -1, -1):
            L = self.fu[i]
            dz = self._bn_conv_backward(L, dy, 1)
            Ho, cout = L["Ho"], L["cout"]
            ncol = L["c1"] if i == 3 else L["c0"] + L["c1"]
            dx = self._buf("fu%d.dx" % i, (L["M"], ncol), bf)
            ops.gemm(dz.view(Bn, Ho, Ho, cout), wb["fu%d.wd" % i], conv=dict(stride=1), out=dx)
            c0 = 0 if i == 3 else L["c0"]
            if i < 3:
                gskip[3 - i] = dx[:, :c0]
            dup = dx.view(Bn, Ho, Ho, ncol)[..., c0:]
            dprev = self._buf("fu%d.dprev" % i, (Bn, Ho // 2, Ho // 2, L["c1"]), bf)
            ops.upsample2x_bwd(dup, out=dprev)
            if i > 0:
                dy = dprev.view(-1, L["c1"])
            else:
                dfmap = dprev
        # ---------------- ConvStream
        dyD = gskip[3]
        for j in range(2, -1, -1):
            L = self.cs[j]
            dz = self._bn_conv_backward(L, dyD, 2)
            if j > 0:
                Ho, cout = L["Ho"], L["cout"]
                u = self._buf("cs%d.u" % j, (Bn, 2 * Ho, 2 * Ho, cout), bf)
                ops.zero_insert2x(dz.view(Bn, Ho, Ho, cout), out=u)
                dxs = self._buf("cs%d.dx" % j, (Bn * 4 * Ho * Ho, L["cin"]), bf)
                ops.gemm(u, wb["cs%d.wd" % j], conv=dict(stride=1), out=dxs)
                dsum = self._buf("cs%d.dsum" % j, (Bn * 4 * Ho * Ho, L["cin"]), bf)
                ops.add_bf16(gskip[j], dxs, out=dsum)
                dyD = dsum
        # ---------------- every decoder parameter gradient, parameter layout, one launch
        _, gdst, _ = self._flat()
        ops.gather_cast(self.acc.data, self.gidx, gdst)
        return dfmap
