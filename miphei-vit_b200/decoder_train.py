"""Training-mode decoder (Detail_Capture, src/generators/mipheivit.py:166-220) on the hand-written kernels: forward with
BatchNorm batch statistics, and the full backward (conv dgrad / wgrad on the implicit GEMM, BN+ReLU backward, bilinear
adjoints, the 16 gated heads).  Activations are NHWC bf16; every reduction (statistics, weight gradients) accumulates in
fp32.

Heads backward in closed form.  With f the 32-channel full-resolution map, a = W1 f + b1, ahat = BN(a), r = relu(ahat),
u = w2 . r + b2, g = sigmoid(u), t = W3 f (per tap), pred = tanh(b3 + sum_tap g t), the kernels materialise only
dt [M,144], du [M,16] and e [M,256] (e = du where ahat > 0 else 0).  Everything else — the BatchNorm batch-statistic
terms included — follows from the small matrices f^T e, f^T f, f^T dt (fp32, 40x256 / 40x32 / 40x144) computed by the
split-K tensor-core GEMM, so no [M, 256] gradient is ever written twice.
"""
import torch

from . import ops, packing

SKIP_CH = [192, 96, 48, 3]       # real channels of the skip input of fusion block i (D3, D2, D1, image)
FUS_OUT = [256, 128, 64, 32]
CS_OUT = [48, 96, 192]


def _pad64(c):
    return (c + 63) // 64 * 64


def _flip_t(w):
    """conv weight [Cout, Cin, 3, 3] -> data-gradient conv weight [Cin, Cout, 3, 3] (taps flipped)."""
    return w.detach().float().flip(2, 3).permute(1, 0, 2, 3).contiguous()


class DecoderTrain:
    def __init__(self, eng):
        self.eng = eng
        self.dec = eng.model.decoder
        self._bufs = {}
        self._head_stats_views = False

    # ------------------------------------------------------------------ weights (re-packed when they change)
    def pack(self):
        dec, eng = self.dec, self.eng
        D = eng.D
        with torch.no_grad():
            self.cs = []
            for i, m in enumerate(dec.convstream.convs):
                w = m.conv.weight
                cin = w.shape[1]
                self.cs.append(dict(w=packing.pack_conv3x3(w, [cin]), wd=packing.pack_conv3x3(_flip_t(w), [w.shape[0]]) if i > 0 else None,
                                    bn=m.bn, conv=m.conv, cin=cin, cout=w.shape[0]))
            self.fu = []
            up_ch = [D, 256, 128, 64]
            for i, m in enumerate(dec.fusion_blks):
                w = m.conv.conv.weight
                c0, c1 = SKIP_CH[i], up_ch[i]
                wt = _flip_t(w)  # [Cin_total, Cout, 3, 3]
                wd_rows = wt if i < 3 else wt[c0:]  # block 3: the image needs no gradient
                self.fu.append(dict(w=packing.pack_conv3x3(w, [c0, c1]), wd=packing.pack_conv3x3(wd_rows, [w.shape[0]]),
                                    bn=m.conv.bn, conv=m.conv.conv, c0=c0, c1=c1, cout=w.shape[0]))
            heads = [getattr(dec, "segmentation_head_%d" % h) for h in range(eng.heads_out)]
            self.heads = heads
            Hh = len(heads)
            dev = eng.device
            W1 = torch.zeros((256, 32), device=dev)
            b1, gam, bet, w2 = (torch.zeros(256, device=dev) for _ in range(4))
            gam.fill_(1.0)
            b2, b3 = torch.zeros(16, device=dev), torch.zeros(16, device=dev)
            W3 = torch.zeros((16, 32, 3, 3), device=dev)
            for h, hd in enumerate(heads):
                psi = hd[0].psi
                W1[16 * h:16 * h + 16] = psi[0].weight.detach().float().flatten(1)
                b1[16 * h:16 * h + 16] = psi[0].bias.detach().float()
                gam[16 * h:16 * h + 16] = psi[1].weight.detach().float()
                bet[16 * h:16 * h + 16] = psi[1].bias.detach().float()
                w2[16 * h:16 * h + 16] = psi[3].weight.detach().float().flatten()
                b2[h] = psi[3].bias.detach().float()[0]
                W3[h] = hd[1].weight.detach().float()[0]
                b3[h] = hd[1].bias.detach().float()[0]
            n1 = 16 * Hh
            self.W1, self.b1, self.gam, self.bet, self.w2, self.b2, self.b3 = W1, b1, gam, bet, w2, b2, b3
            self.n1 = n1
            w1p = torch.zeros((256, 64), device=dev)
            w1p[:, :32] = W1
            self.gate_w = w1p.to(torch.bfloat16)
            self.W1r = self.gate_w[:, :32].float().contiguous()      # the bf16-rounded weights the gate GEMM multiplies by
            self.conv_w = packing.pack_conv3x3(W3, [32])            # [16, 576] for HEAD_CONV
            w3t = W3.permute(2, 3, 0, 1).reshape(144, 32)            # row tap*16 + h, col c
            self.W3t = w3t
            w3tp = torch.zeros((144, 64), device=dev)
            w3tp[:, :32] = w3t
            self.w3t_b = w3tp.to(torch.bfloat16)                     # B operand of T = f W3t^T   (K = 32)
            w3tT = torch.zeros((32, 192), device=dev)
            w3tT[:, :144] = w3t.t()
            self.w3tT_b = w3tT.to(torch.bfloat16)                    # B operand of df1 = dt W3t  (K = 144)
            if not self._head_stats_views:
                # running statistics of the 16 head BatchNorms live in two concatenated buffers (one finalize launch)
                self.rm_cat = torch.zeros(256, device=dev)
                self.rv_cat = torch.ones(256, device=dev)
                for h, hd in enumerate(heads):
                    bn = hd[0].psi[1]
                    self.rm_cat[16 * h:16 * h + 16] = bn.running_mean
                    self.rv_cat[16 * h:16 * h + 16] = bn.running_var
                    bn._buffers["running_mean"] = self.rm_cat[16 * h:16 * h + 16]
                    bn._buffers["running_var"] = self.rv_cat[16 * h:16 * h + 16]
                self._head_stats_views = True

    # ------------------------------------------------------------------ buffers
    def _buf(self, key, shape, dtype, zero=False):
        k = (key, tuple(shape), dtype)
        t = self._bufs.get(k)
        if t is None:
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.eng.device)
            self._bufs[k] = t
        return t

    # ------------------------------------------------------------------ forward (train mode)
    def _conv_bn_relu(self, key, L, src0, src1, stride, Bn, Ho):
        cout = L["cout"]
        M = Bn * Ho * Ho
        bf = torch.bfloat16
        z = self._buf(key + ".z", (M, cout), torch.float32)  # fp32: no bf16-rounding ReLU-mask flips
        y = self._buf(key + ".y", (M, cout), bf)
        stats = self._buf(key + ".stats", (2, cout), torch.float32)
        stats.zero_()
        ops.gemm(src0, L["w"], conv=dict(stride=stride, a2=src1), colstats=stats, out=z)
        bn = L["bn"]
        fin = self._buf(key + ".fin", (4, cout), torch.float32)
        ops.bn_finalize(stats, M, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var,
                        momentum=bn.momentum, eps=bn.eps, out=fin)
        bn.num_batches_tracked += 1
        ops.bn_relu_apply(z, fin[0], fin[1], out=y)
        L["z"], L["y"], L["fin"], L["M"], L["Ho"] = z, y, fin, M, Ho
        return y.view(Bn, Ho, Ho, cout)

    def forward(self, fmap, img8):
        """fmap NHWC bf16 [B, t, t, D], img8 NHWC bf16 [B, S, S, 8] -> pred fp32 NCHW [B, heads, S, S]."""
        eng = self.eng
        Bn, S = img8.shape[0], img8.shape[1]
        bf = torch.bfloat16
        d = [img8]
        for i, L in enumerate(self.cs):
            L["src"] = d[i]
            d.append(self._conv_bn_relu("cs%d" % i, L, d[i], None, 2, Bn, S >> (i + 1)))
        f = fmap
        for i, L in enumerate(self.fu):
            h2 = f.shape[1] * 2
            up = self._buf("fu%d.up" % i, (Bn, h2, h2, f.shape[3]), bf)
            ops.upsample2x(f, out=up)
            L["src0"], L["src1"] = d[3 - i], up
            f = self._conv_bn_relu("fu%d" % i, L, d[3 - i], up, 1, Bn, h2)
        M = Bn * S * S
        f2 = f.view(M, 32)
        self.f, self.f2, self.Bn, self.S, self.M = f, f2, Bn, S, M
        # heads: statistics pass, then the two fused inference kernels with the batch-statistic fold
        # batch statistics of the 256 gate units in closed form from the moments of f (E[f], E[f f^T]): one 64 B/pixel pass
        FF = self._buf("hd.FF", (40, 32), torch.float32)
        ops.gram32(f2, out=FF)                                          # rows 0..31 = f^T f, row 32 = 1^T f
        fin = self._buf("hd.fin", (4, 256), torch.float32)
        ops.heads_bn_from_gram(FF, M, self.W1r, self.b1, self.gam, self.bet, self.rm_cat, self.rv_cat, out=fin)
        for hd in self.heads:
            hd[0].psi[1].num_batches_tracked += 1
        self.hfin = fin
        gate = self._buf("hd.gate", (M, 16), bf, zero=True)  # columns >= heads stay zero
        ops.gemm(f2, self.gate_w[:self.n1, :32], mode=ops.GEMM_HEAD_GATE, scale=fin[0], shift=fin[1], in2=self.w2,
                 resid=self.b2, out=gate[:, :eng.heads_out])
        self.gate = gate
        pred = torch.empty((Bn, eng.heads_out, S, S), dtype=torch.float32, device=eng.device)
        ops.gemm(f, self.conv_w[:eng.heads_out], mode=ops.GEMM_HEAD_CONV, conv=dict(stride=1), shift=self.b3, in2=gate, out=pred)
        self.pred = pred
        return pred

    # ------------------------------------------------------------------ backward
    def _unpack_wgrad(self, dwp, cout, splits_real):
        """packed [Cout, 9 * sum(pad64)] fp32 -> [Cout, sum(real), 3, 3]."""
        v = dwp.view(cout, 9, -1)
        parts, off = [], 0
        for c in splits_real:
            parts.append(v[:, :, off:off + c])
            off += _pad64(c)
        w = torch.cat(parts, 2) if len(parts) > 1 else parts[0]
        return w.reshape(cout, 3, 3, -1).permute(0, 3, 1, 2).contiguous()

    def _bn_conv_backward(self, key, L, dy, srcs, stride, splits_real, grads):
        """dy: [M, Cout] gradient of the layer output (post ReLU). Returns dz (NHWC view)."""
        cout, M = L["cout"], L["M"]
        bn = L["bn"]
        dz = self._buf(key + ".dz", (M, cout), torch.bfloat16)
        sums = self._buf(key + ".sums", (2, cout), torch.float32)
        ops.bn_relu_bwd(dy, L["y"], L["z"], L["fin"][2], L["fin"][3], bn.weight.detach(), sums=sums, dz=dz)
        grads[bn.weight] = sums[1].clone()
        grads[bn.bias] = sums[0].clone()
        ld = (M + 7) // 8 * 8
        dzT = self._buf(key + ".dzT", (cout, ld), torch.bfloat16)
        ops.transpose_bf16(dz, out=dzT)
        kp = 9 * sum(_pad64(s.shape[3]) for s in srcs)
        dwp = self._buf(key + ".dwp", (cout, kp), torch.float32)
        dwp.zero_()
        ops.gemm(dzT[:, :M], srcs[0], mode=ops.GEMM_NN_ATOMIC, conv=dict(stride=stride, a2=srcs[1] if len(srcs) > 1 else None),
                 out=dwp)
        grads[L["conv"].weight] = self._unpack_wgrad(dwp, cout, splits_real)
        return dz

    def backward(self, dpred):
        """dpred fp32 NCHW -> (dict parameter -> gradient, d fmap NHWC bf16)."""
        eng = self.eng
        Bn, S, M = self.Bn, self.S, self.M
        bf = torch.bfloat16
        grads = {}
        f2 = self.f2
        n = float(M)
        # ---------------- heads
        db3 = self._buf("hd.db3", (16,), torch.float32)
        db3.zero_()
        ds = self._buf("hd.ds", (M, 16), bf)
        ops.heads_ds(dpred.contiguous(), self.pred, db3, out=ds)
        T = self._buf("hd.T", (M, 144), bf)
        ops.gemm(f2, self.w3t_b[:, :32], out=T)
        db2 = self._buf("hd.db2", (16,), torch.float32)
        db2.zero_()
        dt = self._buf("hd.dt", (M, 144), bf)
        du = self._buf("hd.du", (M, 16), bf)
        ops.heads_bwd_stencil(T, ds, self.gate, Bn, S, S, db2, dt=dt, du=du)
        ld = (M + 7) // 8 * 8
        fT = self._buf("hd.fT", (40, ld), bf, zero=True)
        ops.transpose_bf16(f2, ones_row=True, out=fT)
        fTm = fT[:, :M]
        G3 = self._buf("hd.G3", (40, 144), torch.float32)
        G3.zero_()
        ops.gemm(fTm, dt, mode=ops.GEMM_NN_ATOMIC, out=G3)            # f^T dt
        df1 = self._buf("hd.df1", (M, 32), torch.float32)
        ops.gemm(dt, self.w3tT_b[:, :144], out=df1)                     # dt W3t
        fin = self.hfin
        scale_g, shift_g, mean, rstd = fin[0], fin[1], fin[2], fin[3]
        e = self._buf("hd.e", (M, 256), bf)
        ops.gemm(f2, self.gate_w[:, :32], scale=scale_g, shift=shift_g, act=ops.ACT_GATE_MASK, in2=du, out=e)
        E = self._buf("hd.E", (40, 256), torch.float32)
        E.zero_()
        ops.gemm(fTm, e, mode=ops.GEMM_NN_ATOMIC, out=E)               # rows 0..31 = f^T e, row 32 = 1^T e
        FF = self._buf("hd.FF", (40, 32), torch.float32)                # f^T f / 1^T f, computed by the forward pass
        # small fp32 algebra ([256, 32]-sized) — closed-form BatchNorm / gate / 1x1-conv gradients
        W1, b1, gam, w2 = self.W1, self.b1, self.gam, self.w2
        EF, E1 = E[:32].t(), E[32]
        F2, F1 = FF[:32], FF[32]
        A = (W1 * EF).sum(1)
        S1 = w2 * E1
        S2 = w2 * rstd * (A + (b1 - mean) * E1)
        dw2 = scale_g * A + shift_g * E1
        XF = rstd[:, None] * (W1 @ F2 + (b1 - mean)[:, None] * F1[None, :])
        gr = gam * rstd
        dW1 = gr[:, None] * (w2[:, None] * EF - (S1 / n)[:, None] * F1[None, :] - (S2 / n)[:, None] * XF)
        Ca = (gr * w2)[:, None] * W1
        K0 = ((gr * S1 / n)[:, None] * W1).sum(0)
        k2 = gr * S2 / n
        Mx = W1.t() @ ((k2 * rstd)[:, None] * W1)
        K1 = ((k2 * rstd * (b1 - mean))[:, None] * W1).sum(0)
        dW3 = G3[:32].reshape(32, 9, 16).permute(2, 0, 1).reshape(16, 32, 3, 3)
        for h, hd in enumerate(self.heads):
            psi = hd[0].psi
            sl = slice(16 * h, 16 * h + 16)
            grads[psi[0].weight] = dW1[sl].reshape(16, 32, 1, 1)
            grads[psi[0].bias] = torch.zeros_like(psi[0].bias)  # analytically zero: BatchNorm removes the mean
            grads[psi[1].weight] = S2[sl]
            grads[psi[1].bias] = S1[sl]
            grads[psi[3].weight] = dw2[sl].reshape(1, 16, 1, 1)
            grads[psi[3].bias] = db2[h:h + 1].clone()
            grads[hd[1].weight] = dW3[h:h + 1].contiguous()
            grads[hd[1].bias] = db3[h:h + 1].clone()
        # d f = dt W3t + e Ca - f Mx - (K0 + K1)
        dfa = self._buf("hd.dfa", (M, 32), torch.float32)
        ops.gemm(e, Ca.t().contiguous().to(bf), resid=df1, shift=(-(K0 + K1)).contiguous(), out=dfa)
        mxb = torch.zeros((32, 64), device=eng.device)
        mxb[:, :32] = (-Mx).t()
        dy = self._buf("hd.dy3", (M, 32), bf)
        ops.gemm(f2, mxb.to(bf)[:, :32], resid=dfa, out=dy)
        # ---------------- fusion blocks
        gskip = {}
        dfmap = None
        for i in range(3, -1, -1):
            L = self.fu[i]
            srcs = [L["src0"], L["src1"]]
            dz = self._bn_conv_backward("fu%d" % i, L, dy, srcs, 1, [L["c0"], L["c1"]], grads)
            Ho, cout = L["Ho"], L["cout"]
            ncol = L["c1"] if i == 3 else L["c0"] + L["c1"]
            dx = self._buf("fu%d.dx" % i, (L["M"], ncol), bf)
            ops.gemm(dz.view(Bn, Ho, Ho, cout), L["wd"], conv=dict(stride=1), out=dx)
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
            dz = self._bn_conv_backward("cs%d" % j, L, dyD, [L["src"]], 2, [L["cin"]], grads)
            if j > 0:
                Ho, cout = L["Ho"], L["cout"]
                u = self._buf("cs%d.u" % j, (Bn, 2 * Ho, 2 * Ho, cout), bf)
                ops.zero_insert2x(dz.view(Bn, Ho, Ho, cout), out=u)
                dxs = self._buf("cs%d.dx" % j, (Bn * 4 * Ho * Ho, L["cin"]), bf)
                ops.gemm(u, L["wd"], conv=dict(stride=1), out=dxs)
                dsum = self._buf("cs%d.dsum" % j, (Bn * 4 * Ho * Ho, L["cin"]), bf)
                ops.add_bf16(gskip[j], dxs, out=dsum)
                dyD = dsum
        return grads, dfmap
