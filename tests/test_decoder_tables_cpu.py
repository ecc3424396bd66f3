"""Host logic of the training decoder without a GPU: the gather tables that replace per-step weight packing / gradient
unpacking (decoder_train.DecoderTrain) must reproduce the reference-layout packing functions exactly."""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import model as om  # noqa: E402


def _build(out_chans):
    from miphei_vit_b200.generators.mipheivit import get_vitmatte
    from miphei_vit_b200 import decoder_train as dtm

    cfg = om.Config(img_size=128, embed_dim=128, depth=1, num_heads=2, hidden=256, out_chans=out_chans)
    m = get_vitmatte("hoptimus0", cfg.img_size, cfg.out_chans, use_lora=True, embed_dim=cfg.embed_dim, depth=cfg.depth,
                     num_heads=cfg.num_heads, hidden=cfg.hidden)
    m.load_state_dict(om.init_state_dict(cfg, seed=3, perturb=True))
    eng = types.SimpleNamespace(model=m, device=torch.device("cpu"), D=cfg.embed_dim, heads_out=out_chans, flat_params=None)
    return m, dtm.DecoderTrain(eng), dtm


def _gather(src, idx):
    out = torch.zeros(idx.numel(), dtype=src.dtype)
    ok = idx >= 0
    out[ok] = src[idx[ok].long()]
    return out


def test_weight_tables_reproduce_the_packing_functions():
    from miphei_vit_b200 import packing

    for heads in (16, 3):
        m, dt, dtm = _build(heads)
        flat = torch.zeros(dt.n_dec)
        for _, p, off, n in dt.layout:
            flat[off:off + n] = p.detach().flatten()
        wf = {k: _gather(flat, dt.wf.idx)[o:o + v.numel()].view(v.shape) for (k, v), o in zip(dt.wf.v.items(), dt.wf.off.values())}
        wb = {k: _gather(flat, dt.wb.idx)[o:o + v.numel()].view(v.shape) for (k, v), o in zip(dt.wb.v.items(), dt.wb.off.values())}
        wc = {k: _gather(flat, dt.wc.idx)[o:o + v.numel()].view(v.shape) for (k, v), o in zip(dt.wc.v.items(), dt.wc.off.values())}
        dec = m.decoder
        for i, mod in enumerate(dec.convstream.convs):
            w = mod.conv.weight.detach()
            assert torch.equal(wf["cs%d.w" % i], packing.pack_conv3x3(w, [w.shape[1]], dtype=None))
            if i > 0:
                assert torch.equal(wb["cs%d.wd" % i], packing.pack_conv3x3(dtm._flip_t(w), [w.shape[0]], dtype=None))
            assert torch.equal(wc["cs%d.gamma" % i], mod.bn.weight.detach())
            assert torch.equal(wc["cs%d.beta" % i], mod.bn.bias.detach())
        for i, mod in enumerate(dec.fusion_blks):
            w = mod.conv.conv.weight.detach()
            c0 = dtm.SKIP_CH[i]
            assert torch.equal(wf["fu%d.w" % i], packing.pack_conv3x3(w, [c0, w.shape[1] - c0], dtype=None))
            wt = dtm._flip_t(w)
            assert torch.equal(wb["fu%d.wd" % i], packing.pack_conv3x3(wt if i < 3 else wt[c0:], [w.shape[0]], dtype=None))
        for h in range(heads):
            hd = getattr(dec, "segmentation_head_%d" % h)
            psi = hd[0].psi
            sl = slice(16 * h, 16 * h + 16)
            assert torch.equal(wc["hd.W1"][sl], psi[0].weight.detach().flatten(1))
            assert torch.equal(wf["hd.gate_w"][sl, :32], psi[0].weight.detach().flatten(1))
            assert torch.equal(wc["hd.b1"][sl], psi[0].bias.detach())
            assert torch.equal(wc["hd.gam"][sl], psi[1].weight.detach())
            assert torch.equal(wc["hd.bet"][sl], psi[1].bias.detach())
            assert torch.equal(wc["hd.w2"][sl], psi[3].weight.detach().flatten())
            assert wc["hd.b2"][h] == psi[3].bias.detach()[0] and wc["hd.b3"][h] == hd[1].bias.detach()[0]
            w3 = hd[1].weight.detach()[0]  # [32, 3, 3]
            assert torch.equal(wf["hd.conv_w"][h].view(9, 64)[:, :32], w3.permute(1, 2, 0).reshape(9, 32))
            for tap in range(9):
                assert torch.equal(wf["hd.w3t"][tap * 16 + h, :32], w3[:, tap // 3, tap % 3])
                assert torch.equal(wb["hd.w3tT"][:, tap * 16 + h], w3[:, tap // 3, tap % 3])
        assert float(wf["hd.gate_w"][:, 32:].abs().sum()) == 0.0 and float(wc["hd.W1"][16 * heads:].abs().sum()) == 0.0


def test_gradient_table_scatters_the_accumulators_into_parameter_layout():
    m, dt, dtm = _build(16)
    acc = torch.randn(dt.acc.data.numel())
    g = _gather(acc, dt.gidx)
    view = lambda k: acc[dt.acc.off[k]:dt.acc.off[k] + dt.acc.v[k].numel()].view(dt.acc.v[k].shape)  # noqa: E731
    got = {name: g[off:off + n].view(p.shape) for name, p, off, n in dt.layout}
    for L in dt.cs + dt.fu:
        wn, gn, bn = ("decoder." + x for x in L["names"])
        assert torch.equal(got[wn], dtm._unpack_wgrad(view(L["key"] + ".dwp"), L["cout"], L["splits"]))
        assert torch.equal(got[gn], view(L["key"] + ".sums")[1]) and torch.equal(got[bn], view(L["key"] + ".sums")[0])
    G3 = view("hd.G3")
    dW3 = G3[:32].reshape(32, 9, 16).permute(2, 0, 1).reshape(16, 32, 3, 3)
    for h in range(16):
        pre = "decoder.segmentation_head_%d." % h
        sl = slice(16 * h, 16 * h + 16)
        assert torch.equal(got[pre + "0.psi.0.weight"].flatten(1), view("hd.dW1")[sl])
        assert float(got[pre + "0.psi.0.bias"].abs().max()) == 0.0
        assert torch.equal(got[pre + "0.psi.1.weight"], view("hd.S2")[sl])
        assert torch.equal(got[pre + "0.psi.1.bias"], view("hd.S1")[sl])
        assert torch.equal(got[pre + "0.psi.3.weight"].flatten(), view("hd.dw2")[sl])
        assert got[pre + "0.psi.3.bias"][0] == view("hd.db2")[h] and got[pre + "1.bias"][0] == view("hd.db3")[h]
        assert torch.equal(got[pre + "1.weight"][0], dW3[h])
    # BatchNorm counters became views of one vector; state-dict keys and values are unchanged
    sd = m.state_dict()
    assert sd["decoder.fusion_blks.0.conv.bn.num_batches_tracked"].shape == ()
    assert len([k for k in sd if k.endswith("num_batches_tracked")]) == dt.nbt.numel()


def test_flat_params_and_decoder_layout_agree():
    from miphei_vit_b200.trainer import FlatParams

    m, dt, _ = _build(16)
    fp = FlatParams(m, torch.device("cpu"))
    assert fp.n_dec == dt.n_dec
    for name, p, off, n in dt.layout:
        assert p.data_ptr() == fp.flat[off:off + n].data_ptr() and p.grad.data_ptr() == fp.gflat[off:off + n].data_ptr()
