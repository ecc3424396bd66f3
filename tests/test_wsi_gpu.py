"""GPU parity of the whole-slide kernels (tile selection, stitching) against the numpy oracle and the reference-generated
goldens — integer / byte work: bit-exact — and of the end-to-end sweep (ring -> infer_stream -> stitcher)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from oracle import model as om  # noqa: E402
from oracle import wsi as ow  # noqa: E402
import make_wsi_golden as mg  # noqa: E402

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wsi_tiling.npz")


def test_std_map_histogram_and_otsu_threshold_are_bit_exact():
    from miphei_vit_b200 import ops
    for seed, hw in [(1, (300, 400)), (2, (257, 191)), (7, (1000, 1333))]:
        thumb = mg.synthetic_thumbnail(seed, hw)
        std, hist = ops.thumb_std_hist(torch.from_numpy(thumb).cuda())
        ref = ow.std_u8(thumb)
        assert np.array_equal(std.cpu().numpy(), ref)
        rh = np.bincount(ref.ravel(), minlength=256)
        assert np.array_equal(hist.cpu().numpy(), rh)
        assert int(ops.otsu_threshold(hist, ref.size).item()) == ow.otsu_threshold_u8(rh, ref.size)
    rnd = np.random.default_rng(3).integers(0, 256, (512, 512, 3)).astype(np.uint8)   # every kind of triple, not just images
    assert np.array_equal(ops.thumb_std_hist(torch.from_numpy(rnd).cuda())[0].cpu().numpy(), np.uint8(rnd.std(axis=-1)))
    for C in (1, 2, 4):
        x = np.random.default_rng(C).integers(0, 256, (64, 80, C)).astype(np.uint8)
        want = x[..., 0] if C == 1 else np.uint8(x.std(axis=-1))
        assert np.array_equal(ops.thumb_std_hist(torch.from_numpy(x).cuda())[0].cpu().numpy(), want)


def test_get_locs_otsu_matches_reference_golden_and_oracle():
    from miphei_vit_b200 import wsi
    g = np.load(GOLDEN)
    for seed, hw, dim, ts, ov, th in mg.CASES:
        thumb = mg.synthetic_thumbnail(seed, hw)
        pos, frac = wsi.get_locs_otsu(thumb, np.array(dim), ts, ov, th)
        assert np.array_equal(pos, g["pos%d" % seed]) and np.array_equal(frac, g["frac%d" % seed])
        assert pos.dtype == g["pos%d" % seed].dtype
        assert wsi.order_tiles_horizontally(pos) == g["order%d" % seed].tolist()
        pm, fm = wsi.get_locs_otsu(thumb.std(axis=-1) > 20, np.array(dim), ts, ov, th)   # boolean mask input
        assert np.array_equal(pm, g["mpos%d" % seed]) and np.array_equal(fm, g["mfrac%d" % seed])
    # edge cases: nothing passes the threshold; single-channel thumbnail
    blank = np.full((50, 60, 3), 200, np.uint8)
    p0, f0 = wsi.get_locs_otsu(blank, np.array((6000, 5000)), 512, 0, 0.0)
    o0 = ow.get_locs_otsu(blank, np.array((6000, 5000)), 512, 0, 0.0)
    assert len(p0) == len(o0[0]) == 0
    grey = mg.synthetic_thumbnail(9, (120, 90))[..., :1]
    assert np.array_equal(wsi.get_locs_otsu(grey, np.array((9000, 12000)), 700, 50, 0.05)[0],
                          ow.get_locs_otsu(grey, np.array((9000, 12000)), 700, 50, 0.05)[0])


@pytest.mark.parametrize("host", [True, False])
def test_stitcher_matches_oracle(host):
    from miphei_vit_b200.wsi import TileStitcher
    rng = np.random.default_rng(5)
    S, C, ov = 64, 5, 6
    keep = S - 2 * ov
    H, W = 3 * keep + 17, 4 * keep + 5
    tiles = rng.integers(0, 256, (12, C, S, S)).astype(np.uint8)
    pos = np.array([[(i % 4) * keep - ov, (i // 4) * keep - ov] for i in range(12)])   # abutting windows, last column / row clipped
    pos[11] = (W - 20, H - 9)
    st = TileStitcher((H, W), C, S, overlap=ov, host=host)
    st.insert(torch.from_numpy(tiles[:7]).cuda(), pos[:7])
    st.insert(torch.from_numpy(tiles[7:]).cuda(), pos[7:], n_valid=5)
    want = ow.stitch(np.zeros((C, H, W), np.uint8), tiles, pos + ov, ov, keep)
    assert np.array_equal(np.asarray(st.result()), want)
    st.close()
    # the reference script's extra shift by one window (cycle_gan_wsi_inference.py:101-104)
    st2 = TileStitcher((H, W), C, S, overlap=ov, host=host, reference_shift=True)
    st2.insert(torch.from_numpy(tiles).cuda(), pos, sequential=True)
    want2 = ow.stitch(np.zeros((C, H, W), np.uint8), tiles, pos + ov + keep, ov, keep)
    assert np.array_equal(np.asarray(st2.result()), want2)
    st2.close()


def test_infer_slide_ring_stream_stitch_end_to_end():
    """tiles of a synthetic slide -> PinnedTileRing (2 loader worker processes) -> infer_stream (uint8 tiles normalised on the
    device, uint8 sink) -> TileStitcher: the canvas equals the per-tile predictions of the plain call, cropped and placed by
    the oracle; sharding over 2 ranks covers every tile exactly once."""
    from miphei_vit_b200 import wsi
    from miphei_vit_b200.generators.mipheivit import get_vitmatte

    cfg = om.Config(img_size=128, embed_dim=128, depth=2, num_heads=2, hidden=256, out_chans=3)
    model = get_vitmatte("hoptimus0", cfg.img_size, cfg.out_chans, use_lora=True, embed_dim=cfg.embed_dim, depth=cfg.depth,
                         num_heads=cfg.num_heads, hidden=cfg.hidden)
    model.load_state_dict(om.init_state_dict(cfg, seed=4, perturb=True))
    model = model.cuda().eval()
    S, ov = 128, 8
    keep = S - 2 * ov
    nx, ny = 5, 3
    slide = np.random.default_rng(0).integers(0, 256, (ny * keep + 2 * ov, nx * keep + 2 * ov, 3)).astype(np.uint8)
    pos = [(i * keep, j * keep) for j in range(ny) for i in range(nx)]
    tiles = [np.ascontiguousarray(slide[y:y + S, x:x + S]) for x, y in pos]
    with torch.no_grad():
        ref = []
        for t in tiles:
            xn = om.normalize_tiles(torch.from_numpy(t).permute(2, 0, 1)[None])
            ref.append(model.engine.infer(xn.cuda(), out_dtype=torch.uint8).cpu().numpy()[0])
    ref = np.stack(ref)
    H, W = slide.shape[:2]
    canvases = []
    for rank in range(2):
        st = wsi.TileStitcher((H, W), cfg.out_chans, S, overlap=ov)
        got = {}
        n = wsi.infer_slide(model, tiles, positions=pos, batch=4, stitcher=st, rank=rank, world=2, num_workers=2,
                            on_batch=lambda p, ids: got.update({i: p[j].clone().numpy() for j, i in enumerate(ids)}))
        assert n == len(range(rank, len(tiles), 2)) and sorted(got) == list(range(rank, len(tiles), 2))
        for i, p in got.items():   # uint8-tile input path vs fp32-input path: normalisation rounding order only
            assert np.abs(p.astype(int) - ref[i].astype(int)).max() <= 2
        canvases.append(np.asarray(st.result()).copy())
        want = ow.stitch(np.zeros((cfg.out_chans, H, W), np.uint8), np.stack([got[i] for i in sorted(got)]),
                         np.array([pos[i] for i in sorted(got)]) + ov, ov, keep)
        assert np.array_equal(canvases[-1], want)
        st.close()
    both = np.maximum(canvases[0], canvases[1])      # disjoint windows: the union is the whole slide
    assert (both[:, ov:ov + ny * keep, ov:ov + nx * keep] > 0).mean() > 0.9
