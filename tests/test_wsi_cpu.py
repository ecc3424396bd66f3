"""Oracle of the whole-slide plumbing (oracle/wsi.py) pinned against the reference's own tiling functions (imported from
/root/reference when present — they need only cv2 + numpy), against the installed cv2 for the Otsu threshold, and against
the committed goldens generated from the reference (tests/golden/make_wsi_golden.py); plus the host logic of the tile ring."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from oracle import wsi as ow  # noqa: E402
import make_wsi_golden as mg  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wsi_tiling.npz")


def test_oracle_matches_golden_from_reference():
    g = np.load(GOLDEN)
    for seed, hw, dim, ts, ov, th in mg.CASES:
        thumb = mg.synthetic_thumbnail(seed, hw)
        pos, frac = ow.get_locs_otsu(thumb, np.array(dim), ts, ov, th)
        assert np.array_equal(pos, g["pos%d" % seed]) and np.array_equal(frac, g["frac%d" % seed])
        assert ow.order_tiles_horizontally(pos) == g["order%d" % seed].tolist()
        pm, fm = ow.get_locs_otsu(thumb.std(axis=-1) > 20, np.array(dim), ts, ov, th)
        assert np.array_equal(pm, g["mpos%d" % seed]) and np.array_equal(fm, g["mfrac%d" % seed])


@pytest.mark.skipif(not os.path.exists(mg.REF), reason="/root/reference not present")
def test_oracle_matches_reference_functions():
    ref = mg.load_reference()
    for seed, hw, dim, ts, ov, th in [(11, (200, 150), (15000, 20000), 900, 60, 0.1), (12, (90, 130), (13000, 9000), 512, 0, 0.0)]:
        thumb = mg.synthetic_thumbnail(seed, hw)
        rp, rf = ref.get_locs_otsu(thumb, np.array(dim), ts, ov, th)
        op, of = ow.get_locs_otsu(thumb, np.array(dim), ts, ov, th)
        assert np.array_equal(rp, op) and np.array_equal(rf, of) and rp.dtype == op.dtype
        assert ref.order_tiles_horizontally(rp) == ow.order_tiles_horizontally(op)
        grey = thumb[..., :1]
        assert np.array_equal(ref.get_locs_otsu(grey, np.array(dim), ts, ov, th)[0], ow.get_locs_otsu(grey, np.array(dim), ts, ov, th)[0])


def test_otsu_restatement_matches_installed_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for t in range(200):
        h, w = int(rng.integers(5, 200)), int(rng.integers(5, 200))
        kind = t % 4
        if kind == 0:
            img = rng.integers(0, 256, (h, w)).astype(np.uint8)
        elif kind == 1:
            img = np.clip(rng.normal(40, 10, (h, w)), 0, 255).astype(np.uint8)
            img[rng.random((h, w)) < 0.3] = rng.integers(100, 200)
        elif kind == 2:
            img = ((rng.random((h, w)) < 0.5) * rng.integers(1, 255)).astype(np.uint8)
        else:
            img = np.full((h, w), rng.integers(0, 256), np.uint8)
        ref, _ = cv2.threshold(img, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
        assert ow.otsu_threshold_u8(np.bincount(img.ravel(), minlength=256), img.size) == int(ref)


def test_std_restatement_is_numpy_std_for_every_uint8_triple():
    a = np.arange(0, 256, 3, dtype=np.uint8)  # a 86^3 lattice + random triples (all 2^24 were checked once, offline)
    tri = np.stack(np.meshgrid(a, a, a, indexing="ij"), -1).reshape(-1, 1, 3)
    rnd = np.random.default_rng(1).integers(0, 256, (200000, 1, 3)).astype(np.uint8)
    for x in (tri, rnd):
        assert np.array_equal(ow.std_u8(x), np.uint8(x.std(axis=-1)))


def test_stitch_oracle_semantics():
    rng = np.random.default_rng(2)
    tiles = rng.integers(0, 256, (5, 2, 16, 16)).astype(np.uint8)
    canvas = np.zeros((2, 40, 50), np.uint8)
    xy = np.array([[0, 0], [12, 0], [44, 30], [-5, -5], [100, 100]])
    ow.stitch(canvas, tiles, xy, crop=2, keep=12)
    assert np.array_equal(canvas[:, 0:7, 0:7], tiles[3, :, 7:14, 7:14])          # clipped at the top-left, later tile wins
    assert np.array_equal(canvas[:, 0:12, 12:24], tiles[1, :, 2:14, 2:14])
    assert np.array_equal(canvas[:, 30:40, 44:50], tiles[2, :, 2:12, 2:8])        # clipped at the bottom-right
    assert canvas[:, 20:30, :40].sum() == 0


def test_tile_ring_dataset_host_logic():
    from miphei_vit_b200.wsi import PinnedTileRing, RingBatchDataset, shard_tiles

    S, n = 8, 23
    tiles = [np.full((S, S, 3), i, np.uint8) for i in range(n)]
    ring = PinnedTileRing(n_slots=6, batch=4, size=S, pin=False)
    ids = list(shard_tiles(n, 1, 2))                 # rank 1 of 2: tiles 1, 3, 5, ...
    ds = RingBatchDataset(tiles, ring, ids)
    assert len(ds) == 3
    loader = torch.utils.data.DataLoader(ds, batch_size=None, shuffle=False, num_workers=2, prefetch_factor=1)
    seen = []
    for k, nv in loader:
        slot = ring.slot(int(k))
        got = [int(slot[j, 0, 0, 0]) for j in range(int(nv))]
        assert got == ids[int(k) * 4:int(k) * 4 + int(nv)]   # written by a WORKER process, visible here (shared memory)
        if nv < 4:
            assert int(slot[int(nv):].sum()) == 0
        seen += got
    assert seen == ids
    assert sorted(list(shard_tiles(n, 0, 2)) + ids) == list(range(n))
