"""Generates tests/golden/wsi_tiling.npz by running the REFERENCE's own get_locs_otsu / order_tiles_horizontally
(/root/reference/slidevips-python/slidevips/tiling.py, imported as a file; needs cv2 + numpy only) on seeded synthetic
thumbnails.  Run in the build container (the reference tree does not exist on the GPU box):
    python tests/golden/make_wsi_golden.py
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = os.path.join(os.environ.get("MIPHEI_REFERENCE", "/root/reference"), "slidevips-python", "slidevips", "tiling.py")

CASES = [  # seed, thumbnail (h, w), slide_dim (W, H), tile_size_lvl0, overlap, mask_thresh
    (1, (300, 400), (40000, 30000), 2048, 0, 0.0),
    (2, (257, 191), (19100, 25700), 1024, 128, 0.01),
    (3, (512, 512), (100000, 90000), 4096, 256, 0.5),
    (4, (64, 96), (960, 640), 256, 32, 0.0),
    (5, (333, 777), (77700, 33300), 3000, 100, 0.25),
]


def synthetic_thumbnail(seed, hw):
    """white-ish background with coloured tissue blobs and noise, uint8 [h, w, 3]"""
    g = np.random.default_rng(seed)
    h, w = hw
    img = np.clip(g.normal(235, 4, (h, w, 3)), 0, 255)
    yy, xx = np.mgrid[0:h, 0:w]
    for _ in range(6):
        cy, cx, r = g.integers(0, h), g.integers(0, w), g.integers(min(h, w) // 10, min(h, w) // 3)
        blob = (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
        colour = np.array([g.integers(120, 220), g.integers(40, 160), g.integers(120, 220)])
        img[blob] = np.clip(colour + g.normal(0, 18, (int(blob.sum()), 3)), 0, 255)
    return img.astype(np.uint8)


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_tiling", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference()
    out = {}
    for seed, hw, dim, ts, ov, th in CASES:
        thumb = synthetic_thumbnail(seed, hw)
        pos, frac = ref.get_locs_otsu(thumb, np.array(dim), ts, ov, th)
        out["pos%d" % seed], out["frac%d" % seed] = pos, frac
        out["order%d" % seed] = np.asarray(ref.order_tiles_horizontally(pos)) if len(pos) else np.zeros(0, dtype=np.int64)
        mask = thumb.std(axis=-1) > 20
        pm, fm = ref.get_locs_otsu(mask, np.array(dim), ts, ov, th)
        out["mpos%d" % seed], out["mfrac%d" % seed] = pm, fm
        print(seed, hw, "tiles", len(pos), "mask tiles", len(pm))
    np.savez_compressed(os.path.join(HERE, "wsi_tiling.npz"), **out)


if __name__ == "__main__":
    main()
