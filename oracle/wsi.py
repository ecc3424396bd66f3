"""CPU restatement of the whole-slide plumbing around the generator — TEST INFRASTRUCTURE ONLY (numpy).

  otsu_threshold_u8     OpenCV's Otsu threshold for 8-bit images (cv2.threshold(..., THRESH_BINARY + THRESH_OTSU), the call at
                        /root/reference/slidevips-python/slidevips/tiling.py:30; cv2 is a third-party dependency of the
                        reference — opencv-python — its published algorithm getThreshVal_Otsu_8u is restated here and pinned
                        against the installed cv2 in tests/test_wsi_cpu.py).
  get_locs_otsu         slidevips/tiling.py:7-65, restated step by step (thumbnail -> std map -> Otsu mask -> float tile grid
                        -> tissue fraction per tile); pinned against the reference function itself (imported from
                        /root/reference with the installed cv2) and tests/golden/wsi_tiling.npz.
  order_tiles_horizontally  slidevips/tiling.py:68-84.
  stitch                preprocessings/cycle_gan/cycle_gan_wsi_inference.py:86-104: crop TILE_OVERLAP off each side, pyvips
                        insert (clipped, later tiles overwrite earlier ones). pyvips is absent here: this part is restated from
                        the source text and pyvips' documented insert semantics only — parity unpinned for stitch.
"""
import numpy as np


def std_u8(thumb):
    """np.uint8(thumbnail.std(axis=-1)) with numpy's operation order written out (tiling.py:27)."""
    x = thumb.astype(np.float64)
    s = x[..., 0].copy()
    for c in range(1, x.shape[-1]):
        s = s + x[..., c]
    mean = s / x.shape[-1]
    q = None
    for c in range(x.shape[-1]):
        d = x[..., c] - mean
        q = d * d if q is None else q + d * d
    return np.sqrt(q / x.shape[-1]).astype(np.uint8)


def otsu_threshold_u8(hist, n):
    scale = 1.0 / n
    mu = 0.0
    for i in range(256):
        mu += i * float(hist[i])
    mu *= scale
    mu1 = q1 = 0.0
    max_sigma, max_val = 0.0, 0
    eps = float(np.finfo(np.float32).eps)
    for i in range(256):
        p_i = float(hist[i]) * scale
        mu1 *= q1
        q1 += p_i
        q2 = 1.0 - q1
        if min(q1, q2) < eps or max(q1, q2) > 1.0 - eps:
            continue
        mu1 = (mu1 + i * p_i) / q1
        mu2 = (mu - q1 * mu1) / q2
        sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2)
        if sigma > max_sigma:
            max_sigma, max_val = sigma, i
    return max_val


def tissue_mask(thumbnail_or_mask):
    if thumbnail_or_mask.dtype == bool:
        return thumbnail_or_mask
    one = std_u8(thumbnail_or_mask) if thumbnail_or_mask.shape[-1] > 1 else thumbnail_or_mask[..., 0]
    t = otsu_threshold_u8(np.bincount(one.ravel(), minlength=256), one.size)
    return one > t


def tile_grid(mask_shape, slide_dim, tile_size_lvl0, tile_overlap):
    """the (thumbnail box, slide position) pairs get_locs_otsu visits, in its order (tiling.py:33-57)."""
    thumb_wh = np.array([mask_shape[1], mask_shape[0]])
    ratio = np.asarray(slide_dim) / thumb_wh
    size_t = tile_size_lvl0 / ratio
    over_t = tile_overlap / ratio
    ys_t = np.arange(0, thumb_wh[1] + 1, size_t[1] - over_t[1])
    ys = np.arange(0, slide_dim[1] + 1, tile_size_lvl0 - tile_overlap)
    xs_t = np.arange(0, thumb_wh[0] + 1, size_t[0] - over_t[0])
    xs = np.arange(0, slide_dim[0] + 1, tile_size_lvl0 - tile_overlap)
    out = []
    for yt, y in zip(ys_t, ys):
        for xt, x in zip(xs_t, xs):
            out.append((int(xt), int(yt), int(xt + size_t[0]), int(yt + size_t[1]), x, y))
    return out


def get_locs_otsu(thumbnail_or_mask, slide_dim, tile_size_lvl0, tile_overlap=0, mask_thresh=0.):
    mask = tissue_mask(thumbnail_or_mask)
    pos, frac = [], []
    for x0, y0, x1, y1, x, y in tile_grid(mask.shape[:2], slide_dim, tile_size_lvl0, tile_overlap):
        tile = mask[y0:y1, x0:x1]
        if tile.size == 0:
            continue
        p = np.count_nonzero(tile) / tile.size
        if p > mask_thresh:
            pos.append([x, y])
            frac.append(p)
    return np.asarray(pos), np.asarray(frac)


def order_tiles_horizontally(coordinates):
    s = coordinates[np.argsort(coordinates[:, 1])]
    s = s[np.lexsort((s[:, 0],))]
    return [int(np.where((coordinates == c).all(axis=1))[0][0]) for c in s]


def stitch(canvas, tiles, xy, crop, keep):
    """canvas uint8 [C, H, W] (modified in place); tiles uint8 [B, C, S, S]; xy [B, 2] canvas position of the kept window."""
    C, H, W = canvas.shape
    for b in range(tiles.shape[0]):
        x, y = int(xy[b][0]), int(xy[b][1])
        win = tiles[b, :, crop:crop + keep, crop:crop + keep]
        y0, y1, x0, x1 = max(y, 0), min(y + keep, H), max(x, 0), min(x + keep, W)
        if y1 <= y0 or x1 <= x0:
            continue
        canvas[:, y0:y1, x0:x1] = win[:, y0 - y:y1 - y, x0 - x:x1 - x]
    return canvas
