"""Whole-slide plumbing around the generator: which tiles of a slide to run, how loader workers hand
them to the GPU, and how the uint8 predictions are assembled into the output image.

  get_locs_otsu            mirror of slidevips.tiling.get_locs_otsu (slidevips-python/slidevips/tiling.py:7-65): same
                           arguments, same (tile_positions, tissue_percentages) return. The pixel work — per-pixel channel
                           std of the thumbnail, its histogram, the Otsu threshold, the tissue count of every tile box — runs in
                           mv_thumb_std_hist / mv_otsu_threshold / mv_tile_tissue (bit-exact integer results); the tile grid
                           itself (a few thousand float coordinates) is host arithmetic.
  order_tiles_horizontally mirror of slidevips.tiling.order_tiles_horizontally (tiling.py:68-84): pure index logic.
  PinnedTileRing           shared-memory ring of uint8 tile batches, page-locked in the consumer process: DataLoader worker
                           processes write tiles straight into pinned memory (no pickling of tile data, no staging copy;
                           src/dataset.py:545-575 / slidevips torch_datasets.py:88-127 produce the tiles in the reference) and
                           engine.infer_stream DMA-copies from it.
  TileStitcher             the overlap-crop-insert writer of preprocessings/cycle_gan/cycle_gan_wsi_inference.py:86-104 on
                           mv_stitch_tiles, into a device canvas or a mapped pinned host canvas (slides exceed HBM).
  infer_slide              tiles -> ring -> engine.infer_stream -> stitcher, sharded round-robin over ranks (BASELINE
                           configs[4]); independent tiles, no collective.
pyvips reading / OME-TIFF pyramid writing stay out of scope (pyvips is not part of this path's arithmetic).
"""
import ctypes

import numpy as np
import torch

from . import lib as _lib
from . import ops


# ------------------------------------------------------------------------------------------------ tile selection
def _tile_grid(mask_hw, slide_dim, tile_size_lvl0, tile_overlap):
    """Thumbnail boxes and level-0 positions in the order the reference visits them (rows of tiles, left to right)."""
    thumb_wh = np.array([mask_hw[1], mask_hw[0]])
    ratio = np.asarray(slide_dim) / thumb_wh                       # level-0 pixels per thumbnail pixel, (x, y)
    size_t, over_t = tile_size_lvl0 / ratio, tile_overlap / ratio  # tile extent / overlap in thumbnail pixels
    step0 = tile_size_lvl0 - tile_overlap
    ys_t = np.arange(0, thumb_wh[1] + 1, size_t[1] - over_t[1])
    xs_t = np.arange(0, thumb_wh[0] + 1, size_t[0] - over_t[0])
    ys = np.arange(0, slide_dim[1] + 1, step0)
    xs = np.arange(0, slide_dim[0] + 1, step0)
    ny, nx = min(len(ys_t), len(ys)), min(len(xs_t), len(xs))      # zip() stops at the shorter sequence
    boxes = np.empty((ny * nx, 4), dtype=np.int64)
    pos = []
    k = 0
    for j in range(ny):
        for i in range(nx):
            boxes[k] = (int(xs_t[i]), int(ys_t[j]), int(xs_t[i] + size_t[0]), int(ys_t[j] + size_t[1]))
            pos.append([xs[i], ys[j]])
            k += 1
    return boxes, pos


def get_locs_otsu(thumbnail_or_mask, slide_dim, tile_size_lvl0, tile_overlap=0, mask_thresh=0., device=None):
    """thumbnail uint8 [h, w, c] (or a boolean tissue mask [h, w]) -> (tile_positions [n, 2] level-0 (x, y),
    tissue_percentages [n]) of the tiles whose tissue fraction exceeds mask_thresh."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    t = torch.as_tensor(np.ascontiguousarray(thumbnail_or_mask)) if not torch.is_tensor(thumbnail_or_mask) else thumbnail_or_mask
    thresh = None
    if t.dtype == torch.bool:
        m = t.to(dev).to(torch.uint8).contiguous()                 # mask > 0
    else:
        if t.dim() == 2:
            t = t[..., None]
        if t.dtype != torch.uint8:
            raise TypeError("thumbnail must be uint8 (or a boolean mask), got %s" % t.dtype)
        m, hist = ops.thumb_std_hist(t.to(dev).contiguous())
        thresh = ops.otsu_threshold(hist, m.numel())
    H, W = m.shape
    boxes, pos = _tile_grid((H, W), slide_dim, tile_size_lvl0, tile_overlap)
    # Python slicing clips boxes to the map; empty boxes are skipped by the reference
    clip = boxes.copy()
    clip[:, [0, 2]] = np.clip(clip[:, [0, 2]], 0, W)
    clip[:, [1, 3]] = np.clip(clip[:, [1, 3]], 0, H)
    size = np.maximum(clip[:, 2] - clip[:, 0], 0) * np.maximum(clip[:, 3] - clip[:, 1], 0)
    keep = np.nonzero(size > 0)[0]
    if len(keep) == 0:
        return np.asarray([]), np.asarray([])
    bt = torch.from_numpy(clip[keep].astype(np.int32)).to(dev)
    counts = ops.tile_tissue(m, bt, thresh=thresh, fixed_thresh=0).cpu().numpy()
    positions, fractions = [], []
    for k, c in zip(keep, counts):
        p = int(c) / int(size[k])
        if p > mask_thresh:
            positions.append(pos[k])
            fractions.append(p)
    return np.asarray(positions), np.asarray(fractions)


def order_tiles_horizontally(coordinates):
    """indices that order tile coordinates the way the reference's writer expects them."""
    by_row = coordinates[np.argsort(coordinates[:, 1])]
    ordered = by_row[np.lexsort((by_row[:, 0],))]
    return [int(np.where((coordinates == c).all(axis=1))[0][0]) for c in ordered]


def shard_tiles(n_tiles, rank, world):
    """Round-robin tile assignment of a whole-slide sweep (independent units, no collective)."""
    return range(rank, n_tiles, world)


# ------------------------------------------------------------------------------------------------ input ring
class PinnedTileRing:
    """n_slots batches of raw uint8 NHWC tiles in ONE shared-memory block; page-locked (cudaHostRegister) in the process that
    feeds the GPU. Worker processes inherit the mapping (fork) and write tiles in place."""

    def __init__(self, n_slots, batch, size, channels=3, pin=None):
        self.n_slots, self.batch, self.size, self.channels = n_slots, batch, size, channels
        self.buf = torch.empty((n_slots, batch, size, size, channels), dtype=torch.uint8).share_memory_()
        self.pinned = False
        if pin is None:
            pin = torch.cuda.is_available()
        if pin:
            _lib.check(_lib.load().mv_host_register(ctypes.c_void_p(self.buf.data_ptr()), self.buf.numel()), "mv_host_register")
            self.pinned = True

    def slot(self, i):
        return self.buf[i % self.n_slots]

    def close(self):
        if self.pinned:
            _lib.check(_lib.load().mv_host_unregister(ctypes.c_void_p(self.buf.data_ptr())), "mv_host_unregister")
            self.pinned = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RingBatchDataset(torch.utils.data.Dataset):
    """Item k = batch k of `tiles` (any indexable of uint8 [S, S, C] arrays / tensors): the worker that gets it writes the
    tiles into ring slot k % n_slots and returns only (k, number of valid tiles). With a DataLoader(batch_size=None,
    num_workers=w, prefetch_factor=p) at most w * p batches are in flight, so n_slots >= w * p + consumer depth + 2 keeps
    producers off the slots the GPU is still reading."""

    def __init__(self, tiles, ring, indices=None):
        self.tiles, self.ring = tiles, ring
        self.indices = list(indices) if indices is not None else list(range(len(tiles)))

    def __len__(self):
        return (len(self.indices) + self.ring.batch - 1) // self.ring.batch

    def __getitem__(self, k):
        B = self.ring.batch
        ids = self.indices[k * B:(k + 1) * B]
        dst = self.ring.slot(k)
        for j, i in enumerate(ids):
            dst[j].copy_(torch.as_tensor(self.tiles[i]))
        if len(ids) < B:
            dst[len(ids):].zero_()
        return k, len(ids)


# ------------------------------------------------------------------------------------------------ output stitching
class TileStitcher:
    """canvas uint8 [channels, H, W], black like the reference's pyvips.Image.black; insert() crops `overlap` pixels off
    every side of each predicted tile and writes the rest at the tile's position (clipped to the canvas).

    positions are the canvas coordinates (x, y) of the UNCROPPED tile's top-left corner; the kept window lands at
    position + overlap. (The reference script adds the window size once more — `insert(tile, x + tile.height, y +
    tile.width)`, cycle_gan_wsi_inference.py:101-104 — which shifts the whole image by one tile; reference_shift=True
    reproduces that.)"""

    def __init__(self, canvas_hw, channels, tile_size, overlap=0, device=None, host=True, reference_shift=False):
        self.H, self.W = int(canvas_hw[0]), int(canvas_hw[1])
        self.C, self.S, self.overlap = int(channels), int(tile_size), int(overlap)
        self.keep = self.S - 2 * self.overlap
        if self.keep <= 0:
            raise ValueError("overlap %d leaves nothing of a %d-pixel tile" % (overlap, tile_size))
        self.shift = self.overlap + (self.keep if reference_shift else 0)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        ops.require_cuda(self.device, "TileStitcher")
        self._host = None
        if host:
            lib = _lib.init(self.device.index or 0)
            hp, dp = ctypes.c_void_p(), ctypes.c_void_p()
            n = self.C * self.H * self.W
            _lib.check(lib.mv_host_alloc_mapped(n, ctypes.byref(hp), ctypes.byref(dp)), "mv_host_alloc_mapped")
            self._host, self._ptr = hp.value, dp.value
            arr = np.ctypeslib.as_array(ctypes.cast(hp.value, ctypes.POINTER(ctypes.c_uint8)), shape=(n,))
            self.canvas = arr.reshape(self.C, self.H, self.W)
        else:
            self._dev = torch.zeros((self.C, self.H, self.W), dtype=torch.uint8, device=self.device)
            self._ptr = self._dev.data_ptr()
            self.canvas = self._dev

    def insert(self, tiles, positions, n_valid=None, sequential=False):
        """tiles uint8 device [B, C, S, S]; positions [B, 2] (x, y) ints (array-like). Asynchronous on the current stream."""
        B = tiles.shape[0] if n_valid is None else int(n_valid)
        if B == 0:
            return
        xy = torch.as_tensor(np.asarray(positions)[:B].astype(np.int64) + self.shift, dtype=torch.int32)
        ops.stitch_tiles(tiles[:B].contiguous(), xy.to(self.device, non_blocking=False).contiguous(), self.overlap, self.keep, self._ptr,
                         self.H, self.W, sequential=sequential)

    def result(self):
        """the finished canvas as a numpy array [C, H, W] (synchronises the device)."""
        torch.cuda.synchronize(self.device)
        return self.canvas if self._host is not None else self._dev.cpu().numpy()

    def close(self):
        if self._host is not None:
            torch.cuda.synchronize(self.device)
            self.canvas = None
            _lib.check(_lib.load().mv_host_free(ctypes.c_void_p(self._host)), "mv_host_free")
            self._host = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------ the sweep
def infer_slide(model, tiles, positions=None, batch=64, stitcher=None, rank=0, world=1, num_workers=0, prefetch_factor=2,
                n_slots=None, on_batch=None, stats=None):
    """Run the generator over the tiles of one slide (BASELINE configs[4]): `tiles` is any indexable of raw uint8 [S, S, 3]
    tiles (a Dataset reading regions of a slide in the reference); this rank takes tiles rank, rank + world, ...; loader
    workers fill a PinnedTileRing, engine.infer_stream normalises on the device and returns uint8 predictions, which go to
    `stitcher` (positions[i] = canvas (x, y) of tile i) and / or `on_batch(pred_u8_host, tile_indices)`.
    Returns the number of tiles processed by this rank; `stats` (a dict) receives the time to the first result and the total.
    The pinned ring is created once per (engine, batch, slots) and reused by later sweeps (page-locking ~150 MB is slow)."""
    import time
    t_start = time.perf_counter()
    eng = model.engine
    eng._ensure_packed()
    ids = list(shard_tiles(len(tiles), rank, world))
    depth = 2
    nw = int(num_workers)
    slots = n_slots if n_slots is not None else max(nw, 1) * max(prefetch_factor, 1) + depth + 2
    rings = eng.__dict__.setdefault("_tile_rings", {})
    ring = rings.get((slots, batch))
    if ring is None:
        ring = rings[(slots, batch)] = PinnedTileRing(slots, batch, eng.S)
    ds = RingBatchDataset(tiles, ring, ids)
    kw = dict(batch_size=None, shuffle=False, num_workers=nw)
    if nw > 0:
        kw.update(prefetch_factor=prefetch_factor, persistent_workers=False)
    loader = torch.utils.data.DataLoader(ds, **kw)
    meta = []

    def batches():
        for k, n in loader:
            meta.append((int(k), int(n)))
            yield ring.slot(int(k))

    done = 0

    def sink(dev_out, j):  # on the compute stream, right after batch j: predictions never leave the device un-stitched
        k, n = meta[j]
        stitcher.insert(dev_out, [positions[i] for i in ids[k * batch:k * batch + n]], n_valid=n)

    stream = eng.infer_stream(batches(), out_dtype=torch.uint8, depth=depth, device_sink=sink if stitcher is not None else None,
                              to_host=on_batch is not None or stitcher is None)
    t_first = None
    for j, pred in enumerate(stream):
        if t_first is None:
            t_first = time.perf_counter() - t_start
        k, n = meta[j]
        if on_batch is not None:
            on_batch(pred[:n], ids[k * batch:k * batch + n])
        done += n
    torch.cuda.synchronize(eng.device)
    if stats is not None:
        stats.update(t_first_s=t_first, t_total_s=time.perf_counter() - t_start, tiles=done, batches=len(meta))
    return done
