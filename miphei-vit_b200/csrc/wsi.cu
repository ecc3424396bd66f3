// Whole-slide plumbing around the generator: tissue-tile selection on a slide thumbnail and the
// overlap-crop-insert stitcher that assembles uint8 predictions into the output canvas.
//
//   mv_thumb_std_hist   get_locs_otsu, slidevips-python/slidevips/tiling.py:25-31: per-pixel np.uint8(thumbnail.std(-1)) and its
//                       256-bin histogram.  The float64 operation sequence of numpy's std (mean, centred squares summed in
//                       order, sqrt, truncation) is reproduced with explicitly rounded double intrinsics (no FMA
//                       contraction), so the uint8 map is bit-identical to the reference's.
//   mv_otsu_threshold   the threshold cv2.threshold(..., THRESH_BINARY + THRESH_OTSU) picks for an 8-bit image (OpenCV's
//                       getThreshVal_Otsu_8u: maximise q1 q2 (mu1 - mu2)^2 over the histogram, first maximum wins) — one thread,
//                       double precision, same update order.
//   mv_tile_tissue      tiling.py:52-60: number of mask pixels (value > threshold) inside each clipped thumbnail box.
//   mv_stitch_tiles     preprocessings/cycle_gan/cycle_gan_wsi_inference.py:98-104: crop `crop` pixels off every side of a
//                       predicted tile and insert the remaining `keep` x `keep` window into the canvas at (x, y), clipped to
//                       the canvas (pyvips insert semantics).  The canvas may live in device memory or in MAPPED pinned host
//                       memory (mv_host_alloc_mapped): slides are larger than HBM, the kernel then writes straight over PCIe.
// All HBM / PCIe-bound byte work: coalesced accesses, 16-byte vectors where alignment allows, no tensor cores.
#include "mv_host.h"
#include "mv_ptx.cuh"
#include <string.h>

namespace mv {

// thumb: uint8 [n_pix, C] (C = 1..4 interleaved); std_u8: uint8 [n_pix]; hist: uint32 [256] (+=)
__global__ void __launch_bounds__(256) thumb_std_hist_kernel(const uint8_t* __restrict__ thumb, long long n_pix, int C,
                                                             uint8_t* __restrict__ std_u8, unsigned int* __restrict__ hist) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  __shared__ unsigned int sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pix; i += (long long)gridDim.x * blockDim.x) {
    unsigned int v;
    if (C == 1) {
      v = thumb[i];
    } else {
      double x[4], s = 0.0;
      for (int c = 0; c < C; ++c) {
        x[c] = (double)thumb[i * C + c];
        s = __dadd_rn(s, x[c]);
      }
      const double mean = __ddiv_rn(s, (double)C);
      double q = 0.0;
      for (int c = 0; c < C; ++c) {
        const double d = __dsub_rn(x[c], mean);
        q = c == 0 ? __dmul_rn(d, d) : __dadd_rn(q, __dmul_rn(d, d));
      }
      v = (unsigned int)__dsqrt_rn(__ddiv_rn(q, (double)C));  // np.uint8(): truncation (std <= 127.5 for uint8 input)
    }
    std_u8[i] = (uint8_t)v;
    atomicAdd(&sh[v & 255u], 1u);
  }
  __syncthreads();
  if (sh[threadIdx.x]) atomicAdd(hist + threadIdx.x, sh[threadIdx.x]);
}

__global__ void otsu_threshold_kernel(const unsigned int* __restrict__ hist, double n_pix, int* __restrict__ thresh) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double scale = __ddiv_rn(1.0, n_pix);
  double mu = 0.0;
  for (int i = 0; i < 256; ++i) mu = __dadd_rn(mu, __dmul_rn((double)i, (double)hist[i]));
  mu = __dmul_rn(mu, scale);
  double mu1 = 0.0, q1 = 0.0, max_sigma = 0.0;
  int max_val = 0;
  const double eps = 1.1920928955078125e-07;  // FLT_EPSILON
  for (int i = 0; i < 256; ++i) {
    const double p_i = __dmul_rn((double)hist[i], scale);
    mu1 = __dmul_rn(mu1, q1);
    q1 = __dadd_rn(q1, p_i);
    const double q2 = __dsub_rn(1.0, q1);
    if (fmin(q1, q2) < eps || fmax(q1, q2) > 1.0 - eps) continue;
    mu1 = __ddiv_rn(__dadd_rn(mu1, __dmul_rn((double)i, p_i)), q1);
    const double mu2 = __ddiv_rn(__dsub_rn(mu, __dmul_rn(q1, mu1)), q2);
    const double d = __dsub_rn(mu1, mu2);
    const double sigma = __dmul_rn(__dmul_rn(__dmul_rn(q1, q2), d), d);
    if (sigma > max_sigma) {
      max_sigma = sigma;
      max_val = i;
    }
  }
  *thresh = max_val;
}

// boxes int32 [n, 4] = (x0, y0, x1, y1), already clipped to the map; counts[i] = #{mask > *thresh} inside box i.
// One CTA per box; mask rows are read as contiguous spans.
__global__ void __launch_bounds__(256) tile_tissue_kernel(const uint8_t* __restrict__ mask, int W, const int* __restrict__ thresh,
                                                          int fixed_thresh, const int* __restrict__ boxes,
                                                          int* __restrict__ counts) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  __shared__ int sh[8];
  const int t = thresh ? *thresh : fixed_thresh;
  const int* b = boxes + 4 * blockIdx.x;
  const int x0 = b[0], y0 = b[1], bw = b[2] - b[0], bh = b[3] - b[1];
  int acc = 0;
  const long long total = (long long)(bw > 0 ? bw : 0) * (bh > 0 ? bh : 0);
  for (long long i = threadIdx.x; i < total; i += blockDim.x) {
    const int yy = (int)(i / bw), xx = (int)(i - (long long)yy * bw);
    acc += mask[(long long)(y0 + yy) * W + x0 + xx] > t ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < 8; ++w) s += sh[w];
    counts[blockIdx.x] = s;
  }
}

// tiles uint8 [B, C, S, S]; xy int32 [B, 2] = canvas position of the kept window's top-left corner; canvas uint8 [C, H, W].
// grid = (row blocks, C, B); a thread copies up to 16 consecutive bytes of one row.
__global__ void __launch_bounds__(256) stitch_tiles_kernel(const uint8_t* __restrict__ tiles, const int* __restrict__ xy, int C,
                                                           int S, int crop, int keep, uint8_t* __restrict__ canvas,
                                                           long long H, long long W, int first_tile) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  const int b = first_tile + blockIdx.z, c = blockIdx.y;
  const int vec_per_row = (keep + 15) / 16;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= keep * vec_per_row) return;
  const int row = idx / vec_per_row, v = idx - row * vec_per_row;
  const long long cx = xy[2 * b], cy = (long long)xy[2 * b + 1] + row;
  if (cy < 0 || cy >= H) return;
  const uint8_t* src = tiles + (((long long)b * C + c) * S + crop + row) * S + crop + v * 16;
  uint8_t* dst = canvas + ((long long)c * H + cy) * W + cx + v * 16;
  const int n = min(16, keep - v * 16);
  const long long x_first = cx + v * 16;
  if (n == 16 && x_first >= 0 && x_first + 16 <= W && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
    return;
  }
  for (int j = 0; j < n; ++j)
    if (x_first + j >= 0 && x_first + j < W) dst[j] = src[j];
}

}  // namespace mv

extern "C" int mv_thumb_std_hist(const void* thumb, int64_t n_pix, int chans, void* std_u8, uint32_t* hist256, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(thumb && std_u8 && hist256 && n_pix > 0 && chans >= 1 && chans <= 4, "mv_thumb_std_hist: 1..4 interleaved channels");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  cudaError_t e = cudaMemsetAsync(hist256, 0, 256 * sizeof(uint32_t), stream);
  if (e != cudaSuccess) {
    set_error("cudaMemsetAsync: %s", cudaGetErrorString(e));
    return (int)e;
  }
  long long blocks = (n_pix + 255) / 256;
  const long long cap = (long long)(device_sms() > 0 ? device_sms() : 148) * 8;
  if (blocks > cap) blocks = cap;
  MV_LAUNCH(thumb_std_hist_kernel, (unsigned)blocks, 256, 0, stream, reinterpret_cast<const uint8_t*>(thumb), (long long)n_pix, chans,
            reinterpret_cast<uint8_t*>(std_u8), hist256);
  MV_CHECK_LAUNCH("thumb_std_hist");
  return MV_OK;
}

extern "C" int mv_otsu_threshold(const uint32_t* hist256, int64_t n_pix, int32_t* thresh, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(hist256 && thresh && n_pix > 0, "mv_otsu_threshold: null/empty");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MV_LAUNCH(otsu_threshold_kernel, 1, 32, 0, stream, hist256, (double)n_pix, thresh);
  MV_CHECK_LAUNCH("otsu_threshold");
  return MV_OK;
}

extern "C" int mv_tile_tissue(const void* mask_u8, int width, const int32_t* thresh_dev, int fixed_thresh, const int32_t* boxes,
                              int n_boxes, int32_t* counts, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(mask_u8 && boxes && counts && n_boxes > 0 && width > 0, "mv_tile_tissue: null/empty");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MV_LAUNCH(tile_tissue_kernel, n_boxes, 256, 0, stream, reinterpret_cast<const uint8_t*>(mask_u8), width, thresh_dev, fixed_thresh,
            boxes, counts);
  MV_CHECK_LAUNCH("tile_tissue");
  return MV_OK;
}

extern "C" int mv_stitch_tiles(const void* tiles_u8, const int32_t* xy, int batch, int chans, int size, int crop, int keep,
                               void* canvas_u8, int64_t canvas_h, int64_t canvas_w, int sequential, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(tiles_u8 && xy && canvas_u8 && batch > 0 && chans > 0, "mv_stitch_tiles: null/empty");
  MV_CHECK_ARG(crop >= 0 && keep > 0 && crop + keep <= size && canvas_h > 0 && canvas_w > 0, "mv_stitch_tiles: crop/keep/size");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int work = keep * ((keep + 15) / 16);
  if (sequential) {  // tiles that overlap each other after cropping: insert in index order, the later tile wins (pyvips insert)
    for (int b = 0; b < batch; ++b) {
      MV_LAUNCH(stitch_tiles_kernel, dim3((work + 255) / 256, chans, 1), 256, 0, stream, reinterpret_cast<const uint8_t*>(tiles_u8),
                xy, chans, size, crop, keep, reinterpret_cast<uint8_t*>(canvas_u8), (long long)canvas_h, (long long)canvas_w, b);
      MV_CHECK_LAUNCH("stitch_tiles");
    }
    return MV_OK;
  }
  MV_LAUNCH(stitch_tiles_kernel, dim3((work + 255) / 256, chans, batch), 256, 0, stream, reinterpret_cast<const uint8_t*>(tiles_u8), xy,
            chans, size, crop, keep, reinterpret_cast<uint8_t*>(canvas_u8), (long long)canvas_h, (long long)canvas_w, 0);
  MV_CHECK_LAUNCH("stitch_tiles");
  return MV_OK;
}

// Pinned host memory the kernels can address directly (output canvas of a slide; zero-initialised).
extern "C" int mv_host_alloc_mapped(int64_t bytes, void** host_ptr, void** device_ptr) {
  using namespace mv;
  MV_CHECK_ARG(bytes > 0 && host_ptr && device_ptr, "mv_host_alloc_mapped: null/empty");
  void* h = nullptr;
  cudaError_t e = cudaHostAlloc(&h, (size_t)bytes, cudaHostAllocMapped | cudaHostAllocPortable);
  if (e != cudaSuccess) {
    set_error("cudaHostAlloc(%lld): %s", (long long)bytes, cudaGetErrorString(e));
    return (int)e;
  }
  memset(h, 0, (size_t)bytes);
  void* d = nullptr;
  e = cudaHostGetDevicePointer(&d, h, 0);
  if (e != cudaSuccess) {
    cudaFreeHost(h);
    set_error("cudaHostGetDevicePointer: %s", cudaGetErrorString(e));
    return (int)e;
  }
  *host_ptr = h;
  *device_ptr = d;
  return MV_OK;
}

extern "C" int mv_host_free(void* host_ptr) {
  using namespace mv;
  cudaError_t e = cudaFreeHost(host_ptr);
  if (e != cudaSuccess) {
    set_error("cudaFreeHost: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return MV_OK;
}

// Page-lock an existing host range (shared memory that DataLoader worker processes fill: SURVEY 8f-2) so that H2D copies
// from it are asynchronous DMA transfers.
extern "C" int mv_host_register(void* ptr, int64_t bytes) {
  using namespace mv;
  MV_CHECK_ARG(ptr && bytes > 0, "mv_host_register: null/empty");
  cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("cudaHostRegister(%lld bytes): %s", (long long)bytes, cudaGetErrorString(e));
    return (int)e;
  }
  return MV_OK;
}

extern "C" int mv_host_unregister(void* ptr) {
  using namespace mv;
  cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("cudaHostUnregister: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return MV_OK;
}
