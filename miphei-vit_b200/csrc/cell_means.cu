// Per-nucleus mean intensities of prediction and target maps — MeanCellExtrator.extract_mean, src/utils.py:49-121 (the
// evaluation step right behind the generator; CellMetrics.update, src/metrics.py:32-74, does the same reduction).
//
// The reference loops over the batch in Python: torch.unique(labels, return_inverse) + 3 scatter_add_ per image, with a
// host sync per image. Here ONE CTA owns one image and keeps everything in shared memory:
//   pass A  distinct positive labels -> open-addressing hash table (64-bit atomicCAS)
//           occupied slots -> bitonic sort by label (torch.unique returns ascending ids) -> slot -> rank
//   pass B  every labelled pixel adds its C prediction / C target values and a count to row `rank` of a [cap, 2C+1] fp32
//           accumulator: each run of equal label inside a warp is reduced by a shuffle reduce-scatter, then ONE warp-wide
//           shared-memory atomic updates its 32 sums
//   out     means = sums / count, ids, counts per image (padded to `cap` rows) + the number of nuclei of the image
// mv_cell_means_pack concatenates the per-image rows in batch order (the reference's torch.cat).
// HBM-bound: reads (2C * 4 + label) bytes per pixel once (the label map twice, from L2).
#include "mv_host.h"
#include "mv_ptx.cuh"

namespace mv {

constexpr int CM_THREADS = 512;
constexpr long long CM_PAD_KEY = 0x7fffffffffffffffll;

template <typename LabelT>
__global__ void __launch_bounds__(CM_THREADS) cell_means_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                                const LabelT* __restrict__ nuclei, int C, int HW, int cap,
                                                                float* __restrict__ means_pred, float* __restrict__ means_target,
                                                                long long* __restrict__ ids, float* __restrict__ counts,
                                                                int* __restrict__ n_unique, int* __restrict__ overflow,
                                                                uint8_t* __restrict__ gws, long long gws_stride,
                                                                float* __restrict__ gsum, int* __restrict__ gcnt) {
  griddep_sync();
  extern __shared__ __align__(16) uint8_t cm_smem[];
  const int hslots = 2 * cap;  // power of two
  // tables and accumulators: shared memory when they fit (the fast path), else this image's slice of a caller-provided
  // global workspace (images with thousands of nuclei: same algorithm, the atomics go to L2)
  uint8_t* base = gws ? gws + (long long)blockIdx.y * gws_stride : cm_smem;
  long long* keys = reinterpret_cast<long long*>(base);                 // [hslots] 0 = empty
  long long* skey = keys + hslots;                                      // [cap] sort keys
  int* dense = reinterpret_cast<int*>(skey + cap);                      // [hslots] slot -> rank
  int* sslot = dense + hslots;                                          // [cap] sort payload (hash slot)
  float* acc = reinterpret_cast<float*>(sslot + cap);                   // [cap][2C + 1]
  __shared__ int s_count, s_over, s_last;
  // grid = (slices, images): every CTA of an image builds the same label table (labels are read by all of them — 1/17 of the
  // traffic at 16 channels), then reduces only its slice of the pixels; partial sums meet in gsum, the last CTA finishes
  const int b = blockIdx.y, tid = threadIdx.x;
  const int slices = gridDim.x;
  const int per_slice = ((HW + slices - 1) / slices + 31) & ~31;
  const int p_lo = blockIdx.x * per_slice, p_hi = min(HW, p_lo + per_slice);
  const int W = 2 * C + 1;
  const LabelT* lab = nuclei + (long long)b * HW;
  const float* pb = pred + (long long)b * C * HW;
  const float* tb = target ? target + (long long)b * C * HW : nullptr;

  for (int i = tid; i < hslots; i += CM_THREADS) { keys[i] = 0; dense[i] = -1; }
  for (int i = tid; i < cap; i += CM_THREADS) { skey[i] = CM_PAD_KEY; sslot[i] = -1; }
  for (int i = tid; i < cap * W; i += CM_THREADS) acc[i] = 0.f;
  if (tid == 0) { s_count = 0; s_over = 0; }
  __syncthreads();

  auto hash = [&](long long k) {
    unsigned long long x = (unsigned long long)k * 0x9E3779B97F4A7C15ull;
    return (int)(x >> 40) & (hslots - 1);
  };
  // ---- pass A: insert distinct labels (4 label loads in flight per thread)
  auto insert = [&](long long k) {
    if (k <= 0) return;
    int slot = hash(k), probes = 0;
    while (true) {
      // plain read first: after its first pixel a nucleus is found here without an atomic (64-bit shared-memory
      // compare-and-swap is slow, and neighbouring lanes carry the same label)
      const long long cur = *reinterpret_cast<volatile long long*>(&keys[slot]);
      if (cur == k) break;
      if (cur == 0) {
        const unsigned long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(&keys[slot]), 0ull, (unsigned long long)k);
        if (prev == 0ull || prev == (unsigned long long)k) break;
      }
      slot = (slot + 1) & (hslots - 1);
      if (++probes >= hslots) { s_over = 1; break; }
    }
  };
  for (int p = tid; p < HW; p += 4 * CM_THREADS) {
    long long k4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) k4[u] = p + u * CM_THREADS < HW ? (long long)lab[p + u * CM_THREADS] : 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) insert(k4[u]);
  }
  __syncthreads();
  // ---- occupied slots -> sort buffer
  for (int i = tid; i < hslots; i += CM_THREADS) {
    if (keys[i] != 0) {
      const int j = atomicAdd(&s_count, 1);
      if (j < cap) { skey[j] = keys[i]; sslot[j] = i; } else s_over = 1;
    }
  }
  __syncthreads();
  const int U = s_count;
  if (s_over || U > cap) {
    if (tid == 0) { atomicExch(overflow, 1); n_unique[b] = 0; }
    return;
  }
  // ---- bitonic sort of (label, slot) by label over `cap` entries (padding sorts last)
  for (int k = 2; k <= cap; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < cap; i += CM_THREADS) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const bool up = (i & k) == 0;
          const long long a = skey[i], c = skey[ixj];
          if ((a > c) == up) {
            skey[i] = c; skey[ixj] = a;
            const int t = sslot[i]; sslot[i] = sslot[ixj]; sslot[ixj] = t;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int r = tid; r < U; r += CM_THREADS) dense[sslot[r]] = r;
  __syncthreads();
  // ---- pass B: accumulate (loop trip count is warp-uniform: __match_any_sync needs every lane)
  const int lane = tid & 31;
  for (int p0 = p_lo + (tid & ~31); p0 < p_hi; p0 += CM_THREADS) {
    const int p = p0 + lane;
    int r = -1;
    if (p < p_hi) {
      const long long k = (long long)lab[p];
      if (k > 0) {
        int slot = hash(k);
        while (keys[slot] != k) slot = (slot + 1) & (hslots - 1);
        r = dense[slot];
      }
    }
    // Runs of equal rank along the 32 consecutive pixels of this warp are reduced one at a time by a shuffle
    // REDUCE-SCATTER (31 exchanges): lane L ends up with the run's sum of value L (16 prediction + 16 target channels), so
    // the 32 accumulator updates of a run are ONE warp-wide shared-memory atomic on 32 different addresses.  (fp32 shared
    // atomics are compare-and-swap loops: 33 serial ones per run, or same-address collisions, cost milliseconds.)
    const int r_prev = __shfl_up_sync(0xffffffffu, r, 1);
    const bool head = lane == 0 || r != r_prev;
    unsigned heads = __ballot_sync(0xffffffffu, head);
    const unsigned labelled = __ballot_sync(0xffffffffu, r >= 0);
    if (labelled == 0u) continue;  // whole warp on background
    for (int c0 = 0; c0 < C; c0 += 16) {
      float v[32];  // this pixel's values: [0,16) prediction channels c0.., [16,32) target channels c0..
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int c = c0 + j;
        const bool ok = r >= 0 && c < C;
        v[j] = ok ? pb[(long long)c * HW + p] : 0.f;
        v[16 + j] = ok && tb ? tb[(long long)c * HW + p] : 0.f;
      }
      unsigned todo = heads;
      while (todo) {  // warp-uniform loop over the runs of this warp
        const int h = __ffs(todo) - 1;
        todo &= todo - 1;
        const int e = todo ? __ffs(todo) - 1 : 32;  // run = lanes [h, e)
        const int rr = __shfl_sync(0xffffffffu, r, h);
        if (rr < 0) continue;  // background run
        const bool mine = lane >= h && lane < e;
        float a[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) a[j] = mine ? v[j] : 0.f;
#pragma unroll
        for (int s2 = 16; s2 >= 1; s2 >>= 1) {
          const bool up = (lane & s2) != 0;
#pragma unroll
          for (int i = 0; i < s2; ++i) {
            const float keep = up ? a[i + s2] : a[i];
            const float send = up ? a[i] : a[i + s2];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, s2);
          }
        }
        // lane L holds the run total of value L
        const int c = c0 + (lane & 15);
        if (c < C && (lane < 16 || tb)) atomicAdd(acc + rr * W + (lane < 16 ? c : C + c), a[0]);
        if (c0 == 0 && lane == 0) atomicAdd(acc + rr * W + 2 * C, (float)(e - h));
      }
    }
  }
  __syncthreads();
  if (slices > 1) {
    // partial sums of this slice -> global; the CTA whose counter increment is the last one reads the totals back
    float* gs = gsum + (long long)b * cap * W;
    for (int i = tid; i < U * W; i += CM_THREADS) {
      const float v = acc[i];
      if (v != 0.f) atomicAdd(gs + i, v);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(gcnt + b, 1) == slices - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int i = tid; i < U * W; i += CM_THREADS) acc[i] = __ldcg(gs + i);
    __syncthreads();
  }
  // ---- means of this image (rows 0..U-1, ascending label)
  for (int i = tid; i < U * C; i += CM_THREADS) {
    const int r = i / C, c = i - r * C;
    const float n = acc[r * W + 2 * C];
    means_pred[((long long)b * cap + r) * C + c] = acc[r * W + c] / n;
    if (means_target) means_target[((long long)b * cap + r) * C + c] = tb ? acc[r * W + C + c] / n : 0.f;
  }
  for (int r = tid; r < U; r += CM_THREADS) {
    ids[(long long)b * cap + r] = skey[r];
    counts[(long long)b * cap + r] = acc[r * W + 2 * C];
  }
  if (tid == 0) n_unique[b] = U;
}

// rows of image b go to [offset_b, offset_b + n_unique[b]) with offset_b = sum of the earlier images' counts
__global__ void cell_means_pack_kernel(const float* __restrict__ means_pred, const float* __restrict__ means_target,
                                       const long long* __restrict__ ids, const float* __restrict__ counts,
                                       const int* __restrict__ n_unique, int C, int cap, float* __restrict__ out_pred,
                                       float* __restrict__ out_target, long long* __restrict__ out_ids,
                                       float* __restrict__ out_counts) {
  griddep_sync();
  const int b = blockIdx.x;
  long long off = 0;
  for (int i = 0; i < b; ++i) off += n_unique[i];
  const int U = n_unique[b];
  for (int i = threadIdx.x; i < U * C; i += blockDim.x) {
    out_pred[off * C + i] = means_pred[(long long)b * cap * C + i];
    if (out_target) out_target[off * C + i] = means_target[(long long)b * cap * C + i];
  }
  for (int r = threadIdx.x; r < U; r += blockDim.x) {
    out_ids[off + r] = ids[(long long)b * cap + r];
    if (out_counts) out_counts[off + r] = counts[(long long)b * cap + r];
  }
}

static size_t cm_smem_bytes(int C, int cap) {
  return (size_t)(2 * cap) * 8 + (size_t)cap * 8 + (size_t)(2 * cap) * 4 + (size_t)cap * 4 + (size_t)cap * (2 * C + 1) * 4;
}
constexpr size_t CM_SMEM_MAX = 220 * 1024;
static int cm_slices(int batch) {
  const int sms = device_sms() > 0 ? device_sms() : 148;
  int s = (2 * sms + batch - 1) / batch;
  return s < 1 ? 1 : (s > 8 ? 8 : s);
}
static size_t cm_gsum_bytes(int batch, int C, int cap) { return ((size_t)batch * cap * (2 * C + 1) * 4 + 255) / 256 * 256; }

// Backward of the per-nucleus means: d pred[b, c, p] = d means[row(b, label p), c] / count[row] for labelled pixels, 0 on
// background. ids / counts / n_unique are the forward's packed outputs (ids ascending per image): one binary search per pixel.
template <typename LabelT>
__global__ void cell_means_bwd_kernel(const float* __restrict__ dmeans, const long long* __restrict__ ids,
                                      const float* __restrict__ counts, const int* __restrict__ n_unique,
                                      const LabelT* __restrict__ nuclei, int B, int C, int HW, float* __restrict__ dmap) {
  griddep_sync();
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  long long off = 0;
  for (int i = 0; i < b; ++i) off += n_unique[i];
  const int U = n_unique[b];
  const long long k = (long long)nuclei[(long long)b * HW + p];
  long long row = -1;
  if (k > 0 && U > 0) {
    int lo = 0, hi = U - 1;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (ids[off + mid] < k) lo = mid + 1; else hi = mid;
    }
    if (ids[off + lo] == k) row = off + lo;
  }
  const float inv = row >= 0 ? 1.f / counts[row] : 0.f;
  float* o = dmap + (long long)b * C * HW + p;
  for (int c = 0; c < C; ++c) o[(long long)c * HW] = row >= 0 ? dmeans[row * C + c] * inv : 0.f;
}

}  // namespace mv

extern "C" int mv_cell_means(const float* pred, const float* target, const void* nuclei, int label_bytes, int batch, int chans,
                             int hw, int cap, float* means_pred, float* means_target, int64_t* ids, float* counts,
                             int32_t* n_unique, int32_t* overflow, void* workspace, int64_t workspace_bytes, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(pred && nuclei && means_pred && ids && counts && n_unique && overflow && batch > 0 && chans > 0 && hw > 0,
               "mv_cell_means: null/empty");
  MV_CHECK_ARG(label_bytes == 4 || label_bytes == 8, "mv_cell_means: labels must be int32 or int64");
  MV_CHECK_ARG(cap >= 32 && (cap & (cap - 1)) == 0, "mv_cell_means: cap must be a power of two >= 32 (got %d)", cap);
  MV_CHECK_ARG((target == nullptr) == (means_target == nullptr), "mv_cell_means: target and means_target go together");
  size_t smem = cm_smem_bytes(chans, cap);
  const long long stride = (long long)((smem + 255) / 256 * 256);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  uint8_t* gws = nullptr;
  float* gsum = nullptr;
  int* gcnt = nullptr;
  int slices = 1;
  const long long need = mv_cell_means_workspace_bytes(batch, chans, cap);
  if (smem > CM_SMEM_MAX) {  // tables in global memory (one CTA per image)
    MV_CHECK_ARG(workspace && workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
                 "mv_cell_means: cap %d x %d channels exceeds shared memory: pass a 256-byte aligned workspace of "
                 "mv_cell_means_workspace_bytes() = %lld bytes", cap, chans, need);
    gws = reinterpret_cast<uint8_t*>(workspace);
    smem = 0;
  } else if (workspace && workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0) {
    // several CTAs per image: enough of them to cover the machine about twice
    slices = cm_slices(batch);
    if (slices > 1) {
      gsum = reinterpret_cast<float*>(workspace);
      gcnt = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(workspace) + cm_gsum_bytes(batch, chans, cap));
      cudaError_t em = cudaMemsetAsync(workspace, 0, (size_t)need, stream);
      if (em != cudaSuccess) {
        set_error("cudaMemsetAsync: %s", cudaGetErrorString(em));
        return (int)em;
      }
    }
  }
  cudaError_t e = cudaSuccess;
  if (label_bytes == 4) {
    e = cudaFuncSetAttribute(cell_means_kernel<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      MV_LAUNCH(cell_means_kernel<int32_t>, dim3(slices, batch), CM_THREADS, smem, stream, pred, target,
                reinterpret_cast<const int32_t*>(nuclei), chans, hw, cap, means_pred, means_target, reinterpret_cast<long long*>(ids),
                counts, n_unique, overflow, gws, stride, gsum, gcnt);
  } else {
    e = cudaFuncSetAttribute(cell_means_kernel<long long>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      MV_LAUNCH(cell_means_kernel<long long>, dim3(slices, batch), CM_THREADS, smem, stream, pred, target,
                reinterpret_cast<const long long*>(nuclei), chans, hw, cap, means_pred, means_target,
                reinterpret_cast<long long*>(ids), counts, n_unique, overflow, gws, stride, gsum, gcnt);
  }
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(cell_means): %s", cudaGetErrorString(e));
    return (int)e;
  }
  MV_CHECK_LAUNCH("cell_means");
  return MV_OK;
}

extern "C" int mv_cell_means_pack(const float* means_pred, const float* means_target, const int64_t* ids, const float* counts,
                                  const int32_t* n_unique, int batch, int chans, int cap, float* out_pred, float* out_target,
                                  int64_t* out_ids, float* out_counts, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(means_pred && ids && n_unique && out_pred && out_ids && batch > 0 && chans > 0 && cap > 0, "mv_cell_means_pack: null/empty");
  MV_CHECK_ARG((means_target == nullptr) == (out_target == nullptr), "mv_cell_means_pack: target buffers go together");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MV_LAUNCH(cell_means_pack_kernel, batch, 256, 0, stream, means_pred, means_target, reinterpret_cast<const long long*>(ids), counts,
            n_unique, chans, cap, out_pred, out_target, reinterpret_cast<long long*>(out_ids), out_counts);
  MV_CHECK_LAUNCH("cell_means_pack");
  return MV_OK;
}

// bytes of global workspace mv_cell_means wants for this (batch, chans, cap): partial sums for several CTAs per image while
// the tables fit in shared memory (optional: without it one CTA per image runs), the tables themselves beyond that (required)
extern "C" int64_t mv_cell_means_workspace_bytes(int batch, int chans, int cap) {
  const size_t smem = mv::cm_smem_bytes(chans, cap);
  if (smem <= mv::CM_SMEM_MAX)  // shared-memory tables: workspace = cross-CTA partial sums + per-image counters
    return mv::cm_slices(batch) > 1 ? (int64_t)(mv::cm_gsum_bytes(batch, chans, cap) + ((size_t)batch * 4 + 255) / 256 * 256) : 0;
  return (int64_t)((smem + 255) / 256 * 256) * batch;
}

extern "C" int mv_cell_means_bwd(const float* dmeans, const int64_t* ids, const float* counts, const int32_t* n_unique,
                                 const void* nuclei, int label_bytes, int batch, int chans, int hw, float* dmap,
                                 void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(ids && counts && n_unique && nuclei && dmap && batch > 0 && chans > 0 && hw > 0, "mv_cell_means_bwd: null/empty");
  MV_CHECK_ARG(label_bytes == 4 || label_bytes == 8, "mv_cell_means_bwd: labels must be int32 or int64");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  dim3 grid((hw + 255) / 256, batch);
  if (label_bytes == 4)
    MV_LAUNCH(cell_means_bwd_kernel<int32_t>, grid, 256, 0, stream, dmeans, reinterpret_cast<const long long*>(ids), counts,
              n_unique, reinterpret_cast<const int32_t*>(nuclei), batch, chans, hw, dmap);
  else
    MV_LAUNCH(cell_means_bwd_kernel<long long>, grid, 256, 0, stream, dmeans, reinterpret_cast<const long long*>(ids), counts,
              n_unique, reinterpret_cast<const long long*>(nuclei), batch, chans, hw, dmap);
  MV_CHECK_LAUNCH("cell_means_bwd");
  return MV_OK;
}
