// bf16 tensor-core GEMM for sm_100a:  D = epilogue(A[M,K] . B[N,K]^T)
//
//   * persistent, one CTA per SM, warp-specialised: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (+ TMEM owner),
//     warps 2..5 = epilogue (TMEM -> registers -> global)
//   * operands: TMA 128-byte-swizzled K-major tiles (BLOCK_K = 64 bf16 = 128 B rows), STAGES-deep mbarrier ring
//   * accumulators: fp32 in TMEM, two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1
//   * M tails: TMA zero-fills out-of-range rows, the epilogue masks the stores; N tails masked per 8 columns
//
// Reference ops replaced: every nn.Linear of the timm ViT created at src/generators/foundation_models.py:53-57,
// QkvWithLoRA (src/generators/lora.py:29-33) through a K-extended weight, and the decoder convs through im2col.
#include "mv_host.h"
#include "mv_ptx.cuh"
#include <stdlib.h>

namespace mv {

// internal mode (never passed through the C ABI): LINEAR with fp32 output + fp32 residual whose epilogue moves the residual in
// and the result out with TMA (bulk async copies through swizzled shared-memory tiles) instead of per-thread ld / st
#ifndef MV_TMA_RES_STAGES
#define MV_TMA_RES_STAGES 2        // residual tiles in flight per epilogue warp (fp32-residual epilogue)
#endif
#ifndef MV_TMA_OUT_STAGES_F32
#define MV_TMA_OUT_STAGES_F32 1    // output staging tiles per epilogue warp, fp32-residual epilogue (5 operand stages remain)
#endif
#ifndef MV_TMA_OUT_STAGES_BF16
#define MV_TMA_OUT_STAGES_BF16 1   // output staging tiles per epilogue warp, bf16 epilogue (6 operand stages remain; 2 -> 5: measured equal)
#endif
constexpr int MV_GEMM_LINEAR_TMA = 16;
// internal mode: LINEAR with bf16 output, no residual (QKV, the dX GEMMs): accumulator -> scale / shift (/ ReLU) -> bf16 ->
// swizzled shared-memory tile -> one TMA store per 32 rows x 64 columns; no transpose staging, no per-thread global store
constexpr int MV_GEMM_LINEAR_TMA_BF16 = 17;
// the same epilogue with EIGHT epilogue warps (two per TMEM lane quadrant, every other 64-column chunk each) and the
// GATE_MASK activation: problems of one or two K blocks over millions of rows (the decoder's 32-channel maps at 256^2) are
// bound by how fast the accumulators leave TMEM.  The transpose-staging epilogue of the 128-wide "light" tiles executed
// ~2500 warp instructions per 128 x 128 tile on 2 warps per scheduler (ncu source page, profiles/r02_skinny_k_gemm.txt):
// e = mask(f W1^T) [2M x 256] took 1130 us for 1 GB of output.
constexpr int MV_GEMM_LINEAR_TMA_BF16_W8 = 18;
__host__ __device__ constexpr bool gemm_tma_bf16(int mode) {
  return mode == MV_GEMM_LINEAR_TMA_BF16 || mode == MV_GEMM_LINEAR_TMA_BF16_W8;
}
constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;
constexpr int GEMM_THREADS = 192;
// the SwiGLU epilogues (exp per element, extra loads / stores) need twice the epilogue warps to keep up with the MMAs:
// two warps per TMEM lane quadrant, each taking every other 32-column chunk
// Role-stall counters (mv_gemm_set_profile_buffer) are compiled in only with -DMV_GEMM_PROFILE=1 (the diagnostic library
// libmiphei_b200_prof.so built by `python miphei-vit_b200/build.py --prof`); the production kernels carry none of it.
#ifndef MV_GEMM_PROFILE
#define MV_GEMM_PROFILE 0
#endif
constexpr bool kProf = MV_GEMM_PROFILE != 0;
#ifndef MV_LINEAR_EPI_WARPS
#define MV_LINEAR_EPI_WARPS 4  // 8 spills on the fp32 + residual path (204-register cap) and loses 30-50 % there
#endif
__host__ __device__ constexpr int gemm_epi_warps(int mode, int block_n, bool light) {
  return (mode == MV_GEMM_SWIGLU || mode == MV_GEMM_SWIGLU_BWD || mode == MV_GEMM_LINEAR_TMA_BF16_W8) ? 8
         : (mode == MV_GEMM_LINEAR && block_n >= 128 && !light) ? MV_LINEAR_EPI_WARPS : 4;
}
__host__ __device__ constexpr int gemm_threads(int mode, int block_n, bool light) {
  return 64 + 32 * gemm_epi_warps(mode, block_n, light);
}

struct GemmDev {
  int m, n, k;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  int act, out_f32;
  void* out;
  long long ldo;
  void* aux;
  long long ldaux;
  const float* scale;
  const float* shift;
  const float* resid;
  long long ldr;
  const void* in2;
  long long ldin2;
  int rows_per_group, group_stride, row_offset, resid_row_mod;
  // implicit-GEMM 3x3 convolution (pad 1) over NHWC sources: A tiles are TMA 4-D boxes, OOB zero fill = padding
  float* colstats;     // optional fp32 [2, N]: per-column sum and sum of squares of the stored values (atomic)
  int splits;          // split-K factor (MV_GEMM_NN_ATOMIC only, else 1)
  int conv;            // 0: A is a plain [M, K] matrix
  int conv_h, conv_w;  // OUTPUT spatial size
  int conv_tw;         // tile = conv_tw x (128 / conv_tw) output pixels (full rows when conv_w < 128)
  int conv_stride;     // 1 or 2
  int conv_rows_per_kb;  // wgrad: map rows covered by one 64-pixel k block when conv_w < 64
  int conv_cb0, conv_cb1;  // 64-channel blocks per tap taken from source 0 / source 1 (channel concat)
  // diagnostics (mv_gemm_set_profile_buffer): per CTA 8 x int64 cycle counters, nullptr in production
  //   [0] producer waiting for a free smem slot   [1] MMA warp waiting for operands   [2] MMA warp waiting for a free
  //   accumulator   [3] epilogue warp 2 waiting for an accumulator   [4] epilogue warp 2 busy   [5] CTA lifetime
  //   [6] tiles of this CTA
  long long* prof;
  int kskip0, kskip1;  // k blocks [kskip0, kskip1) are skipped (their B columns are zero)
  uint32_t idesc_clear;  // instruction-descriptor format bits to clear: bit 7 (A is fp16, not bf16), bit 10 (B is fp16)
  // Stream-K for the last, partial wave of tiles (sk_rem > 0): tiles [0, sk_dp) run whole, round-robin over the scheduling
  // units (CTAs / CTA pairs); the k blocks of the remaining sk_rem tiles are divided evenly over ALL units. Partial
  // accumulators of a tile meet in its fp32 workspace slot (vector red.add at L2); the contributor that arrives last
  // (sk_cnt) reads the sums back, zeroes the slot for the next launch and runs the ordinary epilogue. Nobody waits.
  int sk_dp, sk_rem;
  float* sk_ws;
  int* sk_cnt;
};

__host__ __device__ constexpr bool gemm_sk_mode(int mode, int block_n) {
  return (mode == MV_GEMM_LINEAR && block_n >= 32) || mode == MV_GEMM_SWIGLU || mode == MV_GEMM_SWIGLU_BWD;
}

// The sequence of (tile, k-block range) segments of one scheduling unit — identical in the producer, the MMA issuer and
// the epilogue warps. Whole tiles first (split-K "tiles" of MV_GEMM_NN_ATOMIC included), then this unit's share of the
// stream-K remainder, cut at tile boundaries.
struct SegIter {
  int units, num_mn, splits, nkb, dp_end, sk_dp;
  int tile_next;
  long long k_cur, k_hi;
  __device__ SegIter(const GemmDev& p, int unit, int units_) {
    units = units_;
    num_mn = p.num_m_blocks * p.num_n_blocks;
    splits = p.splits;
    nkb = p.num_k_blocks;
    sk_dp = p.sk_dp;
    dp_end = p.sk_dp;  // = all tiles unless a stream-K remainder or a separately launched tail follows
    tile_next = unit;
    const long long wk = (long long)p.sk_rem * nkb;
    k_cur = wk * unit / units;
    k_hi = wk * (unit + 1) / units;
  }
  __device__ bool next(int& tile, int& kb0, int& kb1, bool& partial) {
    if (tile_next < dp_end) {
      tile = tile_next;
      tile_next += units;
      const int split = tile / num_mn;
      kb0 = (int)((long long)nkb * split / splits);
      kb1 = (int)((long long)nkb * (split + 1) / splits);
      partial = false;
      return true;
    }
    if (k_cur < k_hi) {
      const int r = (int)(k_cur / nkb);
      kb0 = (int)(k_cur - (long long)r * nkb);
      const long long left = k_hi - k_cur;
      const int len = left < (long long)(nkb - kb0) ? (int)left : nkb - kb0;
      kb1 = kb0 + len;
      tile = sk_dp + r;
      partial = len != nkb;
      k_cur += len;
      return true;
    }
    return false;
  }
};

__device__ __forceinline__ void red_add_v4(float4* addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// PAIR: cta_group::2 — two CTAs of a cluster share one 256 x BLOCK_N tile; each stages its own 128 A rows and HALF of
// the B rows (the tensor core reads the other half from the peer's shared memory), halving B traffic per SM.
// LIGHT: shallow smem ring so that TWO CTAs fit on an SM (epilogue-bound problems: K of one or two blocks, or narrow N)
template <int BLOCK_N, int MODE = 0, bool PAIR = false, bool LIGHT = false>
struct GemmCfg {
  static constexpr int kRowsB = PAIR ? BLOCK_N / 2 : BLOCK_N;  // B rows staged by this CTA
  static constexpr int kBoxRowsB = kRowsB < 128 ? kRowsB : (kRowsB % 128 == 0 ? 128 : 96);
  static constexpr int kBoxesB = kRowsB / kBoxRowsB;
  static constexpr int kABytes = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr int kBBytes = kRowsB * GEMM_BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiWarps = gemm_epi_warps(MODE, BLOCK_N, LIGHT);
  static constexpr bool kEpiTma = MODE == MV_GEMM_LINEAR_TMA || gemm_tma_bf16(MODE);
  // LINEAR_TMA: per epilogue warp kTmaResStages residual tiles + kTmaOutStages output tiles of 32 rows x 32 fp32 columns
  // (4 KB each, 128-byte swizzle, 1 KB aligned) directly behind the operand ring; the transpose staging is not needed there
  static constexpr int kTmaResStages = MODE == MV_GEMM_LINEAR_TMA ? MV_TMA_RES_STAGES : 0;
  // (the eight-warp skinny-K form is bound by its epilogue: a second output tile per warp lets a store drain under the next chunk)
  static constexpr int kTmaOutStages = MODE == MV_GEMM_LINEAR_TMA ? MV_TMA_OUT_STAGES_F32
                                       : MODE == MV_GEMM_LINEAR_TMA_BF16_W8 ? 2 : MV_TMA_OUT_STAGES_BF16;
  static constexpr int kEpiTmaBytes = kEpiTma ? kEpiWarps * (kTmaResStages + kTmaOutStages) * 4096 : 0;
  static constexpr int kStagingBytes = kEpiTma ? 0 : kEpiWarps * 32 * 36 * 4;  // per-epilogue-warp 32x32 fp32 transpose tile (padded rows)
  static constexpr int kStatBytes = 2 * 256 * 4;         // CTA-level per-column (sum, sumsq) accumulators
  // LINEAR: per-epilogue-warp copy of the tile's (scale, shift) columns, staged before the accumulator is awaited
  static constexpr int kCoefBytes = ((MODE == MV_GEMM_LINEAR || kEpiTma) && BLOCK_N >= 32) ? kEpiWarps * 2 * BLOCK_N * 4 : 0;
  static constexpr int kBarBytes = 512;  // mbarriers: ring + accumulator stages in the first 256 B, TMA-epilogue barriers behind
  static constexpr int kFixedBytes = 1024 /*align*/ + kBarBytes + kEpiTmaBytes + kStagingBytes + kStatBytes + kCoefBytes;
  static constexpr int kRingBudget = 227 * 1024 - kFixedBytes;
  static constexpr int kStagesDeep = kRingBudget / kStageBytes > 8 ? 8 : kRingBudget / kStageBytes;
  static constexpr int kStages = !LIGHT ? kStagesDeep : (BLOCK_N <= 32 ? 4 : BLOCK_N <= 64 ? 3 : 2);
  // HEAD_CONV keeps the 9 taps in separate 16-column accumulators (144 columns per stage, stage stride 256)
  static constexpr int kAccStride = MODE == MV_GEMM_HEAD_CONV ? 256 : BLOCK_N;
  static constexpr int kTmemRaw = 2 * BLOCK_N;
  static constexpr int kTmemCols = MODE == MV_GEMM_HEAD_CONV ? 512
                                   : kTmemRaw <= 32 ? 32 : kTmemRaw <= 64 ? 64 : kTmemRaw <= 128 ? 128 : kTmemRaw <= 256 ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + kFixedBytes;
  static constexpr int kRingEnd = kStages * kStageBytes + kEpiTmaBytes;  // barriers start here
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
  static_assert(!kEpiTma || (kStageBytes % 1024 == 0 && BLOCK_N % 32 == 0), "TMA epilogue tiles must stay 1 KB aligned");
  static_assert(BLOCK_N % 16 == 0 && BLOCK_N >= 16 && BLOCK_N <= 256, "UMMA N");
  static_assert((kTmemCols & (kTmemCols - 1)) == 0 && kTmemCols <= 512, "TMEM columns must be a power of two <= 512");
};

// sigmoid through one MUFU op: sigma(x) = 0.5 * tanh(x / 2) + 0.5 (tanh.approx: ~2^-11 relative error, far below the
// bf16 rounding of everything these epilogues store); the exp + full-precision divide form costs ~12 instructions and two
// MUFU ops per element, which made the SwiGLU epilogues issue-bound
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_f(float x) { return fmaf(tanh_approx(0.5f * x), 0.5f, 0.5f); }
__device__ __forceinline__ float silu_f(float x) { return x * sigmoid_f(x); }

// ------------------------------------------------------------------ epilogue pieces (one thread = one output row)
__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* dst, const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(dst) = u;
}
__device__ __forceinline__ void load_bf16x8(const __nv_bfloat16* src, float* f) {
  uint4 u = *reinterpret_cast<const uint4*>(src);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

template <int CH>
__device__ __forceinline__ void epilogue_linear_chunk(const GemmDev& p, const uint32_t* v, int m, int n) {
  // v: CH fp32 accumulators for row m, columns n .. n+CH-1
  long long orow = m, rrow = m;
  if (p.rows_per_group > 0) {
    int g = m / p.rows_per_group, r = m - g * p.rows_per_group;
    orow = (long long)g * p.group_stride + r + p.row_offset;
    rrow = p.resid_row_mod ? r : orow;
  }
#pragma unroll
  for (int j8 = 0; j8 < CH / 8; ++j8) {
    const int nn = n + j8 * 8;
    if (nn >= p.n) break;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[j8 * 8 + j]);
    if (p.scale) {
      float4 s0 = __ldg(reinterpret_cast<const float4*>(p.scale + nn));
      float4 s1 = __ldg(reinterpret_cast<const float4*>(p.scale + nn + 4));
      f[0] *= s0.x; f[1] *= s0.y; f[2] *= s0.z; f[3] *= s0.w;
      f[4] *= s1.x; f[5] *= s1.y; f[6] *= s1.z; f[7] *= s1.w;
    }
    if (p.shift) {
      float4 s0 = __ldg(reinterpret_cast<const float4*>(p.shift + nn));
      float4 s1 = __ldg(reinterpret_cast<const float4*>(p.shift + nn + 4));
      f[0] += s0.x; f[1] += s0.y; f[2] += s0.z; f[3] += s0.w;
      f[4] += s1.x; f[5] += s1.y; f[6] += s1.z; f[7] += s1.w;
    }
    if (p.act == MV_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    if (p.resid) {
      const float* r = p.resid + rrow * p.ldr + nn;
      float4 r0 = *reinterpret_cast<const float4*>(r);
      float4 r1 = *reinterpret_cast<const float4*>(r + 4);
      f[0] += r0.x; f[1] += r0.y; f[2] += r0.z; f[3] += r0.w;
      f[4] += r1.x; f[5] += r1.y; f[6] += r1.z; f[7] += r1.w;
    }
    if (p.out_f32) {
      float* o = reinterpret_cast<float*>(p.out) + orow * p.ldo + nn;
      *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(f[4], f[5], f[6], f[7]);
      if (p.aux) store_bf16x8(reinterpret_cast<__nv_bfloat16*>(p.aux) + orow * p.ldaux + nn, f);
    } else {
      store_bf16x8(reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldo + nn, f);
    }
  }
}

// ------------------------------------------------------------------ kernel
template <int BLOCK_N, int MODE, bool PAIR = false, bool LIGHT = false>
__global__ void __launch_bounds__(gemm_threads(MODE, BLOCK_N, LIGHT), LIGHT ? 2 : 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a2,
                    const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_r,
                    const __grid_constant__ CUtensorMap tmap_o, const GemmDev p) {
  using Cfg = GemmCfg<BLOCK_N, MODE, PAIR, LIGHT>;
  constexpr int STAGES = Cfg::kStages;
  unsigned long long prof_ns0 = 0;
  if (kProf && p.prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(prof_ns0));
  static_assert(!LIGHT || (2 * Cfg::kTmemCols <= 512 && !PAIR), "two resident CTAs must share the 512 TMEM columns");
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int tile_start = PAIR ? (blockIdx.x >> 1) : blockIdx.x;
  const int tile_step = PAIR ? (gridDim.x >> 1) : gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Cfg::kRingEnd;
  // barrier layout (8 B each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then tmem ptr (4 B)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  // pair, non-leader CTA: its epilogue warps release an accumulator stage on this LOCAL barrier; one otherwise idle thread
  // forwards the release to the leader (the remote arrival needs cluster-scope release semantics, which compile to
  // MEMBAR.ALL.GPU + ERRBAR: 13 % of the stall samples of an epilogue-bound kernel when every epilogue warp paid for it)
  auto tfwd_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 4 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::kRingEnd + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  float* staging_all = reinterpret_cast<float*>(smem_gen + Cfg::kRingEnd + Cfg::kBarBytes);
  float* cstat = staging_all + Cfg::kStagingBytes / 4;  // [2][256]
  float* coef_all = cstat + 2 * 256;                      // [epilogue warp][scale BLOCK_N | shift BLOCK_N]

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_a2);
    tma_prefetch_desc(&tmap_b);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);  // pair: the leader expects the bytes of both CTAs, the peer's TMA only completes tx
      mbar_init(empty_bar(s), 1);
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      // one arrival per epilogue warp (pair: plus one for the peer CTA, forwarded by its idle MMA-warp thread)
      mbar_init(tempty_bar(s), Cfg::kEpiWarps + (PAIR ? 1 : 0));
      mbar_init(tfwd_bar(s), Cfg::kEpiWarps);
    }
    if constexpr (Cfg::kEpiTma) {
#pragma unroll
      for (int i = 0; i < Cfg::kEpiWarps * Cfg::kTmaResStages; ++i) mbar_init(bar_base + 256u + 8u * i, 1);
      tma_prefetch_desc(&tmap_r);
      tma_prefetch_desc(&tmap_o);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) { tmem_alloc_pair(tmem_slot, Cfg::kTmemCols); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  griddep_sync();  // PDL: the prologue above overlapped the previous kernel's tail; its results are visible from here on
  long long prof_acc[7] = {0, 0, 0, 0, 0, 0, 0};
  const long long prof_t0 = (kProf && p.prof) ? clock64() : 0;
  if (kProf && p.prof && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.prof[16ll * blockIdx.x + 8] = (long long)prof_ns0;
    p.prof[16ll * blockIdx.x + 9] = (long long)t;
  }

  const int num_mn = p.num_m_blocks * p.num_n_blocks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      SegIter it(p, tile_start, tile_step);
      int tile, kb0, kb1;
      bool partial;
      while (it.next(tile, kb0, kb1, partial)) {
        const int mn = tile % num_mn;
        const int m_blk = PAIR ? (mn % p.num_m_blocks) * 2 + (int)cta_rank : mn % p.num_m_blocks;  // 128-row block
        const int n_blk = mn / p.num_m_blocks;
        const int m0 = m_blk * GEMM_BLOCK_M;
        int brow[2];
        if (MODE == MV_GEMM_SWIGLU) {  // 128 gate rows + the matching 128 value rows
          brow[0] = n_blk * 128;
          brow[1] = p.n / 2 + n_blk * 128;
          if (PAIR) brow[0] = brow[cta_rank];  // leader stages the gate rows, the peer the value rows
        } else {
          brow[0] = n_blk * BLOCK_N + (PAIR ? (int)cta_rank * (BLOCK_N / 2) : 0);
          brow[1] = n_blk * BLOCK_N + Cfg::kBoxRowsB;
        }
        int cv_b = 0, cv_y = 0, cv_x = 0;
        if (p.conv && MODE != MV_GEMM_NN_ATOMIC) {  // tile rows are contiguous output pixels: m0 -> (image, y0, x0)
          const int hw = p.conv_h * p.conv_w;
          cv_b = m0 / hw;
          const int rem = m0 - cv_b * hw;
          cv_y = rem / p.conv_w;
          cv_x = rem - cv_y * p.conv_w;
        }
        const int cbt = p.conv_cb0 + p.conv_cb1;
        // the single producer thread is on the critical path of small-tile convolutions: no integer divisions inside
        // the k loop — tap / channel-block indices and the wgrad pixel coordinates advance incrementally
        int f_cbi = 0, f_kx = 0, f_ky = 0;  // forward / dgrad implicit GEMM: k block -> (tap, 64-channel block)
        if (p.conv && MODE != MV_GEMM_NN_ATOMIC && kb0 != 0) {
          const int tap = kb0 / cbt;
          f_cbi = kb0 - tap * cbt;
          f_ky = tap / 3;
          f_kx = tap - f_ky * 3;
        }
        // weight gradient: per tile the (tap, channel block) of each 64-column slice of B is fixed; per k block only the
        // 64-pixel window moves
        constexpr int NSL = BLOCK_N / 64 > 0 ? BLOCK_N / 64 : 1;
        int w_ok[NSL], w_c[NSL], w_dx[NSL], w_dy[NSL], w_src[NSL];
        int w_nvalid = 0, w_pb = 0, w_py = 0, w_px = 0;
        if (MODE == MV_GEMM_NN_ATOMIC && p.conv) {
#pragma unroll
          for (int ns = 0; ns < NSL; ++ns) {
            const int nb = n_blk * NSL + ns;
            w_ok[ns] = nb < 9 * cbt;
            const int tap = nb / cbt, cbi = nb - tap * cbt;
            const int ky = tap / 3, kx = tap - ky * 3;
            w_src[ns] = cbi < p.conv_cb0 ? 0 : 1;
            w_c[ns] = (cbi < p.conv_cb0 ? cbi : cbi - p.conv_cb0) * 64;
            w_dx[ns] = kx - 1;
            w_dy[ns] = ky - 1;
            w_nvalid += w_ok[ns] ? 1 : 0;
          }
          const int hw = p.conv_h * p.conv_w;
          const long long pm0 = (long long)kb0 * GEMM_BLOCK_K;
          w_pb = (int)(pm0 / hw);
          const int prem = (int)(pm0 - (long long)w_pb * hw);
          w_py = prem / p.conv_w;
          w_px = prem - w_py * p.conv_w;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          if (kb >= p.kskip0 && kb < p.kskip1) continue;
          const long long t0_ = (kProf && p.prof) ? clock64() : 0;
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (kProf && p.prof) prof_acc[0] += clock64() - t0_;
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + Cfg::kABytes;
          const bool nn_conv = MODE == MV_GEMM_NN_ATOMIC && p.conv;
          if (PAIR) {
            // both CTAs' TMA bytes land on the LEADER's full barrier; it expects the sum, the peer only arrives
            if (leader) mbar_expect_tx(full_bar(stage), 2 * Cfg::kStageBytes);
            tma_load_2d_pair(sa, &tmap_a, full_bar(stage), kb * GEMM_BLOCK_K, m0);
            tma_load_2d_pair(sb, &tmap_b, full_bar(stage), kb * GEMM_BLOCK_K, brow[0]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          if (!nn_conv) mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
          if (nn_conv) {
          } else if (p.conv) {
            const int ix = cv_x * p.conv_stride + f_kx - 1;
            const int iy = cv_y * p.conv_stride + f_ky - 1;
            if (f_cbi < p.conv_cb0) tma_load_4d(sa, &tmap_a, full_bar(stage), f_cbi * 64, ix, iy, cv_b);
            else tma_load_4d(sa, &tmap_a2, full_bar(stage), (f_cbi - p.conv_cb0) * 64, ix, iy, cv_b);
            if (++f_cbi == cbt) {
              f_cbi = 0;
              if (++f_kx == 3) { f_kx = 0; ++f_ky; }
            }
          } else
          tma_load_2d(sa, &tmap_a, full_bar(stage), kb * GEMM_BLOCK_K, m0);
#pragma unroll
          if constexpr (MODE == MV_GEMM_NN_ATOMIC) {
            // B is [K, N] row-major: [64 k x 64 n] boxes, one per 64 output columns (MN-major UMMA operand)
            if (p.conv) {
              // weight gradient of a 3x3 conv: k = output pixel, n = (tap, input channel); the B box is the input map
              // shifted by the tap (TMA 4-D tile of 64 pixels x 64 channels, zero fill = padding)
              mbar_expect_tx(full_bar(stage), Cfg::kABytes + w_nvalid * 8192);
              tma_load_2d(sa, &tmap_a, full_bar(stage), kb * GEMM_BLOCK_K, m0);
#pragma unroll
              for (int ns = 0; ns < NSL; ++ns) {
                if (w_ok[ns]) {
                  const int ix = w_px * p.conv_stride + w_dx[ns], iy = w_py * p.conv_stride + w_dy[ns];
                  tma_load_4d(sb + ns * 8192, w_src[ns] ? &tmap_a2 : &tmap_b, full_bar(stage), w_c[ns], ix, iy, w_pb);
                }
              }
              // next 64-pixel window (full rows when the map is narrower than 64 pixels)
              if (p.conv_w >= GEMM_BLOCK_K) {
                w_px += GEMM_BLOCK_K;
                if (w_px >= p.conv_w) { w_px = 0; ++w_py; }
              } else {
                w_py += p.conv_rows_per_kb;
              }
              if (w_py >= p.conv_h) { w_py = 0; ++w_pb; }
            } else {
#pragma unroll
            for (int ns = 0; ns < BLOCK_N / 64; ++ns)
              tma_load_2d(sb + ns * 8192, &tmap_b, full_bar(stage), n_blk * BLOCK_N + ns * 64, kb * GEMM_BLOCK_K);
            }
          } else {
#pragma unroll
          for (int bx = 0; bx < Cfg::kBoxesB; ++bx)
            tma_load_2d(sb + bx * (Cfg::kBoxRowsB * 128), &tmap_b, full_bar(stage), kb * GEMM_BLOCK_K, brow[bx]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && leader) {
      // a_format / b_format of kind::f16: 1 = bf16 (default), 0 = fp16 (training-mode decoder operands)
      const uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * GEMM_BLOCK_M : GEMM_BLOCK_M, BLOCK_N) & ~p.idesc_clear;
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      // descriptors of stage 0; the start-address field counts 16-byte units, so stage / k offsets are plain adds
      // (this single thread issues every MMA: with small tiles its instruction count per k block is the bottleneck)
      const uint64_t da0 = umma_desc_sw128(smem_base);
      const uint64_t db0 = MODE == MV_GEMM_NN_ATOMIC
                               ? umma_desc_sw128(smem_base + Cfg::kABytes, 1024, 8192)  // MN-major: 64-column atoms 8 KB apart
                               : umma_desc_sw128(smem_base + Cfg::kABytes);
      SegIter it(p, tile_start, tile_step);
      int tile, kb0, kb1;
      bool partial;
      while (it.next(tile, kb0, kb1, partial)) {
        long long t0_ = (kProf && p.prof) ? clock64() : 0;
        mbar_wait(tempty_bar(as), aphase ^ 1);
        if (kProf && p.prof) prof_acc[2] += clock64() - t0_;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * Cfg::kAccStride;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (kb >= p.kskip0 && kb < p.kskip1) continue;
          t0_ = (kProf && p.prof) ? clock64() : 0;
          mbar_wait(full_bar(stage), phase);
          if (kProf && p.prof) prof_acc[1] += clock64() - t0_;
          tc_fence_after();
          const uint64_t da = da0 + (uint64_t)(stage * (Cfg::kStageBytes >> 4));
          const uint64_t db = db0 + (uint64_t)(stage * (Cfg::kStageBytes >> 4));
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
            // advancing K by 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in the (addr >> 4) field
            if constexpr (MODE == MV_GEMM_HEAD_CONV) {
              umma_bf16(d_tmem + kb * 16, da + 2 * k, db + 2 * k, idesc, k != 0);
            } else if constexpr (MODE == MV_GEMM_NN_ATOMIC) {
              // B is MN-major: 16 reduction rows = 2 KB per k step; all BLOCK_N columns (64-column swizzle atoms,
              // leading-dimension offset 8 KB) in ONE instruction
              const uint32_t idesc_mn = umma_idesc_bf16(GEMM_BLOCK_M, BLOCK_N, 0, 1) & ~p.idesc_clear;
              umma_bf16(d_tmem, da + 2 * k, db + (2048 >> 4) * k, idesc_mn, (kb != kb0) || (k != 0));
            } else if constexpr (PAIR) {
              umma_bf16_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb != kb0) || (k != 0));
            } else {
              umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb != kb0) || (k != 0));
            }
          }
          // smem slot reusable once these MMAs have read it (pair: signalled in both CTAs)
          if (PAIR) umma_commit_pair(empty_bar(stage), 3); else umma_commit(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (PAIR) umma_commit_pair(tfull_bar(as), 3); else umma_commit(tfull_bar(as));  // accumulator complete
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
    if (PAIR && lane == 0 && !leader) {
      // ---- peer CTA: forward "accumulator stage released" from this CTA's epilogue warps to the leader's MMA thread
      SegIter it(p, tile_start, tile_step);
      int tile, kb0, kb1, as = 0;
      uint32_t fph = 0;
      bool partial;
      while (it.next(tile, kb0, kb1, partial)) {
        mbar_wait(tfwd_bar(as), fph);
        tc_fence_after();
        tc_fence_before();
        mbar_arrive_cluster(tempty_bar(as), 0);
        if (++as == 2) { as = 0; fph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: warps 2..5, TMEM lane quadrant = warp % 4 =====================
    const int quad = warp & 3;
    const int ew = warp - 2;           // epilogue warp index
    const int egrp = ew >> 2;          // which share of the column chunks (only with 8 epilogue warps)
    constexpr int EGRPS = Cfg::kEpiWarps / 4;
    const int row_in_tile = quad * 32 + lane;
    const int ep_tid = threadIdx.x - 64;  // 0..127 (4-warp modes)
    int as = 0;
    uint32_t aphase = 0;
    int stat_nblk = -1;
    constexpr int EPT = 32 * Cfg::kEpiWarps;  // epilogue threads
    auto ep_bar = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(EPT) : "memory"); };
    auto flush_stats = [&](int nb) {
      ep_bar();
      for (int i = ep_tid; i < 2 * BLOCK_N; i += EPT) {
        const float v = cstat[i];
        const int col = nb * BLOCK_N + (i % BLOCK_N);
        if (v != 0.f && col < p.n) atomicAdd(p.colstats + (i / BLOCK_N) * p.n + col, v);
        cstat[i] = 0.f;
      }
      ep_bar();
    };
    if (MODE == MV_GEMM_LINEAR && p.colstats) {
      for (int i = ep_tid; i < 2 * BLOCK_N; i += EPT) cstat[i] = 0.f;
      ep_bar();
    }
    __shared__ int sk_last;
    // LINEAR_TMA: this warp's residual / output tiles and residual barriers; running chunk counters give stage and parity
    constexpr int TR = Cfg::kTmaResStages, TO = Cfg::kTmaOutStages;
    const uint32_t tma_epi_base = smem_base + STAGES * Cfg::kStageBytes + (uint32_t)ew * (TR + TO) * 4096u;
    uint8_t* tma_epi_gen = smem_gen + STAGES * Cfg::kStageBytes + (size_t)ew * (TR + TO) * 4096;
    auto rbar = [&](int s_) { return bar_base + 256u + 8u * (uint32_t)(ew * TR + s_); };
    uint32_t tma_issued = 0, tma_waited = 0, tma_stores = 0;
    SegIter it(p, tile_start, tile_step);
    int tile, kb0_, kb1_;
    bool partial;
    while (it.next(tile, kb0_, kb1_, partial)) {
      const int mn = tile % num_mn;
      const int m_blk = PAIR ? (mn % p.num_m_blocks) * 2 + (int)cta_rank : mn % p.num_m_blocks;  // 128-row block
      const int n_blk = mn / p.num_m_blocks;
      if (MODE == MV_GEMM_LINEAR && p.colstats && n_blk != stat_nblk) {
        if (stat_nblk >= 0) flush_stats(stat_nblk);
        stat_nblk = n_blk;
      }
      const int m = m_blk * GEMM_BLOCK_M + row_in_tile;
      const bool row_ok = m < p.m;
      // row remap of the patch-embedding GEMM (patch rows -> token rows); identity for everything else
      auto map_rows = [&](int mm, long long& orow, long long& rrow) {
        orow = mm;
        rrow = mm;
        if (p.rows_per_group > 0) {
          const int g = mm / p.rows_per_group, rr = mm - g * p.rows_per_group;
          orow = (long long)g * p.group_stride + rr + p.row_offset;
          rrow = p.resid_row_mod ? rr : orow;
        }
      };
      // fp32-output path: the residual of a 32-column chunk (8 row slots per lane) is fetched one chunk ahead, the first
      // one BEFORE waiting for the accumulator (the load latency hides behind the MMAs of this tile)
      float4 qn[8];
      auto load_resid = [&](int c, float4* q) {
        const int nn = n_blk * BLOCK_N + c * 32 + (lane & 7) * 4;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int mm = m_blk * GEMM_BLOCK_M + quad * 32 + it * 4 + (lane >> 3);
          q[it] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.resid && mm < p.m && nn < p.n) {
            long long orow, rrow;
            map_rows(mm, orow, rrow);
            q[it] = *reinterpret_cast<const float4*>(p.resid + rrow * p.ldr + nn);
          }
        }
      };
      float* coef = coef_all + ew * (2 * BLOCK_N);
      if constexpr ((MODE == MV_GEMM_LINEAR || Cfg::kEpiTma) && BLOCK_N >= 32) {
        // this tile's (scale, shift) columns -> warp-private smem, also ahead of the accumulator wait
#pragma unroll
        for (int i = lane * 4; i < BLOCK_N; i += 128) {
          const int nn = n_blk * BLOCK_N + i;
          float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
          if (nn < p.n) {
            if (p.scale) sc = __ldg(reinterpret_cast<const float4*>(p.scale + nn));
            if (p.shift) sh = __ldg(reinterpret_cast<const float4*>(p.shift + nn));
          }
          *reinterpret_cast<float4*>(coef + i) = sc;
          *reinterpret_cast<float4*>(coef + BLOCK_N + i) = sh;
        }
        if constexpr (MODE == MV_GEMM_LINEAR) {
          if (p.out_f32) load_resid(egrp, qn);
        }
        __syncwarp();
      }
      // SWIGLU_BWD: the saved pre-activations [g | v] of a 32-column chunk (4 row groups per lane), fetched one chunk
      // ahead — the first one before waiting for the accumulator — so their DRAM latency hides behind the MMAs
      uint4 hgn[4], hvn[4];
      auto load_h = [&](int c, uint4* hg, uint4* hv) {
        const int jj = n_blk * BLOCK_N + c * 32 + (lane & 3) * 8;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int mm = m_blk * GEMM_BLOCK_M + quad * 32 + it * 8 + (lane >> 2);
          hg[it] = make_uint4(0, 0, 0, 0);
          hv[it] = hg[it];
          if (mm < p.m && jj < p.n) {
            const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(p.in2) + (long long)mm * p.ldin2;
            hg[it] = *reinterpret_cast<const uint4*>(h + jj);
            hv[it] = *reinterpret_cast<const uint4*>(h + p.n + jj);
          }
        }
      };
      if constexpr (MODE == MV_GEMM_SWIGLU_BWD) load_h(egrp, hgn, hvn);
      // skinny-K TMA epilogue with GATE_MASK: this row's gate gradients for every chunk this warp converts (4 heads = 8 bytes
      // per 64-column chunk), requested before the accumulator is awaited
      constexpr int kDuChunks = MODE == MV_GEMM_LINEAR_TMA_BF16_W8 ? (BLOCK_N / 64 + EGRPS - 1) / EGRPS : 1;
      uint2 du_row[kDuChunks];
      if constexpr (MODE == MV_GEMM_LINEAR_TMA_BF16_W8) {
#pragma unroll
        for (int i = 0; i < kDuChunks; ++i) {
          du_row[i] = make_uint2(0u, 0u);
          const int c0_ = n_blk * BLOCK_N + (egrp + i * EGRPS) * 64;
          if (p.act == MV_ACT_GATE_MASK && row_ok && c0_ < p.n)
            du_row[i] = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.in2) + (long long)m * p.ldin2 + (c0_ >> 4));
        }
      }
      // GATE_MASK (heads backward, 128-wide two-CTA-per-SM tiles): the per-(row, head) gate gradients of the WHOLE tile are
      // fetched before the accumulator wait — this epilogue-bound kernel otherwise exposes their latency once per chunk
      constexpr bool kPreloadDu = MODE == MV_GEMM_LINEAR && LIGHT && BLOCK_N == 128;
      float du_tile[kPreloadDu ? 4 : 1][4];
      if constexpr (kPreloadDu) {
        if (p.act == MV_ACT_GATE_MASK) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int nnm = n_blk * BLOCK_N + c * 32 + (lane & 3) * 8;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int mm = m_blk * GEMM_BLOCK_M + quad * 32 + it * 8 + (lane >> 2);
              du_tile[c][it] = 0.f;
              if (mm < p.m && nnm < p.n)
                du_tile[c][it] = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.in2)[(long long)mm * p.ldin2 + (nnm >> 4)]);
            }
          }
        }
      }
      // LINEAR_TMA: residual tiles of the first chunks are requested before the accumulator is awaited
      auto issue_res = [&](int chunk) {
        const uint32_t s_ = tma_issued % TR;
        ++tma_issued;
        mbar_expect_tx(rbar(s_), 4096u);
        tma_load_2d(tma_epi_base + s_ * 4096u, &tmap_r, rbar(s_), n_blk * BLOCK_N + chunk * 32, m_blk * GEMM_BLOCK_M + quad * 32);
      };
      if constexpr (MODE == MV_GEMM_LINEAR_TMA) {
        if (lane == 0) {
#pragma unroll
          for (int c = 0; c < TR && c < BLOCK_N / 32; ++c) issue_res(c);
        }
      }
      const long long te0_ = (kProf && p.prof) ? clock64() : 0;
      mbar_wait(tfull_bar(as), aphase);
      const long long te1_ = (kProf && p.prof) ? clock64() : 0;
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * Cfg::kAccStride;
      // ---- stream-K segment of a remainder tile: add the partial accumulators to the tile's workspace slot; only the
      // contributor that arrives last goes on to the epilogue, reading the sums back instead of TMEM
      bool from_ws = false;
      float4* ws4 = nullptr;
      if constexpr (gemm_sk_mode(MODE, BLOCK_N) && !LIGHT) {
        if (partial) {
          const int slot = (tile - p.sk_dp) * (PAIR ? 2 : 1) + (int)cta_rank;
          ws4 = reinterpret_cast<float4*>(p.sk_ws) + (size_t)slot * (BLOCK_N / 4) * 128;
#pragma unroll 1
          for (int c = egrp; c < BLOCK_N / 32; c += EGRPS) {
            uint32_t v[32];
            tmem_ld32(taddr + c * 32, v);
            tmem_ld_wait();
            float4* pp = ws4 + (size_t)(c * 8) * 128 + row_in_tile;  // [column / 4][row][4]: a warp's 32 rows are contiguous
#pragma unroll
            for (int q = 0; q < 8; ++q) red_add_v4(pp + q * 128, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {  // the accumulator stage is free again
            if (PAIR && !leader) mbar_arrive(tfwd_bar(as)); else mbar_arrive(tempty_bar(as));
          }
          __threadfence();
          ep_bar();
          if (ep_tid == 0) {
            // contributors of this tile = units whose k range meets [r nkb, (r + 1) nkb)  (ranges: [wk u / U, wk (u + 1) / U))
            const long long nkb = p.num_k_blocks, wk = (long long)p.sk_rem * nkb, U = tile_step;
            const long long s0 = (long long)(tile - p.sk_dp) * nkb, e0 = s0 + nkb;
            const int ua = (int)(((s0 + 1) * U + wk - 1) / wk) - 1, ub = (int)((e0 * U + wk - 1) / wk) - 1;
            const int n_contrib = ub - ua + 1;
            const int old = atomicAdd(p.sk_cnt + slot, 1);
            const int last = old == n_contrib - 1;
            if (last) p.sk_cnt[slot] = 0;  // ready for the next launch
            sk_last = last;
          }
          ep_bar();
          const bool last = sk_last != 0;
          ep_bar();  // sk_last may be rewritten by the next segment only after everyone has read it
          if (!last) {
            if (++as == 2) { as = 0; aphase ^= 1; }
            continue;
          }
          __threadfence();
          from_ws = true;
        }
      }
      // accumulator chunk source of the epilogues below: TMEM, or (stream-K finisher) the summed workspace slot, which is
      // zeroed as it is read
      auto acc_ld32 = [&](int col0, uint32_t (&v)[32]) {
        if (!(gemm_sk_mode(MODE, BLOCK_N) && !LIGHT) || !from_ws) {
          tmem_ld32(taddr + col0, v);
          return;
        }
        float4* pp = ws4 + (size_t)(col0 / 4) * 128 + row_in_tile;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 t = __ldcg(pp + q * 128);
          __stcg(pp + q * 128, make_float4(0.f, 0.f, 0.f, 0.f));
          v[4 * q] = __float_as_uint(t.x); v[4 * q + 1] = __float_as_uint(t.y);
          v[4 * q + 2] = __float_as_uint(t.z); v[4 * q + 3] = __float_as_uint(t.w);
        }
      };
      auto acc_wait = [&]() { if (!(gemm_sk_mode(MODE, BLOCK_N) && !LIGHT) || !from_ws) tmem_ld_wait(); };

      if constexpr (MODE == MV_GEMM_LINEAR && BLOCK_N >= 32) {
        // TMEM (one row per lane) -> padded smem tile -> row-contiguous global accesses (coalesced residual read,
        // output write); scale / shift are per column, so they are applied after the transpose (fixed per lane).
        // The TMEM load of chunk c+1 is issued as soon as chunk c sits in smem, so it overlaps the store phase.
        float* stg = staging_all + ew * (32 * 36);
        const int m_warp = m_blk * GEMM_BLOCK_M + quad * 32;
        constexpr int NC = BLOCK_N / 32;
        uint32_t v[32];
        acc_ld32(egrp * 32, v);
#pragma unroll 1
        for (int c = egrp; c < NC; c += EGRPS) {
          // GATE_MASK: the per-(row, head) gate gradients of this chunk, fetched before the TMEM wait (hidden latency)
          float du4[4] = {0.f, 0.f, 0.f, 0.f};
          if constexpr (kPreloadDu) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
              if (cc == c) {
#pragma unroll
                for (int it = 0; it < 4; ++it) du4[it] = du_tile[cc][it];
              }
          } else if (p.act == MV_ACT_GATE_MASK) {
            const int nnm = n_blk * BLOCK_N + c * 32 + (lane & 3) * 8;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int mm = m_warp + it * 8 + (lane >> 2);
              if (mm < p.m && nnm < p.n)
                du4[it] = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.in2)[(long long)mm * p.ldin2 + (nnm >> 4)]);
            }
          }
          acc_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(stg + lane * 36 + j * 4) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          if (c + EGRPS < NC) acc_ld32((c + EGRPS) * 32, v);
          float4 q_[8];
          if (p.out_f32) {  // residual of this chunk was prefetched; start fetching the next chunk's now
#pragma unroll
            for (int it = 0; it < 8; ++it) q_[it] = qn[it];
            if (c + EGRPS < NC) load_resid(c + EGRPS, qn);
          }
          __syncwarp();
          const int n0 = n_blk * BLOCK_N + c * 32;
          if (p.out_f32) {
            const int col = (lane & 7) * 4;
            const int nn = n0 + col;
            float4 cs4 = make_float4(0.f, 0.f, 0.f, 0.f), cq4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (nn < p.n) {
              const float4 sc = *reinterpret_cast<const float4*>(coef + c * 32 + col);
              const float4 sh = *reinterpret_cast<const float4*>(coef + BLOCK_N + c * 32 + col);
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                const int r = it * 4 + (lane >> 3);
                const int mm = m_warp + r;
                if (mm < p.m) {
                  float4 a = *reinterpret_cast<const float4*>(stg + r * 36 + col);
                  a.x = a.x * sc.x + sh.x; a.y = a.y * sc.y + sh.y; a.z = a.z * sc.z + sh.z; a.w = a.w * sc.w + sh.w;
                  if (p.act == MV_ACT_RELU) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
                  a.x += q_[it].x; a.y += q_[it].y; a.z += q_[it].z; a.w += q_[it].w;
                  cs4.x += a.x; cs4.y += a.y; cs4.z += a.z; cs4.w += a.w;
                  cq4.x += a.x * a.x; cq4.y += a.y * a.y; cq4.z += a.z * a.z; cq4.w += a.w * a.w;
                  long long orow, rrow;
                  map_rows(mm, orow, rrow);
                  *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + orow * p.ldo + nn) = a;
                  if (p.aux) {
                    uint2 u;
                    u.x = pack_bf16x2(a.x, a.y);
                    u.y = pack_bf16x2(a.z, a.w);
                    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.aux) + orow * p.ldaux + nn) = u;
                  }
                }
              }
            }
            if (p.colstats) {  // rows live in lane >> 3: fold the 4 row groups, one atomic per column per warp
#pragma unroll
              for (int o = 8; o < 32; o <<= 1) {
                cs4.x += __shfl_xor_sync(0xffffffffu, cs4.x, o); cs4.y += __shfl_xor_sync(0xffffffffu, cs4.y, o);
                cs4.z += __shfl_xor_sync(0xffffffffu, cs4.z, o); cs4.w += __shfl_xor_sync(0xffffffffu, cs4.w, o);
                cq4.x += __shfl_xor_sync(0xffffffffu, cq4.x, o); cq4.y += __shfl_xor_sync(0xffffffffu, cq4.y, o);
                cq4.z += __shfl_xor_sync(0xffffffffu, cq4.z, o); cq4.w += __shfl_xor_sync(0xffffffffu, cq4.w, o);
              }
              if (lane < 8 && nn < p.n) {
                float* cs = cstat + c * 32 + col;
                atomicAdd(cs + 0, cs4.x); atomicAdd(cs + 1, cs4.y); atomicAdd(cs + 2, cs4.z); atomicAdd(cs + 3, cs4.w);
                atomicAdd(cs + BLOCK_N + 0, cq4.x); atomicAdd(cs + BLOCK_N + 1, cq4.y);
                atomicAdd(cs + BLOCK_N + 2, cq4.z); atomicAdd(cs + BLOCK_N + 3, cq4.w);
              }
            }
          } else {
            const int col = (lane & 3) * 8;
            const int nn = n0 + col;
            const bool col_ok = nn < p.n;
            float sc[8], sh[8], csum[8], csq[8];
            {
              const float4 s0 = *reinterpret_cast<const float4*>(coef + c * 32 + col);
              const float4 s1 = *reinterpret_cast<const float4*>(coef + c * 32 + col + 4);
              const float4 h0 = *reinterpret_cast<const float4*>(coef + BLOCK_N + c * 32 + col);
              const float4 h1 = *reinterpret_cast<const float4*>(coef + BLOCK_N + c * 32 + col + 4);
              sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
              sh[0] = h0.x; sh[1] = h0.y; sh[2] = h0.z; sh[3] = h0.w; sh[4] = h1.x; sh[5] = h1.y; sh[6] = h1.z; sh[7] = h1.w;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) { csum[j] = 0.f; csq[j] = 0.f; }
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int r = it * 8 + (lane >> 2);
              const int mm = m_warp + r;
              if (mm < p.m && col_ok) {
                const float4 a0 = *reinterpret_cast<const float4*>(stg + r * 36 + col);
                const float4 a1 = *reinterpret_cast<const float4*>(stg + r * 36 + col + 4);
                float f[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  f[j] = f[j] * sc[j] + sh[j];
                  if (p.act == MV_ACT_RELU) f[j] = fmaxf(f[j], 0.f);
                }
                if (p.act == MV_ACT_GATE_MASK) {  // e = du[m, head] where the (batch-normalised) unit is active
                  const float du = du4[it];
#pragma unroll
                  for (int j = 0; j < 8; ++j) f[j] = f[j] > 0.f ? du : 0.f;
                }
                long long orow, rrow;
                map_rows(mm, orow, rrow);
                if (p.resid) {
                  const float* q = p.resid + rrow * p.ldr + nn;
                  const float4 q0 = *reinterpret_cast<const float4*>(q), q1 = *reinterpret_cast<const float4*>(q + 4);
                  f[0] += q0.x; f[1] += q0.y; f[2] += q0.z; f[3] += q0.w; f[4] += q1.x; f[5] += q1.y; f[6] += q1.z; f[7] += q1.w;
                }
                if (p.colstats) {  // BatchNorm batch statistics of exactly what is stored (bf16-rounded)
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float fr = p.out ? __bfloat162float(__float2bfloat16(f[j])) : f[j];
                    csum[j] += fr;
                    csq[j] += fr * fr;
                  }
                }
                if (p.out) store_bf16x8(reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldo + nn, f);
              }
            }
            if (p.colstats) {  // rows live in lane >> 2: fold the 8 row groups, then one atomic per column per warp
#pragma unroll
              for (int j = 0; j < 8; ++j) {
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                  csum[j] += __shfl_xor_sync(0xffffffffu, csum[j], o);
                  csq[j] += __shfl_xor_sync(0xffffffffu, csq[j], o);
                }
              }
              if (lane < 4 && col_ok) {
                float* cs = cstat + c * 32 + col;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  atomicAdd(cs + j, csum[j]);
                  atomicAdd(cs + BLOCK_N + j, csq[j]);
                }
              }
            }
          }
          __syncwarp();
        }
      } else if constexpr (MODE == MV_GEMM_LINEAR_TMA) {
        // x_out = acc * scale + shift + residual, fp32, 32 rows x 32 columns per step and warp.  The accumulator arrives one
        // row per lane (tcgen05.ld 32x32b); the residual tile sits in shared memory in the 128-byte-swizzled layout TMA
        // wrote (16-byte chunk j of row r at r * 128 + ((j ^ (r & 7)) << 4)), which a row-per-lane reader walks without bank
        // conflicts; the result goes back through the same layout and one TMA store. No per-thread global access at all:
        // the bytes in flight are set by the bulk copies, not by how many loads four warps can keep outstanding.
        constexpr int NC = BLOCK_N / 32;
        const int m_warp = m_blk * GEMM_BLOCK_M + quad * 32;
        const int swz = lane & 7;
        uint32_t v[32];
        acc_ld32(0, v);
#pragma unroll 1
        for (int c = 0; c < NC; ++c) {
          acc_wait();
          float a[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) a[j] = __uint_as_float(v[j]);
          if (c + 1 < NC) acc_ld32((c + 1) * 32, v);
          const uint32_t s_ = tma_waited % TR;
          mbar_wait(rbar(s_), (tma_waited / TR) & 1u);
          ++tma_waited;
          const uint8_t* rs = tma_epi_gen + s_ * 4096 + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 r4 = *reinterpret_cast<const float4*>(rs + ((j ^ swz) << 4));
            const float4 sc = *reinterpret_cast<const float4*>(coef + c * 32 + 4 * j);            // same address in every lane:
            const float4 sh = *reinterpret_cast<const float4*>(coef + BLOCK_N + c * 32 + 4 * j);  // shared-memory broadcast
            a[4 * j + 0] = fmaf(a[4 * j + 0], sc.x, sh.x) + r4.x;
            a[4 * j + 1] = fmaf(a[4 * j + 1], sc.y, sh.y) + r4.y;
            a[4 * j + 2] = fmaf(a[4 * j + 2], sc.z, sh.z) + r4.z;
            a[4 * j + 3] = fmaf(a[4 * j + 3], sc.w, sh.w) + r4.w;
          }
          fence_proxy_async_smem();  // our generic reads of the residual stage precede the async-proxy refill below
          __syncwarp();
          if (lane == 0) {
            if (c + TR < NC) issue_res(c + TR);
            tma_store_wait_read<TO - 1>();  // the store that last used this output stage has drained its shared-memory source
          }
          __syncwarp();
          uint8_t* os = tma_epi_gen + (TR + (c % TO)) * 4096 + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(os + ((j ^ swz) << 4)) = make_float4(a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
          fence_proxy_async_smem();  // generic writes -> visible to the TMA store
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmap_o, tma_epi_base + (TR + (c % TO)) * 4096u, n_blk * BLOCK_N + c * 32, m_warp);
            tma_store_commit();
          }
        }
      } else if constexpr (gemm_tma_bf16(MODE)) {
        // out = act(acc * scale + shift) in bf16: 32 rows x 64 columns (128-byte rows) per step and warp, written row per lane
        // into the 128-byte-swizzled tile a TMA store expects; output stages alternate so that a store drains while the next
        // chunk is converted.  With eight epilogue warps the two warps of a lane quadrant take every other chunk.
        constexpr int NC = BLOCK_N / 64;
        const int m_warp = m_blk * GEMM_BLOCK_M + quad * 32;
        const int swz = lane & 7;
        const bool relu = p.act == MV_ACT_RELU;
        const bool gmask = p.act == MV_ACT_GATE_MASK;
        const int n0 = n_blk * BLOCK_N;
        // chunks that start beyond N hold nothing (N = 144 on a 256-wide tile): never loaded, converted or stored
        auto live = [&](int c_) { return c_ < NC && n0 + c_ * 64 < p.n; };
        uint32_t v0[32], v1[32];
        if (live(egrp)) {
          acc_ld32(egrp * 64, v0);
          acc_ld32(egrp * 64 + 32, v1);
        }
#pragma unroll 1
        for (int c = egrp; live(c); c += EGRPS) {
          // GATE_MASK: out = du[row, column / 16] where the unit is active — the row's four gate gradients of this chunk
          uint2 dug = make_uint2(0u, 0u);
          if constexpr (MODE == MV_GEMM_LINEAR_TMA_BF16_W8) {
#pragma unroll
            for (int i = 0; i < kDuChunks; ++i)
              if (c == egrp + i * EGRPS) dug = du_row[i];
          }
          acc_wait();
          uint32_t pk[32];  // 64 bf16
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t* vv = h ? v1 : v0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 sc = *reinterpret_cast<const float4*>(coef + c * 64 + h * 32 + 4 * j);
              const float4 sh = *reinterpret_cast<const float4*>(coef + BLOCK_N + c * 64 + h * 32 + 4 * j);
              float f0 = fmaf(__uint_as_float(vv[4 * j + 0]), sc.x, sh.x), f1 = fmaf(__uint_as_float(vv[4 * j + 1]), sc.y, sh.y);
              float f2 = fmaf(__uint_as_float(vv[4 * j + 2]), sc.z, sh.z), f3 = fmaf(__uint_as_float(vv[4 * j + 3]), sc.w, sh.w);
              if (relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); f2 = fmaxf(f2, 0.f); f3 = fmaxf(f3, 0.f); }
              if (gmask) {  // columns 32 h + 4 j .. + 3 belong to head (2 h + j / 4) of the chunk's four
                const uint32_t w = h ? dug.y : dug.x;
                const float du = __uint_as_float(j < 4 ? (w << 16) : (w & 0xffff0000u));
                f0 = f0 > 0.f ? du : 0.f; f1 = f1 > 0.f ? du : 0.f; f2 = f2 > 0.f ? du : 0.f; f3 = f3 > 0.f ? du : 0.f;
              }
              pk[h * 16 + 2 * j] = pack_bf16x2(f0, f1);
              pk[h * 16 + 2 * j + 1] = pack_bf16x2(f2, f3);
            }
          }
          if (live(c + EGRPS)) {
            acc_ld32((c + EGRPS) * 64, v0);
            acc_ld32((c + EGRPS) * 64 + 32, v1);
          }
          const int oslot = (int)(tma_stores++ % (uint32_t)TO);  // running count: a warp may convert ONE chunk per tile (N = 144)
          if (lane == 0) tma_store_wait_read<TO - 1>();  // the store that last used this output stage has drained it
          __syncwarp();
          uint8_t* os = tma_epi_gen + oslot * 4096 + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(os + ((j ^ swz) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmap_o, tma_epi_base + oslot * 4096u, n0 + c * 64, m_warp);
            tma_store_commit();
          }
        }
      } else if constexpr (MODE == MV_GEMM_LINEAR) {
        constexpr int CH = 16;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / CH; ++c) {
          uint32_t v[CH];
          tmem_ld16(taddr + c * CH, v);
          tmem_ld_wait();
          if (row_ok) epilogue_linear_chunk<CH>(p, v, m, n_blk * BLOCK_N + c * CH);
        }
      } else if constexpr (MODE == MV_GEMM_SWIGLU) {
        // tile columns [0,128) = gate, [128,256) = value for hidden units n_blk*128 ..; silu(g)*v is formed per thread,
        // then transposed through smem so that every output row segment is written contiguously
        const int half = p.n / 2;
        float* stg = staging_all + ew * (32 * 36);
        const int m_warp = m_blk * GEMM_BLOCK_M + quad * 32;
        auto stage_and_store = [&](const float* f32vals, __nv_bfloat16* dst, long long ld, int col0) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(stg + lane * 36 + j * 4) =
                make_float4(f32vals[4 * j], f32vals[4 * j + 1], f32vals[4 * j + 2], f32vals[4 * j + 3]);
          __syncwarp();
          const int col = (lane & 3) * 8;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int r = it * 8 + (lane >> 2);
            const int mm = m_warp + r;
            if (mm < p.m) {
              const float4 a0 = *reinterpret_cast<const float4*>(stg + r * 36 + col);
              const float4 a1 = *reinterpret_cast<const float4*>(stg + r * 36 + col + 4);
              const float f[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
              store_bf16x8(dst + (long long)mm * ld + col0 + col, f);
            }
          }
          __syncwarp();
        };
#pragma unroll 1
        for (int c = egrp; c < 4; c += EGRPS) {
          uint32_t g[32], u[32];
          acc_ld32(c * 32, g);
          acc_ld32(128 + c * 32, u);
          acc_wait();
          const int j0 = n_blk * 128 + c * 32;
          float fg[32], fv[32], fo[32];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.shift + j0 + j4 * 4));
            const float4 c0 = __ldg(reinterpret_cast<const float4*>(p.shift + half + j0 + j4 * 4));
            fg[4 * j4 + 0] = __uint_as_float(g[4 * j4 + 0]) + b0.x; fv[4 * j4 + 0] = __uint_as_float(u[4 * j4 + 0]) + c0.x;
            fg[4 * j4 + 1] = __uint_as_float(g[4 * j4 + 1]) + b0.y; fv[4 * j4 + 1] = __uint_as_float(u[4 * j4 + 1]) + c0.y;
            fg[4 * j4 + 2] = __uint_as_float(g[4 * j4 + 2]) + b0.z; fv[4 * j4 + 2] = __uint_as_float(u[4 * j4 + 2]) + c0.z;
            fg[4 * j4 + 3] = __uint_as_float(g[4 * j4 + 3]) + b0.w; fv[4 * j4 + 3] = __uint_as_float(u[4 * j4 + 3]) + c0.w;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) fo[j] = silu_f(fg[j]) * fv[j];
          stage_and_store(fo, reinterpret_cast<__nv_bfloat16*>(p.out), p.ldo, j0);
          if (p.aux) {
            stage_and_store(fg, reinterpret_cast<__nv_bfloat16*>(p.aux), p.ldaux, j0);
            stage_and_store(fv, reinterpret_cast<__nv_bfloat16*>(p.aux), p.ldaux, half + j0);
          }
        }
      } else if constexpr (MODE == MV_GEMM_NN_ATOMIC) {
        // split-K partial sums: fp32 reductions into the (pre-zeroed) output
#pragma unroll 1
        for (int c = 0; c < BLOCK_N / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
          if (row_ok) {
            float* o = reinterpret_cast<float*>(p.out) + (long long)m * p.ldo + n_blk * BLOCK_N + c * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n_blk * BLOCK_N + c * 32 + j < p.n) atomicAdd(o + j, __uint_as_float(v[j]));
          }
        }
      } else if constexpr (MODE == MV_GEMM_HEAD_GATE) {
        // columns = heads x 16 hidden units of AttentionBlock.psi: g_h = sigmoid(w2_h . relu(scale*acc+shift) + b2_h)
        const int heads = p.n / 16;
        const float* w2 = reinterpret_cast<const float*>(p.in2);
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)m * p.ldo;
        // the tile's BLOCK_N / 16 gates of this row are collected in registers and stored as 16-byte vectors (they used to
        // leave as sixteen 2-byte stores per row, each lane on a different line)
        constexpr int HPT = BLOCK_N / 16;  // heads per tile
        float gate[HPT];
#pragma unroll
        for (int hl = 0; hl < HPT; ++hl) {
          const int hd = n_blk * HPT + hl;
          gate[hl] = 0.f;
          if (hd < heads) {  // warp-uniform
            uint32_t v[16];
            tmem_ld16(taddr + hl * 16, v);
            tmem_ld_wait();
            float acc = __ldg(p.resid + hd);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const int n = hd * 16 + j4 * 4;
              const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + n));
              const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + n));
              const float4 ww = __ldg(reinterpret_cast<const float4*>(w2 + n));
              acc += fmaxf(__uint_as_float(v[j4 * 4 + 0]) * sc.x + sh.x, 0.f) * ww.x;
              acc += fmaxf(__uint_as_float(v[j4 * 4 + 1]) * sc.y + sh.y, 0.f) * ww.y;
              acc += fmaxf(__uint_as_float(v[j4 * 4 + 2]) * sc.z + sh.z, 0.f) * ww.z;
              acc += fmaxf(__uint_as_float(v[j4 * 4 + 3]) * sc.w + sh.w, 0.f) * ww.w;
            }
            gate[hl] = sigmoid_f(acc);
          }
        }
        if (row_ok) {
          const int h0 = n_blk * HPT;
          if constexpr (HPT % 8 == 0) {
            if (p.ldo % 8 == 0 && h0 + HPT <= heads && (reinterpret_cast<uintptr_t>(o + h0) & 15) == 0) {
#pragma unroll
              for (int q = 0; q < HPT / 8; ++q) store_bf16x8(o + h0 + 8 * q, gate + 8 * q);
            } else {
#pragma unroll
              for (int hl = 0; hl < HPT; ++hl)
                if (h0 + hl < heads) o[h0 + hl] = __float2bfloat16(gate[hl]);
            }
          } else {
#pragma unroll
            for (int hl = 0; hl < HPT; ++hl)
              if (h0 + hl < heads) o[h0 + hl] = __float2bfloat16(gate[hl]);
          }
        }
      } else if constexpr (MODE == MV_GEMM_HEAD_CONV) {
        // columns [16*tap, 16*tap+16): t[h] = sum_c W3[h, c, tap] * f[c](pixel + tap); the 3x3 conv acts on f * g_h,
        // so each tap is weighted by the gate of the NEIGHBOUR pixel before the taps are summed, then bias + tanh.
        const int hw = p.conv_h * p.conv_w;
        const int bimg = m / hw;
        const int rem = m - bimg * hw;
        const int py = rem / p.conv_w, px = rem - py * p.conv_w;
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.f;
        // three taps per round: their TMEM loads and the neighbour-gate loads are all in flight before the first use
#pragma unroll 1
        for (int t3 = 0; t3 < 3; ++t3) {
          uint32_t v[3][16];
          uint4 g0[3], g1[3];
          bool inb[3];
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const int tap = t3 * 3 + q;
            tmem_ld16(taddr + tap * 16, v[q]);
            const int yy = py + t3 - 1, xx = px + q - 1;
            inb[q] = yy >= 0 && yy < p.conv_h && xx >= 0 && xx < p.conv_w;
            g0[q] = make_uint4(0, 0, 0, 0);
            g1[q] = g0[q];
            if (inb[q]) {
              const __nv_bfloat16* gp = reinterpret_cast<const __nv_bfloat16*>(p.in2) +
                                        ((long long)bimg * hw + (long long)yy * p.conv_w + xx) * p.ldin2;
              if (p.ldin2 == 16) {
                g0[q] = *reinterpret_cast<const uint4*>(gp);
                g1[q] = *reinterpret_cast<const uint4*>(gp + 8);
              } else {
                __nv_bfloat16 tmp[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) tmp[j] = j < p.n ? gp[j] : __float2bfloat16(0.f);
                g0[q] = *reinterpret_cast<const uint4*>(tmp);
                g1[q] = *reinterpret_cast<const uint4*>(tmp + 8);
              }
            }
          }
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            const uint32_t gu[8] = {g0[q].x, g0[q].y, g0[q].z, g0[q].w, g1[q].x, g1[q].y, g1[q].z, g1[q].w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float2 gf = unpack_bf16x2(gu[j]);
              acc[2 * j] += gf.x * __uint_as_float(v[q][2 * j]);
              acc[2 * j + 1] += gf.y * __uint_as_float(v[q][2 * j + 1]);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (j < p.n) {
            const float z = acc[j] + __ldg(p.shift + j);
            const float e = __expf(2.f * z);
            const float t = 1.f - __fdividef(2.f, e + 1.f);  // tanh (fast divide: 2 ulp, e + 1 >= 1)
            const long long oi = ((long long)bimg * p.n + j) * hw + rem;  // NCHW
            if (p.out_f32 == 1) reinterpret_cast<float*>(p.out)[oi] = t;
            else if (p.out_f32 == 0) reinterpret_cast<__nv_bfloat16*>(p.out)[oi] = __float2bfloat16(t);
            else {  // uint8 sink of SavePredictionsCallback (src/callbacks.py:345-346): truncating conversion
              const float q = fminf(fmaxf((t + 0.9f) / 1.8f, 0.f), 1.f) * 255.f;
              reinterpret_cast<uint8_t*>(p.out)[oi] = static_cast<uint8_t>(q);
            }
          }
        }
      } else {  // MV_GEMM_SWIGLU_BWD: acc = dU tile (128 hidden units); N == H
        const int H = p.n;
        float* stg = staging_all + ew * (32 * 36);
        const int m_warp = m_blk * GEMM_BLOCK_M + quad * 32;
#pragma unroll 1
        for (int c = egrp; c < BLOCK_N / 32; c += EGRPS) {
          uint32_t v[32];
          acc_ld32(c * 32, v);
          acc_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(stg + lane * 36 + j * 4) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          __syncwarp();
          const int col = (lane & 3) * 8;
          const int jj = n_blk * BLOCK_N + c * 32 + col;
          uint4 hg[4], hv[4];
#pragma unroll
          for (int it = 0; it < 4; ++it) { hg[it] = hgn[it]; hv[it] = hvn[it]; }
          if (c + EGRPS < BLOCK_N / 32) load_h(c + EGRPS, hgn, hvn);
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int r = it * 8 + (lane >> 2);
            const int mm = m_warp + r;
            if (mm < p.m && jj < H) {
              const float4 a0 = *reinterpret_cast<const float4*>(stg + r * 36 + col);
              const float4 a1 = *reinterpret_cast<const float4*>(stg + r * 36 + col + 4);
              const float du[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
              const uint32_t* pg = &hg[it].x;
              const uint32_t* pv = &hv[it].x;
              float dg[8], dv[8];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 fg = unpack_bf16x2(pg[j]), fv = unpack_bf16x2(pv[j]);
                const float s0 = sigmoid_f(fg.x), s1 = sigmoid_f(fg.y);
                dg[2 * j] = du[2 * j] * fv.x * (s0 * (1.f + fg.x * (1.f - s0)));
                dg[2 * j + 1] = du[2 * j + 1] * fv.y * (s1 * (1.f + fg.y * (1.f - s1)));
                dv[2 * j] = du[2 * j] * fg.x * s0;
                dv[2 * j + 1] = du[2 * j + 1] * fg.y * s1;
              }
              __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (long long)mm * p.ldo;
              store_bf16x8(o + jj, dg);
              store_bf16x8(o + H + jj, dv);
            }
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0 && !from_ws) {  // (a stream-K finisher released its accumulator stage before the counter)
        if (PAIR && !leader) mbar_arrive(tfwd_bar(as)); else mbar_arrive(tempty_bar(as));
      }
      if (kProf && p.prof && warp == 2) {
        prof_acc[3] += te1_ - te0_;
        prof_acc[4] += clock64() - te1_;
        prof_acc[6] += 1;
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (MODE == MV_GEMM_LINEAR && p.colstats && stat_nblk >= 0) flush_stats(stat_nblk);
    if constexpr (Cfg::kEpiTma) {
      if (lane == 0) tma_store_wait<0>();  // every output tile has landed before the CTA retires
    }
  }

  if (kProf && p.prof && lane == 0 && warp <= 2) {
    long long* q = p.prof + 16ll * blockIdx.x;
    if (warp == 0) { q[0] = prof_acc[0]; q[5] = clock64() - prof_t0; }
    if (warp == 1) { q[1] = prof_acc[1]; q[2] = prof_acc[2]; }
    if (warp == 2) { q[3] = prof_acc[3]; q[4] = prof_acc[4]; q[6] = prof_acc[6]; }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (kProf && p.prof && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.prof[16ll * blockIdx.x + 10] = (long long)t;
  }
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols); else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

static long long* g_gemm_prof = nullptr;  // diagnostics only (mv_gemm_set_profile_buffer)

// ------------------------------------------------------------------ host launch
// tile_limit > 0: only tiles [0, tile_limit) in m-fastest order are computed (the rest is launched separately, see
// launch_with_tail)
template <int BLOCK_N, int MODE, bool PAIR = false, bool LIGHT = false>
static int launch_gemm(const mv_gemm_args& a, cudaStream_t stream, int tile_limit = 0) {
  using Cfg = GemmCfg<BLOCK_N, MODE, PAIR, LIGHT>;
  static std::atomic<uint64_t> attr_set{0};  // one bit per device: function attributes are per device
  auto kern = gemm_bf16_tc_kernel<BLOCK_N, MODE, PAIR, LIGHT>;
  if (first_use_on_device(attr_set)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm<%d,%d>): %s", BLOCK_N, MODE, cudaGetErrorString(e));
      return (int)e;
    }
  }
  const CUtensorMap* ta = nullptr;
  const CUtensorMap* ta2 = nullptr;
  const CUtensorMap* tb_conv = nullptr;
  if (a.conv && MODE == MV_GEMM_NN_ATOMIC) {
    const int tw = a.conv_w < 64 ? a.conv_w : 64;
    const int th = 64 / tw;
    ta = get_tmap_2d_bf16(a.a, a.m, a.k, a.lda, GEMM_BLOCK_M);
    tb_conv = get_tmap_nhwc_bf16(a.b, a.conv_batch, a.conv_h * a.conv_stride, a.conv_w * a.conv_stride, a.conv_c0, tw, th,
                                 a.conv_stride);
    ta2 = a.conv_c1 > 0 ? get_tmap_nhwc_bf16(a.a2, a.conv_batch, a.conv_h * a.conv_stride, a.conv_w * a.conv_stride,
                                             a.conv_c1, tw, th, a.conv_stride)
                        : tb_conv;
    if (!tb_conv) return MV_ERR_ARG;
  } else if (a.conv) {
    const int tw = a.conv_w < 128 ? a.conv_w : 128;
    const int th = 128 / tw;
    ta = get_tmap_nhwc_bf16(a.a, a.conv_batch, a.conv_h * a.conv_stride, a.conv_w * a.conv_stride, a.conv_c0, tw, th,
                            a.conv_stride);
    ta2 = a.conv_c1 > 0 ? get_tmap_nhwc_bf16(a.a2, a.conv_batch, a.conv_h * a.conv_stride, a.conv_w * a.conv_stride,
                                             a.conv_c1, tw, th, a.conv_stride)
                        : ta;
  } else {
    ta = get_tmap_2d_bf16(a.a, a.m, a.k, a.lda, GEMM_BLOCK_M);
    ta2 = ta;
  }
  const CUtensorMap* tr = ta;  // residual / output maps of the TMA epilogue (dummies elsewhere)
  const CUtensorMap* to = ta;
  if (MODE == MV_GEMM_LINEAR_TMA) {
    tr = get_tmap_2d_f32(a.resid, a.m, a.n, a.ldr, 32);
    to = get_tmap_2d_f32(a.out, a.m, a.n, a.ldo, 32);
    if (!tr || !to) return MV_ERR_ARG;
  } else if (gemm_tma_bf16(MODE)) {
    to = get_tmap_2d_bf16(a.out, a.m, a.n, a.ldo, 32, 64);
    if (!to) return MV_ERR_ARG;
  }
  const CUtensorMap* tb = tb_conv ? tb_conv
                          : MODE == MV_GEMM_NN_ATOMIC ? get_tmap_2d_bf16(a.b, a.k, a.n, a.ldb, 64)
                                                      : get_tmap_2d_bf16(a.b, a.n, a.k, a.ldb, Cfg::kBoxRowsB);
  if (!ta || !ta2 || !tb) return MV_ERR_ARG;

  GemmDev p;
  p.m = a.m; p.n = a.n; p.k = a.k;
  p.num_m_blocks = PAIR ? (a.m + 2 * GEMM_BLOCK_M - 1) / (2 * GEMM_BLOCK_M) : (a.m + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  p.num_n_blocks = MODE == MV_GEMM_SWIGLU ? (a.n / 2) / 128 : (a.n + BLOCK_N - 1) / BLOCK_N;
  p.num_k_blocks = (a.k + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
  p.act = a.act; p.out_f32 = a.out_f32;
  p.out = a.out; p.ldo = a.ldo; p.aux = a.aux; p.ldaux = a.ldaux;
  p.scale = a.scale; p.shift = a.shift; p.resid = a.resid; p.ldr = a.ldr;
  p.in2 = a.in2; p.ldin2 = a.ldin2;
  p.rows_per_group = a.rows_per_group; p.group_stride = a.group_stride; p.row_offset = a.row_offset;
  p.resid_row_mod = a.resid_row_mod;
  p.splits = 1;
  p.colstats = a.colstats;
  p.prof = g_gemm_prof;
  p.kskip0 = a.kskip_begin / GEMM_BLOCK_K;
  p.kskip1 = a.kskip_end / GEMM_BLOCK_K;
  p.idesc_clear = ((a.ab_f16 & 1) ? (1u << 7) : 0u) | ((a.ab_f16 & 2) ? (1u << 10) : 0u);
  if (MODE == MV_GEMM_NN_ATOMIC) {
    const int sms = device_sms() > 0 ? device_sms() : 148;
    const int mn = p.num_m_blocks * p.num_n_blocks;
    int sp = a.reserved_splits > 0 ? a.reserved_splits : (sms + mn - 1) / mn;
    if (sp > p.num_k_blocks) sp = p.num_k_blocks;
    if (sp < 1) sp = 1;
    p.splits = sp;
  }
  p.conv = a.conv;
  p.conv_h = a.conv_h; p.conv_w = a.conv_w;
  p.conv_tw = a.conv_w < 128 ? a.conv_w : 128;
  p.conv_stride = a.conv_stride;
  p.conv_rows_per_kb = (a.conv && a.conv_w > 0 && a.conv_w < GEMM_BLOCK_K) ? GEMM_BLOCK_K / a.conv_w : 1;
  p.conv_cb0 = (a.conv_c0 + 63) / 64;
  p.conv_cb1 = (a.conv_c1 + 63) / 64;

  int tiles = p.num_m_blocks * p.num_n_blocks * p.splits;
  if (tile_limit > 0 && tile_limit < tiles) tiles = tile_limit;
  const int sms = device_sms() > 0 ? device_sms() : 148;
  int grid = sms;
  if (PAIR) grid &= ~1;
  if (LIGHT) grid *= 2;  // two co-resident CTAs per SM
  const int units_full = PAIR ? grid / 2 : grid;
  // ---- stream-K for the partial last wave (see GemmDev): needs a caller-provided zeroed workspace
  p.sk_dp = tiles;
  p.sk_rem = 0;
  p.sk_ws = nullptr;
  p.sk_cnt = nullptr;
  static const int sk_env = [] { const char* e = getenv("MV_GEMM_SK"); return e ? atoi(e) : 1; }();  // 0 off, 1 auto, 2 always
  if (gemm_sk_mode(MODE, BLOCK_N) && !LIGHT && sk_env != 0 && a.reserved3 != 2 && a.workspace && tile_limit <= 0 && !a.conv && !a.colstats && a.rows_per_group == 0 &&
      a.kskip_end == 0 && p.splits == 1 && (reinterpret_cast<uintptr_t>(a.workspace) & 255) == 0) {
    const int rem = tiles % units_full;
    const long long wk = (long long)rem * p.num_k_blocks;
    const long long need = 4096 + (long long)rem * (PAIR ? 2 : 1) * 128 * BLOCK_N * 4;
    // worth it when the last wave leaves a good part of the machine idle and every unit still gets a few k blocks
    const bool worth = sk_env == 2 || a.reserved3 == 1 || (rem * 100 <= units_full * 85 && wk >= 2ll * units_full);
    if (rem > 0 && wk >= units_full && worth && need <= a.workspace_bytes && rem * (PAIR ? 2 : 1) <= 1024) {
      p.sk_dp = tiles - rem;
      p.sk_rem = rem;
      p.sk_cnt = reinterpret_cast<int*>(a.workspace);
      p.sk_ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a.workspace) + 4096);
    }
  }
  if (PAIR) {
    if (p.sk_rem == 0 && 2 * tiles < grid) grid = 2 * tiles;
    (void)launch_pdl(kern, dim3(grid), dim3(gemm_threads(MODE, BLOCK_N, LIGHT)), (size_t)Cfg::kSmemBytes, stream, 2, *ta, *ta2, *tb, *tr, *to, p);
    MV_CHECK_LAUNCH("gemm_bf16_tc_pair");
    return MV_OK;
  }
  if (p.sk_rem == 0 && tiles < grid) grid = tiles;
  MV_LAUNCH(kern, grid, gemm_threads(MODE, BLOCK_N, LIGHT), Cfg::kSmemBytes, stream, *ta, *ta2, *tb, *tr, *to, p);
  MV_CHECK_LAUNCH("gemm_bf16_tc");
  return MV_OK;
}

// Tail re-tiling for the CTA-pair LINEAR GEMMs: T tiles of 256 x 256 on U = SMs / 2 pairs leave a last wave of R = T mod U
// tiles in which U - R pairs idle for a whole tile time. When those R tiles form one rectangle (they do whenever R <= the
// number of 256-row blocks: tiles are ordered m-fastest, so they are the bottom rows of the last 256-column block), the main
// launch stops after the full waves and the rectangle is computed by a second launch of 128 x 128 single-CTA tiles, which
// spreads it over up to 148 SMs at a quarter of the work each — no reduction, no extra traffic (stream-K, which needs one, was
// measured slower: DESIGN.md 3.1). NEGATIVE RESULT as well, kept behind a switch (see below). Returns the number of leading tiles the main launch keeps (0 = no split) and fills `tail`.
static int plan_tail(const mv_gemm_args& a, mv_gemm_args* tail) {
  // measured SLOWER than leaving the last wave partly idle (QKV 60.6 -> 71.8 us at M = 5264: the tail launch pays its own
  // prologue, runs the register epilogue on few SMs and cannot overlap the main launch, whose CTAs all end together), so it is
  // off unless asked for: MV_GEMM_TAIL=1 or reserved3 == 3
  static const int tail_env = [] { const char* e = getenv("MV_GEMM_TAIL"); return e ? atoi(e) : 0; }();
  if ((tail_env == 0 && a.reserved3 != 3) || a.aux || a.colstats || a.rows_per_group || a.kskip_end || a.conv || a.act == MV_ACT_GATE_MASK) return 0;
  const int sms = device_sms() > 0 ? device_sms() : 148;
  const int units = (sms & ~1) / 2;
  const int mp = (a.m + 255) / 256, nb = (a.n + 255) / 256;
  const int tiles = mp * nb, full = tiles / units, rem = tiles % units;
  if (full < 1 || rem == 0 || rem > mp) return 0;
  const int row0 = (mp - rem) * 256, col0 = (nb - 1) * 256;
  const int rows = a.m - row0, cols = a.n - col0;
  const int tail_tiles = ((rows + 127) / 128) * ((cols + 127) / 128);
  // cost of the tail in pair-tile times: quarter-size tiles, ~1.3x less efficient, in waves of `sms` tiles
  const double tail_cost = 0.5 * 1.3 * ((tail_tiles + sms - 1) / sms);
  if (tail_cost > 0.85) return 0;
  *tail = a;
  const size_t osz = a.out_f32 == 1 ? 4 : 2;
  tail->a = reinterpret_cast<const uint8_t*>(a.a) + (size_t)row0 * a.lda * 2;
  tail->b = reinterpret_cast<const uint8_t*>(a.b) + (size_t)col0 * a.ldb * 2;
  tail->m = rows;
  tail->n = cols;
  tail->out = reinterpret_cast<uint8_t*>(a.out) + ((size_t)row0 * a.ldo + col0) * osz;
  if (a.resid) tail->resid = a.resid + (size_t)row0 * a.ldr + col0;
  if (a.scale) tail->scale = a.scale + col0;
  if (a.shift) tail->shift = a.shift + col0;
  tail->block_n = 128;
  tail->reserved2 = 1;  // single CTAs
  tail->workspace = nullptr;
  return tiles - rem;
}

}  // namespace mv

extern "C" void mv_gemm_set_profile_buffer(void* buf) { mv::g_gemm_prof = reinterpret_cast<long long*>(buf); }

extern "C" int mv_gemm_bf16(const mv_gemm_args* args, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(args != nullptr, "mv_gemm_bf16: null args");
  const mv_gemm_args& a = *args;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MV_CHECK_ARG(a.a && a.b && (a.out || a.colstats), "mv_gemm_bf16: null operand");
  MV_CHECK_ARG(!a.colstats || (a.mode == MV_GEMM_LINEAR && a.n >= 32), "mv_gemm_bf16: colstats needs LINEAR mode, N >= 32");
  MV_CHECK_ARG(a.m > 0 && a.n > 0 && a.k > 0, "mv_gemm_bf16: empty problem m=%d n=%d k=%d", a.m, a.n, a.k);
  // measured on B200: a kind::f16 instruction descriptor whose A and B formats differ is an ILLEGAL INSTRUCTION
  MV_CHECK_ARG(a.ab_f16 == 0 || a.ab_f16 == 3, "mv_gemm_bf16: A and B must share one 16-bit format (both bf16 or both fp16)");
  MV_CHECK_ARG(a.n % 8 == 0 || a.mode == MV_GEMM_HEAD_CONV, "mv_gemm_bf16: N=%d must be a multiple of 8", a.n);
  MV_CHECK_ARG(a.k % 8 == 0 || a.mode == MV_GEMM_NN_ATOMIC, "mv_gemm_bf16: K=%d must be a multiple of 8", a.k);
  MV_CHECK_ARG((a.conv || a.lda % 8 == 0) && a.ldb % 8 == 0, "mv_gemm_bf16: lda/ldb must be multiples of 8 elements");
  if (a.conv && a.mode == MV_GEMM_NN_ATOMIC) {
    MV_CHECK_ARG(a.conv_stride == 1 || a.conv_stride == 2, "mv_gemm_bf16: conv stride must be 1 or 2");
    MV_CHECK_ARG(a.conv_c0 > 0 && a.conv_c0 % 8 == 0 && a.conv_c1 % 8 == 0 && (a.conv_c1 == 0 || a.a2),
                 "mv_gemm_bf16: conv channel counts must be multiples of 8");
    MV_CHECK_ARG(a.conv_w > 0 && a.conv_h > 0 && (a.conv_w % 64 == 0 || 64 % a.conv_w == 0) && (a.conv_h * a.conv_w) % 64 == 0,
                 "mv_gemm_bf16: wgrad output map %dx%d must tile into 64-pixel row blocks", a.conv_h, a.conv_w);
    MV_CHECK_ARG(a.k == a.conv_batch * a.conv_h * a.conv_w, "mv_gemm_bf16: wgrad K must be batch*H*W");
    MV_CHECK_ARG(a.n == 9 * 64 * ((a.conv_c0 + 63) / 64 + (a.conv_c1 + 63) / 64), "mv_gemm_bf16: wgrad N must be 9 x padded channels");
  } else if (a.conv) {
    MV_CHECK_ARG(a.mode == MV_GEMM_LINEAR || a.mode == MV_GEMM_HEAD_CONV,
                 "mv_gemm_bf16: conv A operand only with MV_GEMM_LINEAR / MV_GEMM_HEAD_CONV");
    MV_CHECK_ARG(a.conv_stride == 1 || a.conv_stride == 2, "mv_gemm_bf16: conv stride must be 1 or 2");
    MV_CHECK_ARG(a.conv_c0 > 0 && a.conv_c0 % 8 == 0 && a.conv_c1 % 8 == 0 && (a.conv_c1 == 0 || a.a2),
                 "mv_gemm_bf16: conv channel counts must be multiples of 8");
    MV_CHECK_ARG(a.conv_w > 0 && a.conv_h > 0 && (a.conv_w % 128 == 0 || 128 % a.conv_w == 0) &&
                     (a.conv_h * a.conv_w) % 128 == 0,
                 "mv_gemm_bf16: conv output %dx%d must tile into 128-pixel row blocks", a.conv_h, a.conv_w);
    MV_CHECK_ARG(a.m == a.conv_batch * a.conv_h * a.conv_w, "mv_gemm_bf16: conv M must be batch*H*W");
    MV_CHECK_ARG(a.k == 9 * 64 * ((a.conv_c0 + 63) / 64 + (a.conv_c1 + 63) / 64),
                 "mv_gemm_bf16: conv K must be 9 taps x 64-padded channel blocks (got %d)", a.k);
  }
  MV_CHECK_ARG(a.mode == MV_GEMM_HEAD_CONV || a.mode == MV_GEMM_HEAD_GATE || a.ldo % (a.out_f32 ? 4 : 8) == 0,
               "mv_gemm_bf16: ldo alignment");
  MV_CHECK_ARG((reinterpret_cast<uintptr_t>(a.out) & 15) == 0, "mv_gemm_bf16: out must be 16-byte aligned");
  MV_CHECK_ARG(!a.resid || ((reinterpret_cast<uintptr_t>(a.resid) & 15) == 0 && a.ldr % 4 == 0), "mv_gemm_bf16: resid alignment");
  MV_CHECK_ARG(!a.aux || ((reinterpret_cast<uintptr_t>(a.aux) & 15) == 0 && a.ldaux % 8 == 0), "mv_gemm_bf16: aux alignment");
  MV_CHECK_ARG(!a.scale || (reinterpret_cast<uintptr_t>(a.scale) & 15) == 0, "mv_gemm_bf16: scale alignment");
  MV_CHECK_ARG(!a.shift || (reinterpret_cast<uintptr_t>(a.shift) & 15) == 0, "mv_gemm_bf16: shift alignment");
  MV_CHECK_ARG(a.kskip_end == 0 || (a.mode == MV_GEMM_LINEAR && !a.conv && a.kskip_begin > 0 && a.kskip_begin % 64 == 0 &&
                                    a.kskip_end % 64 == 0 && a.kskip_end > a.kskip_begin && a.kskip_end <= a.k),
               "mv_gemm_bf16: kskip must be a non-empty 64-aligned K range after the first k block, LINEAR mode only");

  // CTA pairs (cta_group::2) for the big plain GEMMs; reserved2: 0 = auto, 1 = never, 2 = always (when legal)
  static const int pair_env = [] { const char* e = getenv("MV_GEMM_PAIR"); return e ? atoi(e) : -1; }();  // 0 disables
  const bool pair_legal = !a.conv && !a.colstats && a.m >= 256 && pair_env != 0;
  const bool use_pair = pair_legal && (a.reserved2 == 2 || (a.reserved2 == 0 && a.m >= 1024 && a.n >= 512));
  switch (a.mode) {
    case MV_GEMM_SWIGLU:
      MV_CHECK_ARG(a.n % 256 == 0 && a.shift && !a.out_f32, "mv_gemm_bf16(SWIGLU): N %% 256 == 0, bias required, bf16 out");
      // with 8 epilogue warps the CTA pair wins (M=5264: 112 -> 100 us, 1320 TFLOP/s); reserved2 == 1 forces single CTAs
      if (pair_legal && (a.reserved2 == 2 || (a.reserved2 == 0 && a.m >= 1024))) return launch_gemm<256, MV_GEMM_SWIGLU, true>(a, stream);
      return launch_gemm<256, MV_GEMM_SWIGLU>(a, stream);
    case MV_GEMM_SWIGLU_BWD:
      MV_CHECK_ARG(a.in2 && a.ldin2 % 8 == 0 && !a.out_f32, "mv_gemm_bf16(SWIGLU_BWD): in2 required, bf16 out");
      // 256 hidden units per tile on a CTA pair by default (M=10528: 131 us vs 168 us for 128-wide single-CTA tiles)
      if (a.n % 256 == 0 && (a.block_n == 256 || (a.block_n == 0 && a.m >= 1024))) {
        if (pair_legal && (a.reserved2 == 2 || (a.reserved2 == 0 && a.m >= 1024)))
          return launch_gemm<256, MV_GEMM_SWIGLU_BWD, true>(a, stream);
        return launch_gemm<256, MV_GEMM_SWIGLU_BWD>(a, stream);
      }
      if (pair_legal && a.reserved2 == 2) return launch_gemm<128, MV_GEMM_SWIGLU_BWD, true>(a, stream);
      return launch_gemm<128, MV_GEMM_SWIGLU_BWD>(a, stream);
    case MV_GEMM_NN_ATOMIC:
      MV_CHECK_ARG(a.out_f32 == 1 && a.n % 8 == 0, "mv_gemm_bf16(NN_ATOMIC): fp32 output, N %% 8 == 0");
      return launch_gemm<128, MV_GEMM_NN_ATOMIC>(a, stream);
    case MV_GEMM_HEAD_CONV:
      MV_CHECK_ARG(a.conv && a.conv_c1 == 0 && a.conv_c0 <= 64 && a.conv_stride == 1 && a.n >= 1 && a.n <= 16 && a.shift &&
                       a.in2 && a.ldin2 >= a.n,
                   "mv_gemm_bf16(HEAD_CONV): one <=64-channel NHWC source, <=16 heads, bias (shift) and gates (in2) required");
      return launch_gemm<16, MV_GEMM_HEAD_CONV>(a, stream);
    case MV_GEMM_HEAD_GATE:
      MV_CHECK_ARG(a.n % 16 == 0 && a.n <= 256 && a.scale && a.shift && a.in2 && a.resid && !a.out_f32,
                   "mv_gemm_bf16(HEAD_GATE): N = 16*heads <= 256; scale, shift, in2 (w2) and resid (b2) required");
      MV_CHECK_ARG((reinterpret_cast<uintptr_t>(a.in2) & 15) == 0, "mv_gemm_bf16(HEAD_GATE): w2 alignment");
      return launch_gemm<128, MV_GEMM_HEAD_GATE, false, true>(a, stream);
    case MV_GEMM_LINEAR: {
      int bn = a.block_n;
      if (bn == 0) {
        if (a.n <= 16) bn = 16;
        else if (a.n <= 32) bn = 32;
        else if (a.n <= 64) bn = 64;
        else if (a.n <= 128) bn = 128;
        else {
          // tile width with the least (waves x width), the narrower tiles paying their lower flop/byte
          const int sms = device_sms() > 0 ? device_sms() : 148;
          const int units = use_pair ? sms / 2 : sms;
          const int mb = use_pair ? (a.m + 255) / 256 : (a.m + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
          auto cost = [&](int b, double penalty) {
            const int t = mb * ((a.n + b - 1) / b);
            return (double)((t + units - 1) / units) * b * penalty;
          };
          double best = cost(256, 1.0);
          bn = 256;
          // 192-wide tiles move 14 % more operand bytes per flop; steady-state measurements (tools/gemm_roles.py: QKV at
          // M=5264, 7 waves of 192 vs 6 of 256: 75.4 vs 73.3 us) put their break-even at 1.2x
          if (a.n % 192 == 0 && !a.conv && !a.colstats && cost(192, 1.2) < best) { best = cost(192, 1.2); bn = 192; }
          if (cost(128, 1.25) < best) { best = cost(128, 1.25); bn = 128; }
          if (a.n % 256 != 0 && bn == 256 && a.n % 128 == 0 && a.n < 256) bn = 128;
        }
      }
      {
        // one or two K blocks, more than 128 bf16 output columns, many rows: CTA pairs on 256-wide tiles whose eight epilogue
        // warps hand 32 x 64 tiles to TMA stores (MV_GEMM_SKINNY_TMA=0 keeps the 128-wide two-CTA-per-SM schedule)
        static const int skinny_env = [] { const char* e = getenv("MV_GEMM_SKINNY_TMA"); return e ? atoi(e) : 1; }();
        const int kb_ = (a.k + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
        if (skinny_env != 0 && a.block_n == 0 && pair_legal && a.reserved2 != 1 && kb_ <= 2 && a.n > 128 && a.m >= 4096 &&
            a.out_f32 == 0 && a.out && !a.resid && !a.aux && !a.colstats && a.rows_per_group == 0 && a.kskip_end == 0 &&
            a.ldo % 8 == 0 &&
            (a.act == MV_ACT_NONE || a.act == MV_ACT_RELU ||
             (a.act == MV_ACT_GATE_MASK && a.in2 && a.ldin2 % 4 == 0 && (reinterpret_cast<uintptr_t>(a.in2) & 7) == 0)))
          return launch_gemm<256, MV_GEMM_LINEAR_TMA_BF16_W8, true>(a, stream);
      }
      {
        // epilogue-bound shapes (one or two K blocks, or a narrow N over many rows): two CTAs per SM
        const int kblocks = (a.k + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
        const long long tiles128 = (long long)((a.m + 127) / 128) * ((a.n + 127) / 128);
        const bool many = tiles128 >= 2ll * (device_sms() > 0 ? device_sms() : 148);
        if (a.block_n == 0 && many && !use_pair) {
          if (kblocks <= 2 && a.n >= 128) return launch_gemm<128, MV_GEMM_LINEAR, false, true>(a, stream);
          if (bn == 64) return launch_gemm<64, MV_GEMM_LINEAR, false, true>(a, stream);
          if (bn == 32) return launch_gemm<32, MV_GEMM_LINEAR, false, true>(a, stream);
        }
      }
      if (use_pair && bn == 192) return launch_gemm<192, MV_GEMM_LINEAR, true>(a, stream);
      if (bn == 192) return launch_gemm<192, MV_GEMM_LINEAR>(a, stream);
      if (use_pair && bn == 256) {
        // fp32 residual-stream update (attn.proj / fc2 with LayerScale): TMA epilogue. MV_GEMM_EPI_TMA=0 keeps the old one.
        static const int epi_env = [] { const char* e = getenv("MV_GEMM_EPI_TMA"); return e ? atoi(e) : 1; }();
        mv_gemm_args tail;
        const int keep = a.block_n == 0 ? plan_tail(a, &tail) : 0;  // (an explicit block_n pins the single-launch schedule)
        int rc;
        if (epi_env != 0 && a.out_f32 == 1 && a.resid && !a.aux && a.act == MV_ACT_NONE && a.rows_per_group == 0 && a.n % 32 == 0 &&
            a.ldo % 4 == 0 && a.ldr % 4 == 0)
          rc = launch_gemm<256, MV_GEMM_LINEAR_TMA, true>(a, stream, keep);
        else if (epi_env != 0 && epi_env != 2 && a.out_f32 == 0 && a.out && !a.resid && !a.aux && !a.colstats &&
                 a.rows_per_group == 0 && (a.act == MV_ACT_NONE || a.act == MV_ACT_RELU) && a.ldo % 8 == 0)
          rc = launch_gemm<256, MV_GEMM_LINEAR_TMA_BF16, true>(a, stream, keep);
        else
          rc = launch_gemm<256, MV_GEMM_LINEAR, true>(a, stream, keep);
        if (rc != MV_OK || keep == 0) return rc;
        return mv_gemm_bf16(&tail, stream_);  // the last partial wave as 128 x 128 single-CTA tiles
      }
      switch (bn) {
        case 16: return launch_gemm<16, MV_GEMM_LINEAR>(a, stream);
        case 32: return launch_gemm<32, MV_GEMM_LINEAR>(a, stream);
        case 64: return launch_gemm<64, MV_GEMM_LINEAR>(a, stream);
        case 128: return launch_gemm<128, MV_GEMM_LINEAR>(a, stream);
        case 256: return launch_gemm<256, MV_GEMM_LINEAR>(a, stream);
        default: set_error("mv_gemm_bf16: unsupported block_n %d", bn); return MV_ERR_ARG;
      }
    }
    default:
      set_error("mv_gemm_bf16: unknown mode %d", a.mode);
      return MV_ERR_ARG;
  }
}
