// Multi-head self-attention forward for short sequences (16 <= N <= 448 tokens, head_dim 64) on tcgen05 / TMEM.
//
// Replaces timm Attention.forward's F.scaled_dot_product_attention(q, k, v) (no mask, scale 1/8, dropout 0; the ViT is
// created at src/generators/foundation_models.py:53-57) including the reshape(B,N,3,H,64).permute(2,0,3,1,4) split
// and the transpose(1,2).reshape(B,N,C) merge: it reads q/k/v straight from the fused qkv rows [B*N, 3*D] and writes
// token-major O [B*N, D].
//
// Persistent kernel, one CTA per SM; work item = (image, head, 128-query tile), each CTA takes a contiguous range of
// items so the K/V of a head are loaded once for its query tiles.  The whole key range of a head is resident, so the
// softmax is exact and single pass (no running rescale):
//   warp 0 (1 thread) : TMA producer — Q tiles (2 stages) and K/V (1 or 2 stages), 128-byte swizzle
//   warp 1 (1 thread) : tcgen05.mma issuer — S = Q K^T (SS), then O = P V (A = P from TMEM, B = V MN-major from smem)
//   warps 2..9        : softmax + epilogue, two warps per TMEM lane quadrant (each takes half of the key columns):
//                       row max -> p = exp2((s - max) * scale*log2e) -> bf16 P written back over S in TMEM -> O / rowsum
// TMEM: S occupies key_pad fp32 columns (P aliases its first key_pad/2), O is double buffered at columns 384 and 448.
#include "mv_host.h"
#include "mv_ptx.cuh"

namespace mv {

constexpr int ATT_THREADS = 320;
constexpr int ATT_MAX_KEYS = 384;   // S columns [0, 384); O buffers at 384 / 448
constexpr int ATT_O_COL = 384;
constexpr int ATT_MAXC = 6;         // 32-column chunks per half row (384 / 32 / 2)

struct AttnDev {
  int n_tok;     // tokens per image
  int key_pad;   // n_tok rounded up to 16
  int kv_box;    // key_pad / 2 rows per TMA box
  int heads, dim;
  int q_tiles;
  int total_tiles;
  int kv_stages;
  float scale_log2e;
  float scale;
  __nv_bfloat16* out;
  long long ldo;
  float* lse;  // [B, heads, n_tok] or null
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(ATT_THREADS, 1) attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q,
                                                                  const __grid_constant__ CUtensorMap tmap_kv,
                                                                  const AttnDev p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t kv_bytes = p.key_pad * 128;
  const uint32_t sQ = smem_base;                       // 2 x 16 KB
  const uint32_t sKV = sQ + 2 * 16384;                 // kv_stages x (K | V)
  const uint32_t misc_off = 2 * 16384 + p.kv_stages * 2 * kv_bytes;
  const uint32_t bar_base = smem_base + misc_off;
  auto q_full = [&](int s) { return bar_base + 8u * s; };
  auto q_empty = [&](int s) { return bar_base + 8u * (2 + s); };
  auto kv_full = [&](int s) { return bar_base + 8u * (4 + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (6 + s); };
  const uint32_t bar_s = bar_base + 8u * 8, bar_p = bar_base + 8u * 9;
  auto bar_o = [&](int s) { return bar_base + 8u * (10 + s); };
  auto o_empty = [&](int s) { return bar_base + 8u * (12 + s); };
  const uint32_t tmem_slot = bar_base + 8u * 14;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + misc_off + 8 * 14);
  float* xch_max = reinterpret_cast<float*>(smem_gen + misc_off + 128);  // [2][128]
  float* xch_sum = xch_max + 256;                                        // [2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // contiguous range of work items for this CTA
  const int t0 = (int)((long long)p.total_tiles * blockIdx.x / gridDim.x);
  const int t1 = (int)((long long)p.total_tiles * (blockIdx.x + 1) / gridDim.x);

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    for (int s = 0; s < 2; ++s) {
      mbar_init(q_full(s), 1);
      mbar_init(q_empty(s), 1);
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
      mbar_init(bar_o(s), 1);
      mbar_init(o_empty(s), 256);
    }
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 256);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int qs = 0, ks = 0, prev_bh = -1;
      uint32_t qph = 0, kph = 0;
      for (int t = t0; t < t1; ++t) {
        const int bh = t / p.q_tiles, qt = t - bh * p.q_tiles;
        const int b = bh / p.heads, h = bh - b * p.heads;
        const int row0 = b * p.n_tok;
        if (bh != prev_bh) {
          prev_bh = bh;
          mbar_wait(kv_empty(ks), kph ^ 1);
          const uint32_t sK = sKV + ks * 2 * kv_bytes, sV = sK + kv_bytes;
          mbar_expect_tx(kv_full(ks), 2 * kv_bytes);
          tma_load_2d(sK, &tmap_kv, kv_full(ks), p.dim + h * 64, row0);
          tma_load_2d(sK + p.kv_box * 128, &tmap_kv, kv_full(ks), p.dim + h * 64, row0 + p.kv_box);
          tma_load_2d(sV, &tmap_kv, kv_full(ks), 2 * p.dim + h * 64, row0);
          tma_load_2d(sV + p.kv_box * 128, &tmap_kv, kv_full(ks), 2 * p.dim + h * 64, row0 + p.kv_box);
          if (++ks == p.kv_stages) { ks = 0; kph ^= 1; }
        }
        mbar_wait(q_empty(qs), qph ^ 1);
        mbar_expect_tx(q_full(qs), 16384);
        tma_load_2d(sQ + qs * 16384, &tmap_q, q_full(qs), h * 64, row0 + qt * 128);
        if (++qs == 2) { qs = 0; qph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      int qs = 0, ks = 0, os = 0, prev_bh = -1, cur_ks = 0;
      uint32_t qph = 0, kph = 0, oph = 0, tph = 0;
      const uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);
      const int ksteps = p.key_pad / 16;
      for (int t = t0; t < t1; ++t) {
        const int bh = t / p.q_tiles;
        if (bh != prev_bh) {
          prev_bh = bh;
          mbar_wait(kv_full(ks), kph);
          cur_ks = ks;
          if (++ks == p.kv_stages) { ks = 0; kph ^= 1; }
        }
        const uint32_t sK = sKV + cur_ks * 2 * kv_bytes, sV = sK + kv_bytes;
        mbar_wait(q_full(qs), qph);
        tc_fence_after();
        // ---- S = Q K^T : [128, key_pad], K = 64 (4 UMMA k-steps), keys in chunks of <= 256
        const uint64_t dq = umma_desc_sw128(sQ + qs * 16384);
        for (int n0 = 0; n0 < p.key_pad; n0 += 256) {
          const int nn = p.key_pad - n0 < 256 ? p.key_pad - n0 : 256;
          const uint32_t idesc = umma_idesc_bf16(128, nn);
          const uint64_t dk = umma_desc_sw128(sK + n0 * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + n0, dq + 2 * k, dk + 2 * k, idesc, k != 0);
        }
        umma_commit(bar_s);
        umma_commit(q_empty(qs));
        // ---- O = P V : [128, 64], K = key_pad (P from TMEM, 8 columns per 16 keys; V MN-major, 2 KB per 16 keys)
        mbar_wait(bar_p, tph);
        mbar_wait(o_empty(os), oph ^ 1);
        tc_fence_after();
        const uint32_t d_o = tmem_base + ATT_O_COL + os * 64;
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t dv = umma_desc_sw128(sV + k * 2048, 1024, 1024);
          umma_bf16_ts(d_o, tmem_base + k * 8, dv, idesc_pv, k != 0);
        }
        umma_commit(bar_o(os));
        if (t + 1 == t1 || (t + 1) / p.q_tiles != bh) umma_commit(kv_empty(cur_ks));
        if (++qs == 2) { qs = 0; qph ^= 1; }
        if (++os == 2) { os = 0; oph ^= 1; }
        tph ^= 1;
      }
    }
  } else {
    // ===================== softmax + epilogue =====================
    const int sw = warp - 2;         // 0..7
    const int quad = warp & 3;       // TMEM lane quadrant this warp may access
    const int half = sw >> 2;        // which half of the key columns / O columns
    const int r = quad * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int n_tok = p.n_tok;
    const int full32 = p.key_pad / 32;
    const bool tail16 = (p.key_pad & 31) != 0;
    const int h0 = (full32 + 1) / 2;
    const int c_begin = half == 0 ? 0 : h0;
    const int c_end = half == 0 ? h0 : full32;
    const bool my_tail = tail16 && half == 1;
    int os = 0;
    uint32_t oph = 0, tph = 0;
    for (int t = t0; t < t1; ++t) {
      const int bh = t / p.q_tiles, qt = t - bh * p.q_tiles;
      const int b = bh / p.heads, h = bh - b * p.heads;
      mbar_wait(bar_s, tph);
      tc_fence_after();
      // ---- pass 1: row max over this warp's columns
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = c_begin; c < c_end; ++c) {
        uint32_t v[32];
        tmem_ld32(trow + c * 32, v);
        tmem_ld_wait();
        if (c * 32 + 32 <= n_tok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c * 32 + j < n_tok) mx = fmaxf(mx, __uint_as_float(v[j]));
        }
      }
      if (my_tail) {
        uint32_t v[16];
        tmem_ld16(trow + full32 * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (full32 * 32 + j < n_tok) mx = fmaxf(mx, __uint_as_float(v[j]));
      }
      xch_max[half * 128 + r] = mx;
      named_bar_sync(1, 256);
      mx = fmaxf(mx, xch_max[(half ^ 1) * 128 + r]);
      const float mxs = mx * p.scale_log2e;
      // ---- pass 2: probabilities, kept in registers until every warp has finished reading S
      uint32_t pk[ATT_MAXC][16];
      uint32_t pkt[8];
      float sum = 0.f;
#pragma unroll
      for (int ci = 0; ci < ATT_MAXC; ++ci) {
        const int c = c_begin + ci;
        if (c < c_end) {
          uint32_t v[32];
          tmem_ld32(trow + c * 32, v);
          tmem_ld_wait();
          const bool fullc = c * 32 + 32 <= n_tok;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float e0 = ex2_approx(__uint_as_float(v[2 * j]) * p.scale_log2e - mxs);
            float e1 = ex2_approx(__uint_as_float(v[2 * j + 1]) * p.scale_log2e - mxs);
            if (!fullc) {
              if (c * 32 + 2 * j >= n_tok) e0 = 0.f;
              if (c * 32 + 2 * j + 1 >= n_tok) e1 = 0.f;
            }
            sum += e0 + e1;
            pk[ci][j] = pack_bf16x2(e0, e1);
          }
        }
      }
      if (my_tail) {
        uint32_t v[16];
        tmem_ld16(trow + full32 * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int col = full32 * 32 + 2 * j;
          const float e0 = col < n_tok ? ex2_approx(__uint_as_float(v[2 * j]) * p.scale_log2e - mxs) : 0.f;
          const float e1 = col + 1 < n_tok ? ex2_approx(__uint_as_float(v[2 * j + 1]) * p.scale_log2e - mxs) : 0.f;
          sum += e0 + e1;
          pkt[j] = pack_bf16x2(e0, e1);
        }
      }
      xch_sum[half * 128 + r] = sum;
      named_bar_sync(2, 256);  // all reads of S are done: P may now overwrite it
#pragma unroll
      for (int ci = 0; ci < ATT_MAXC; ++ci) {
        const int c = c_begin + ci;
        if (c < c_end) tmem_st16(trow + c * 16, pk[ci]);
      }
      if (my_tail) {
        uint32_t z[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) { z[j] = pkt[j]; z[8 + j] = 0u; }
        tmem_st16(trow + full32 * 16, z);  // upper 8 columns fall beyond key_pad/2, inside the dead part of S
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_p);

      // ---- epilogue: this warp normalises and stores 32 of the 64 output columns
      mbar_wait(bar_o(os), oph);
      tc_fence_after();
      const float total = sum + xch_sum[(half ^ 1) * 128 + r];
      const float inv = 1.f / total;
      uint32_t o[32];
      tmem_ld32(trow + ATT_O_COL + os * 64 + half * 32, o);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(o_empty(os));
      const int q = qt * 128 + r;
      if (q < n_tok) {
        __nv_bfloat16* orow = p.out + (long long)(b * n_tok + q) * p.ldo + h * 64 + half * 32;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[8 * j + 0]) * inv, __uint_as_float(o[8 * j + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv);
          reinterpret_cast<uint4*>(orow)[j] = u;
        }
        if (p.lse && half == 0) p.lse[((long long)b * p.heads + h) * n_tok + q] = mx * p.scale + logf(total);
      }
      if (++os == 2) { os = 0; oph ^= 1; }
      tph ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace mv

// qkv bf16 [B*N, 3*D] rows = tokens (q | k | v, each heads x 64); out bf16 [B*N, D]; lse fp32 [B, heads, N] or NULL.
extern "C" int mv_attn_fwd(const void* qkv, int64_t ldqkv, void* out, int64_t ldo, float* lse, int batch, int n_tok,
                           int heads, float scale, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(qkv && out && batch > 0 && heads > 0, "mv_attn_fwd: null/empty");
  MV_CHECK_ARG(n_tok >= 16 && n_tok <= ATT_MAX_KEYS, "mv_attn_fwd: n_tok=%d outside [16, %d] (single-pass kernel)", n_tok,
               ATT_MAX_KEYS);
  MV_CHECK_ARG(ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "mv_attn_fwd: out alignment");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  AttnDev p;
  p.n_tok = n_tok;
  p.key_pad = (n_tok + 15) / 16 * 16;
  p.kv_box = p.key_pad / 2;
  p.heads = heads;
  p.dim = heads * 64;
  p.q_tiles = (n_tok + 127) / 128;
  p.total_tiles = batch * heads * p.q_tiles;
  p.scale = scale;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.lse = lse;
  const int misc = 128 + 4 * 128 * 4 + 1024;
  const int kvb = 2 * p.key_pad * 128;
  p.kv_stages = (2 * 16384 + 2 * kvb + misc <= 227 * 1024) ? 2 : 1;
  const int smem = 2 * 16384 + p.kv_stages * kvb + misc;
  const uint64_t rows = (uint64_t)batch * n_tok;
  const CUtensorMap* tq = get_tmap_2d_bf16(qkv, rows, 3ull * p.dim, ldqkv, 128);
  const CUtensorMap* tkv = get_tmap_2d_bf16(qkv, rows, 3ull * p.dim, ldqkv, p.kv_box);
  if (!tq || !tkv) return MV_ERR_ARG;
  static int smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attn_fwd): %s", cudaGetErrorString(e));
      return (int)e;
    }
    smem_set = smem;
  }
  int grid = device_sms() > 0 ? device_sms() : 148;
  if (grid > p.total_tiles) grid = p.total_tiles;
  attn_fwd_kernel<<<grid, ATT_THREADS, smem, stream>>>(*tq, *tkv, p);
  MV_CHECK_LAUNCH("attn_fwd");
  return MV_OK;
}
