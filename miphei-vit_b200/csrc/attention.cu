// Multi-head self-attention forward for short sequences (N <= 448 tokens, head_dim 64) on tcgen05 / TMEM.
//
// Replaces timm Attention.forward's F.scaled_dot_product_attention(q, k, v) (no mask, scale 1/8, dropout 0; the ViT is
// created at src/generators/foundation_models.py:53-57) including the reshape(B,N,3,H,64).permute(2,0,3,1,4) split
// and the transpose(1,2).reshape(B,N,C) merge: it reads q/k/v straight from the fused qkv rows [B*N, 3*D] and writes
// token-major O [B*N, D].
//
// One CTA per (batch, head, 128-query tile); the whole key range of the head is resident, so softmax is exact and
// single pass (no running rescale):
//   thread 0        : TMA loads Q [128,64], K [Np,64], V [Np,64] (128-byte swizzle) -> S = Q K^T  (tcgen05.mma, SS)
//   warps 1..4      : one query row per thread: row max, p = exp2((s - max) * scale*log2e), row sum;
//                     P (bf16) written back into TMEM over S (tcgen05.st)
//   thread 0        : O = P V  (tcgen05.mma, A from TMEM, V as an MN-major smem operand)
//   warps 1..4      : O / rowsum -> bf16 -> global; log-sum-exp saved for the backward pass
// TMEM: S occupies Np fp32 columns (P aliases its first Np/2), O the 64 columns at 448.
#include "mv_host.h"
#include "mv_ptx.cuh"

namespace mv {

constexpr int ATT_THREADS = 160;
constexpr int ATT_MAX_KEYS = 448;
constexpr int ATT_O_COL = 448;

struct AttnDev {
  int n_tok;     // tokens per image
  int key_pad;   // n_tok rounded up to 16
  int kv_box;    // key_pad / 2 rows per TMA box
  int heads, dim;
  int q_tiles;
  float scale_log2e;
  float scale;
  __nv_bfloat16* out;
  long long ldo;
  float* lse;  // [B, heads, n_tok] or null
};

__global__ void __launch_bounds__(ATT_THREADS, 2) attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q,
                                                               const __grid_constant__ CUtensorMap tmap_kv,
                                                               const AttnDev p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t kv_bytes = p.key_pad * 128;
  const uint32_t sQ = smem_base;
  const uint32_t sK = sQ + 128 * 128;
  const uint32_t sV = sK + kv_bytes;
  const uint32_t bar_base = sV + kv_bytes;
  const uint32_t bar_qk = bar_base, bar_v = bar_base + 8, bar_s = bar_base + 16, bar_p = bar_base + 24,
                 bar_o = bar_base + 32;
  const uint32_t tmem_slot = bar_base + 40;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (smem_base - smem_u32(smem_raw)) + 128 * 128 + 2 * kv_bytes + 40);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x % p.q_tiles;
  const int bh = blockIdx.x / p.q_tiles;
  const int h = bh % p.heads;
  const int b = bh / p.heads;
  const int row0 = b * p.n_tok;  // first token row of this image
  const int q0 = qt * 128;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    fence_barrier_init();
    // loads first: they overlap the (possibly blocking) TMEM allocation below
    mbar_expect_tx(bar_qk, 128 * 128 + kv_bytes);
    tma_load_2d(sQ, &tmap_q, bar_qk, h * 64, row0 + q0);
    tma_load_2d(sK, &tmap_kv, bar_qk, p.dim + h * 64, row0);
    tma_load_2d(sK + p.kv_box * 128, &tmap_kv, bar_qk, p.dim + h * 64, row0 + p.kv_box);
    mbar_expect_tx(bar_v, kv_bytes);
    tma_load_2d(sV, &tmap_kv, bar_v, 2 * p.dim + h * 64, row0);
    tma_load_2d(sV + p.kv_box * 128, &tmap_kv, bar_v, 2 * p.dim + h * 64, row0 + p.kv_box);
  }
  __syncwarp();
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ---- S = Q K^T : [128, key_pad], K = 64 (4 UMMA k-steps), keys in chunks of <= 256
      mbar_wait(bar_qk, 0);
      tc_fence_after();
      const uint64_t dq = umma_desc_sw128(sQ);
      for (int n0 = 0; n0 < p.key_pad; n0 += 256) {
        const int nn = p.key_pad - n0 < 256 ? p.key_pad - n0 : 256;
        const uint32_t idesc = umma_idesc_bf16(128, nn);
        const uint64_t dk = umma_desc_sw128(sK + n0 * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + n0, dq + 2 * k, dk + 2 * k, idesc, k != 0);
      }
      umma_commit(bar_s);
      // ---- O = P V : [128, 64], K = key_pad (P from TMEM, 8 columns per 16 keys; V MN-major, 2 KB per 16 keys)
      mbar_wait(bar_v, 0);
      mbar_wait(bar_p, 0);
      tc_fence_after();
      const uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);
      const int ksteps = p.key_pad / 16;
      for (int k = 0; k < ksteps; ++k) {
        const uint64_t dv = umma_desc_sw128(sV + k * 2048, 1024, 1024);
        umma_bf16_ts(tmem_base + ATT_O_COL, tmem_base + k * 8, dv, idesc_pv, k != 0);
      }
      umma_commit(bar_o);
    }
  } else {
    // ---- softmax + epilogue: thread <-> query row <-> TMEM lane
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int n_tok = p.n_tok;
    const int full = p.key_pad / 32;       // full 32-column chunks
    const bool tail16 = (p.key_pad & 31) != 0;
    mbar_wait(bar_s, 0);
    tc_fence_after();
    float mx = -INFINITY;
    for (int c = 0; c < full; ++c) {
      uint32_t v[32];
      tmem_ld32(trow + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c * 32 + j < n_tok) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    if (tail16) {
      uint32_t v[16];
      tmem_ld16(trow + full * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (full * 32 + j < n_tok) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    const float mxs = mx * p.scale_log2e;
    float sum = 0.f;
    for (int c = 0; c < full; ++c) {
      uint32_t v[32];
      tmem_ld32(trow + c * 32, v);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int col = c * 32 + 2 * j;
        float e0 = col < n_tok ? exp2f(__uint_as_float(v[2 * j]) * p.scale_log2e - mxs) : 0.f;
        float e1 = col + 1 < n_tok ? exp2f(__uint_as_float(v[2 * j + 1]) * p.scale_log2e - mxs) : 0.f;
        pk[j] = pack_bf16x2(e0, e1);
        const float2 rr = unpack_bf16x2(pk[j]);  // the sum must match what the MMA will see
        sum += rr.x + rr.y;
      }
      tmem_st16(trow + c * 16, pk);
    }
    if (tail16) {
      uint32_t v[16];
      tmem_ld16(trow + full * 32, v);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int col = full * 32 + 2 * j;
        float e0 = col < n_tok ? exp2f(__uint_as_float(v[2 * j]) * p.scale_log2e - mxs) : 0.f;
        float e1 = col + 1 < n_tok ? exp2f(__uint_as_float(v[2 * j + 1]) * p.scale_log2e - mxs) : 0.f;
        pk[j] = pack_bf16x2(e0, e1);
        const float2 rr = unpack_bf16x2(pk[j]);
        sum += rr.x + rr.y;
      }
#pragma unroll
      for (int j = 8; j < 16; ++j) pk[j] = 0u;
      // 8 valid columns; the 8 zero columns land in [key_pad/2, key_pad/2 + 8) which is still inside S (never read)
      tmem_st16(trow + full * 16, pk);
    }
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(bar_p);

    mbar_wait(bar_o, 0);
    tc_fence_after();
    const int q = q0 + r;
    const float inv = 1.f / sum;
    uint32_t o0[32], o1[32];
    tmem_ld32(trow + ATT_O_COL, o0);
    tmem_ld32(trow + ATT_O_COL + 32, o1);
    tmem_ld_wait();
    if (q < n_tok) {
      __nv_bfloat16* orow = p.out + (long long)(row0 + q) * p.ldo + h * 64;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(o0[8 * j + 0]) * inv, __uint_as_float(o0[8 * j + 1]) * inv);
        u.y = pack_bf16x2(__uint_as_float(o0[8 * j + 2]) * inv, __uint_as_float(o0[8 * j + 3]) * inv);
        u.z = pack_bf16x2(__uint_as_float(o0[8 * j + 4]) * inv, __uint_as_float(o0[8 * j + 5]) * inv);
        u.w = pack_bf16x2(__uint_as_float(o0[8 * j + 6]) * inv, __uint_as_float(o0[8 * j + 7]) * inv);
        reinterpret_cast<uint4*>(orow)[j] = u;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(o1[8 * j + 0]) * inv, __uint_as_float(o1[8 * j + 1]) * inv);
        u.y = pack_bf16x2(__uint_as_float(o1[8 * j + 2]) * inv, __uint_as_float(o1[8 * j + 3]) * inv);
        u.z = pack_bf16x2(__uint_as_float(o1[8 * j + 4]) * inv, __uint_as_float(o1[8 * j + 5]) * inv);
        u.w = pack_bf16x2(__uint_as_float(o1[8 * j + 6]) * inv, __uint_as_float(o1[8 * j + 7]) * inv);
        reinterpret_cast<uint4*>(orow)[4 + j] = u;
      }
      if (p.lse) p.lse[((long long)b * p.heads + h) * n_tok + q] = mx * p.scale + logf(sum);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
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
  p.scale = scale;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.lse = lse;
  const uint64_t rows = (uint64_t)batch * n_tok;
  const CUtensorMap* tq = get_tmap_2d_bf16(qkv, rows, 3ull * p.dim, ldqkv, 128);
  const CUtensorMap* tkv = get_tmap_2d_bf16(qkv, rows, 3ull * p.dim, ldqkv, p.kv_box);
  if (!tq || !tkv) return MV_ERR_ARG;
  const int smem = 128 * 128 + 2 * p.key_pad * 128 + 1024 + 64;
  static int smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attn_fwd): %s", cudaGetErrorString(e));
      return (int)e;
    }
    smem_set = smem;
  }
  const int grid = batch * heads * p.q_tiles;
  attn_fwd_kernel<<<grid, ATT_THREADS, smem, stream>>>(*tq, *tkv, p);
  MV_CHECK_LAUNCH("attn_fwd");
  return MV_OK;
}
