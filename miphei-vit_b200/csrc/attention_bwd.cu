// Self-attention backward (head_dim 64, any N) on tcgen05 / TMEM: dQ, dK, dV from dO, the saved q/k/v rows and the
// saved log-sum-exp.  Autograd counterpart of timm Attention.forward's F.scaled_dot_product_attention.
//
//   D[q]   = sum_d dO[q,d] * O[q,d]                                   (mv_attn_bwd_prep, HBM-bound)
//   P      = exp(S * scale - LSE),  S = Q K^T                          (recomputed, never stored)
//   dP     = dO V^T,   dS = P .* (dP - D) * scale
//   dV     = P^T dO,   dK = dS^T Q,   dQ = dS K
//
// One templated persistent kernel, instantiated twice (no atomics, every output written once):
//   DKV = true : work item = (image, head, 128-key block); the key block's K, V rows are the TMEM-lane ("row")
//                operands, the kernel loops over 128-query tiles and accumulates dV, dK of the block in TMEM.
//   DKV = false: work item = (image, head, 128-query tile); Q, dO are the row operands, the loop runs over key blocks and
//                accumulates dQ.  (S and dP are recomputed by both instances: 7 MMAs per tile pair instead of 5.)
// Per (128-row block, 64-column tile) pair: X = R1 C1^T and Y = R2 C2^T (SS MMAs) -> the math warps turn X into P and Y
// into dS (bf16, written back over X / Y in TMEM) -> acc1 += P C2 (DKV only) and acc2 += dS C1 with the A operand read
// from TMEM and the C tiles re-read from the SAME smem tiles as MN-major operands.
// Two schedules of this: attn_bwd_kernel (round 1, MV_ATTN_BWD_V=1) runs two CTAs per SM (64 KB smem, 256 TMEM columns
// each) so that one CTA's exponentials overlap the other CTA's MMAs; attn_bwd2_kernel (default, further down) runs one CTA
// per SM with X / Y triple-buffered, two MMA-issuing threads and sixteen math warps.
#include <stdlib.h>

#include "mv_host.h"
#include "mv_ptx.cuh"

namespace mv {

constexpr int ATB_THREADS = 320;      // warp 0 TMA, warp 1 MMA, warps 2..9 softmax-backward math: TWO threads per row
                                      // (TMEM lane quadrant = warp % 4, column half = (warp - 2) / 4), 32 columns each
constexpr int ATB_RTILE = 128 * 128;   // bytes of one [128 x 64] bf16 row-operand tile
constexpr int ATB_CTILE = 64 * 128;    // bytes of one [64 x 64] bf16 column-operand tile
constexpr int ATB_TMEM = 256;          // X 64 | Y 64 | acc1 64 | acc2 64 : two CTAs per SM share the 512 columns

struct AttnBwdDev {
  int n_tok, heads, dim;
  int rblocks;  // ceil(n_tok / 128): outer (row) blocks
  int cblocks;  // ceil(n_tok / 64): inner (column) blocks
  int total_items;
  float scale, scale_log2e;
  const float* lse;    // [B, heads, n_tok]
  const float* dsum;   // [B, heads, n_tok]
  __nv_bfloat16* dqkv; // [B*n_tok, 3*dim]
  long long lddqkv;
  long long* prof;     // diagnostics (MV_GEMM_PROFILE builds only): per CTA 16 x int64 cycle sums, see tools/attn_roles.py
};

#ifndef MV_GEMM_PROFILE
#define MV_GEMM_PROFILE 0
#endif
constexpr bool kAtbProf = MV_GEMM_PROFILE != 0;
extern long long* g_attn_prof;

__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void nbar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Two CTAs are resident per SM (64 KB smem, 256 TMEM columns, <= 170 registers each): while one CTA's warps do the
// exponentials of a tile pair, the other CTA's MMAs own the tensor core — the overlap needs no intra-kernel ping-pong.
template <bool DKV>
__global__ void __launch_bounds__(ATB_THREADS, 2) attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv,
                                                                  const __grid_constant__ CUtensorMap tmap_do,
                                                                  const __grid_constant__ CUtensorMap tmap_qkv64,
                                                                  const __grid_constant__ CUtensorMap tmap_do64,
                                                                  const AttnBwdDev p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // smem: R1 | R2 (single stage), C tiles 2 stages x (C1 | C2), barriers, lse/D staging (2 buffers)
  const uint32_t sR = smem_base, sC = smem_base + 2 * ATB_RTILE;
  const uint32_t misc_off = 2 * ATB_RTILE + 4 * ATB_CTILE;
  const uint32_t bar_base = smem_base + misc_off;
  const uint32_t r_full = bar_base, r_empty = bar_base + 8;
  auto c_full = [&](int s) { return bar_base + 8u * (2 + s); };
  auto c_empty = [&](int s) { return bar_base + 8u * (4 + s); };
  const uint32_t bar_xy = bar_base + 8u * 6, bar_pd = bar_base + 8u * 7, bar_acc = bar_base + 8u * 8,
                 acc_empty = bar_base + 8u * 9;
  const uint32_t tmem_slot = bar_base + 8u * 10;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + misc_off + 8 * 10);
  // DKV: per-column (query) statistics of the current inner tile, one private copy per math warp: [8][lse * log2e 64 | D 64]
  float* s_stat = reinterpret_cast<float*>(smem_gen + misc_off + 128);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = (int)((long long)p.total_items * blockIdx.x / gridDim.x);
  const int t1 = (int)((long long)p.total_items * (blockIdx.x + 1) / gridDim.x);
  const int nrb = p.rblocks, ncb = p.cblocks;
  constexpr uint32_t COL_X = 0, COL_Y = 64, COL_A1 = 128, COL_A2 = 192;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_do);
    tma_prefetch_desc(&tmap_qkv64);
    tma_prefetch_desc(&tmap_do64);
    mbar_init(r_full, 1);
    mbar_init(r_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(c_full(s), 1);
      mbar_init(c_empty(s), 1);
    }
    mbar_init(bar_xy, 1);
    mbar_init(bar_pd, 8);      // one arrival per math warp
    mbar_init(bar_acc, 1);
    mbar_init(acc_empty, 8);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, ATB_TMEM);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  griddep_sync();  // PDL: prologue overlapped the previous kernel; its results are visible from here on

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int cs = 0;
      uint32_t rph = 0, cph = 0;
      for (int t = t0; t < t1; ++t) {
        const int bh = t / nrb, ob = t - bh * nrb;  // outer block: 128 keys (DKV) or 128 queries
        const int b = bh / p.heads, h = bh - b * p.heads;
        const int row0 = b * p.n_tok;
        mbar_wait(r_empty, rph ^ 1);
        mbar_expect_tx(r_full, 2 * ATB_RTILE);
        if (DKV) {
          tma_load_2d(sR, &tmap_qkv, r_full, p.dim + h * 64, row0 + ob * 128);                  // K_j
          tma_load_2d(sR + ATB_RTILE, &tmap_qkv, r_full, 2 * p.dim + h * 64, row0 + ob * 128);  // V_j
        } else {
          tma_load_2d(sR, &tmap_qkv, r_full, h * 64, row0 + ob * 128);                          // Q_i
          tma_load_2d(sR + ATB_RTILE, &tmap_do, r_full, h * 64, row0 + ob * 128);               // dO_i
        }
        rph ^= 1;
        for (int ib = 0; ib < ncb; ++ib) {
          mbar_wait(c_empty(cs), cph ^ 1);
          mbar_expect_tx(c_full(cs), 2 * ATB_CTILE);
          const uint32_t c1 = sC + cs * 2 * ATB_CTILE, c2 = c1 + ATB_CTILE;
          if (DKV) {
            tma_load_2d(c1, &tmap_qkv64, c_full(cs), h * 64, row0 + ib * 64);            // Q (64 queries)
            tma_load_2d(c2, &tmap_do64, c_full(cs), h * 64, row0 + ib * 64);             // dO
          } else {
            tma_load_2d(c1, &tmap_qkv64, c_full(cs), p.dim + h * 64, row0 + ib * 64);      // K (64 keys)
            tma_load_2d(c2, &tmap_qkv64, c_full(cs), 2 * p.dim + h * 64, row0 + ib * 64);  // V
          }
          if (++cs == 2) { cs = 0; cph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      int cs = 0;
      uint32_t rph = 0, cph = 0, xph = 0, aph = 0;
      const uint32_t idesc_xy = umma_idesc_bf16(128, 64);
      const uint32_t idesc_acc = umma_idesc_bf16(128, 64, 0, 1);
      // all smem descriptors are built once: this single thread sits on the per-tile critical path
      // (XY MMAs -> math -> accumulate MMAs), so its instruction count per tile matters
      const uint64_t dr1 = umma_desc_sw128(sR), dr2 = umma_desc_sw128(sR + ATB_RTILE);
      const uint64_t dk1 = umma_desc_sw128(sC), dk2 = umma_desc_sw128(sC + ATB_CTILE);                            // K-major
      const uint64_t dm1 = umma_desc_sw128(sC, 1024, 1024), dm2 = umma_desc_sw128(sC + ATB_CTILE, 1024, 1024);  // MN-major
      constexpr uint64_t kStageStep = (2 * ATB_CTILE) >> 4;
      long long pm[4] = {0, 0, 0, 0};  // waiting for operands | issuing XY | waiting for P, dS | issuing the accumulations
      const long long pm_t0 = kAtbProf ? clock64() : 0;
      long long pt = pm_t0;
      auto lap = [&](int i) { if (kAtbProf) { const long long n = clock64(); pm[i] += n - pt; pt = n; } };
      for (int t = t0; t < t1; ++t) {
        mbar_wait(r_full, rph);
        for (int ib = 0; ib < ncb; ++ib) {
          mbar_wait(c_full(cs), cph);
          lap(0);
          tc_fence_after();
          const uint64_t so = cs ? kStageStep : 0;
          const uint64_t dc1 = dk1 + so, dc2 = dk2 + so;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + COL_X, dr1 + 2 * k, dc1 + 2 * k, idesc_xy, k != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + COL_Y, dr2 + 2 * k, dc2 + 2 * k, idesc_xy, k != 0);
          umma_commit(bar_xy);
          lap(1);
          mbar_wait(bar_pd, xph);
          if (ib == 0) mbar_wait(acc_empty, aph ^ 1);  // previous item's accumulators drained
          lap(2);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // K = 64 columns of this inner tile; 16 reduction rows = 2 KB per step
            // P / dS of column half h sit packed in the first 16 TMEM columns of that half's own 32-column range
            const uint32_t acol = (k >> 1) * 32 + (k & 1) * 8;
            if (DKV) umma_bf16_ts(tmem_base + COL_A1, tmem_base + COL_X + acol, dm2 + so + (2048 >> 4) * k, idesc_acc, (ib | k) != 0);
            umma_bf16_ts(tmem_base + COL_A2, tmem_base + COL_Y + acol, dm1 + so + (2048 >> 4) * k, idesc_acc, (ib | k) != 0);
          }
          umma_commit(c_empty(cs));
          if (ib == ncb - 1) {
            umma_commit(bar_acc);
            umma_commit(r_empty);
          }
          if (++cs == 2) { cs = 0; cph ^= 1; }
          xph ^= 1;
          lap(3);
        }
        rph ^= 1;
        aph ^= 1;
      }
      if (kAtbProf && p.prof) {
        long long* q = p.prof + 16ll * (blockIdx.x + (DKV ? 0 : gridDim.x));  // dK/dV instance first, dQ instance behind
        q[8] = pm[0]; q[9] = pm[1]; q[10] = pm[2]; q[11] = pm[3]; q[12] = clock64() - pm_t0;
      }
    }
  } else {
    // ===================== softmax-backward math + epilogue (4 warps, thread = row) =====================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;   // which 32 of the 64 tile columns (and of the 64 accumulator columns) this thread owns
    const int r = quad * 32 + lane;
    const int tid = threadIdx.x - 64;  // 0..255
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int n_tok = p.n_tok;
    uint32_t xph = 0, aph = 0;
    int itn = 0;
    long long pa[6] = {0, 0, 0, 0, 0, 0};  // statistics staging | wait X, Y | ld + math | st + arrive | item epilogue
    const long long pa_t0 = kAtbProf ? clock64() : 0;
    long long pt = pa_t0;
    auto lap = [&](int i) { if (kAtbProf) { const long long n = clock64(); pa[i] += n - pt; pt = n; } };
    for (int t = t0; t < t1; ++t) {
      const int bh = t / nrb, ob = t - bh * nrb;
      const int b = bh / p.heads, h = bh - b * p.heads;
      const long long vec0 = (long long)bh * n_tok;
      const int orow = ob * 128 + r;  // key (DKV) or query index of this thread's row
      float lse_r = 0.f, dsum_r = 0.f;
      if (!DKV && orow < n_tok) {
        lse_r = p.lse[vec0 + orow] * 1.4426950408889634f;
        dsum_r = p.dsum[vec0 + orow];
      }
      // DKV: per-column (query) statistics of an inner tile: every warp keeps its own copy (lane l fetches queries l and
      // l + 32, lse and D) ONE TILE AHEAD into registers and parks it in its private smem row — no CTA-wide barrier per tile.
      // (Measured neutral against the shared copy behind a named barrier, 213 vs 215 us at B = 32: the time a math warp
      // spends here, ~0.9 k of 4.8 k cycles per tile pair in tools/attn_bwd_roles.py, moves into "wait X, Y" — the tile
      // pair is a serial chain  math -> accumulate MMAs -> next X, Y MMAs  and only the SM's second CTA overlaps it.)
      float* my_stat = s_stat + (warp - 2) * 128;
      // RAW loads with clamped indices and no arithmetic or select on the result: a warp issues in order, so anything that
      // consumes the loaded value here would stall it for the whole L2 round trip; queries beyond the sequence are masked
      // in the math
      auto load_stat = [&](int ib, float (&v)[4]) {
        const int q0 = min(ib * 64 + lane, n_tok - 1), q1 = min(ib * 64 + 32 + lane, n_tok - 1);
        v[0] = __ldg(p.lse + vec0 + q0);
        v[1] = __ldg(p.lse + vec0 + q1);
        v[2] = __ldg(p.dsum + vec0 + q0);
        v[3] = __ldg(p.dsum + vec0 + q1);
      };
      float stat_next[4] = {0.f, 0.f, 0.f, 0.f};
      if (DKV) load_stat(0, stat_next);
      for (int ib = 0; ib < ncb; ++ib, ++itn) {
        const float* sl = my_stat;
        const float* sd = my_stat + 64;
        if (DKV) {
          __syncwarp();  // every lane has finished reading the previous tile's copy
          my_stat[lane] = stat_next[0] * 1.4426950408889634f; my_stat[32 + lane] = stat_next[1] * 1.4426950408889634f;
          my_stat[64 + lane] = stat_next[2]; my_stat[96 + lane] = stat_next[3];
          __syncwarp();
          if (ib + 1 < ncb) load_stat(ib + 1, stat_next);
        }
        lap(0);
        mbar_wait(bar_xy, xph);
        lap(1);
        tc_fence_after();
        // this thread's 32 columns [32*half, 32*half+32) of X and Y: one TMEM round trip; P / dS (bf16 pairs) go back into
        // the first 16 columns of the SAME range, so the two threads of a row never touch each other's columns
        uint32_t x[32], y[32], pp[16], pd[16];
        tmem_ld32(trow + COL_X + half * 32, x);
        tmem_ld32(trow + COL_Y + half * 32, y);
        tmem_ld_wait();
        // dS is formed WITHOUT the softmax scale (applied once per item to the accumulator in the epilogue), and tiles
        // that lie completely inside the sequence skip the per-element bounds predicates
        const bool interior = (ob * 128 + 127 < n_tok) && (ib * 64 + 63 < n_tok);  // CTA-uniform
        if (interior) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float pv[2], dv[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int col = half * 32 + 2 * j + e;
              const float l2 = DKV ? sl[col] : lse_r;
              const float dd = DKV ? sd[col] : dsum_r;
              const float pr = ex2a(__uint_as_float(x[2 * j + e]) * p.scale_log2e - l2);
              pv[e] = pr;
              dv[e] = pr * (__uint_as_float(y[2 * j + e]) - dd);
            }
            pp[j] = pack_bf16x2(pv[0], pv[1]);
            pd[j] = pack_bf16x2(dv[0], dv[1]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float pv[2], dv[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int col = half * 32 + 2 * j + e;
              const int icol = ib * 64 + col;
              const bool ok = orow < n_tok && icol < n_tok;
              const float l2 = DKV ? sl[col] : lse_r;
              const float dd = DKV ? sd[col] : dsum_r;
              const float pr = ok ? ex2a(__uint_as_float(x[2 * j + e]) * p.scale_log2e - l2) : 0.f;
              pv[e] = pr;
              dv[e] = pr * (__uint_as_float(y[2 * j + e]) - dd);
            }
            pp[j] = pack_bf16x2(pv[0], pv[1]);
            pd[j] = pack_bf16x2(dv[0], dv[1]);
          }
        }
        lap(2);
        if (DKV) tmem_st16(trow + COL_X + half * 32, pp);
        tmem_st16(trow + COL_Y + half * 32, pd);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_pd);  // 8 arrivals, not 256 serialised updates of one shared-memory word
        xph ^= 1;
        lap(3);
      }
      // ---- epilogue: accumulators -> bf16 rows of dqkv
      mbar_wait(bar_acc, aph);
      tc_fence_after();
      uint32_t a0[32], b0[32];
      tmem_ld32(trow + COL_A2 + half * 32, a0);
      if (DKV) tmem_ld32(trow + COL_A1 + half * 32, b0);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      {
        // four neighbouring lanes write one row's 64 bytes (8 rows per store instruction, not 32 rows of 16 bytes)
        const int row4 = ob * 128 + quad * 32 + (lane & ~3);  // first of this lane group's four rows
        __nv_bfloat16* base = p.dqkv + (long long)(b * n_tok + row4) * p.lddqkv + h * 64 + half * 32 + (lane & 3) * 8;
        auto store32 = [&](__nv_bfloat16* dst, const uint32_t* v, float mul) {
          uint4 c[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            c[j].x = pack_bf16x2(__uint_as_float(v[8 * j + 0]) * mul, __uint_as_float(v[8 * j + 1]) * mul);
            c[j].y = pack_bf16x2(__uint_as_float(v[8 * j + 2]) * mul, __uint_as_float(v[8 * j + 3]) * mul);
            c[j].z = pack_bf16x2(__uint_as_float(v[8 * j + 4]) * mul, __uint_as_float(v[8 * j + 5]) * mul);
            c[j].w = pack_bf16x2(__uint_as_float(v[8 * j + 6]) * mul, __uint_as_float(v[8 * j + 7]) * mul);
          }
          lane4_transpose_u4(c, lane);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (row4 + j < n_tok) *reinterpret_cast<uint4*>(dst + (long long)j * p.lddqkv) = c[j];
        };
        if (DKV) {
          store32(base + p.dim, a0, p.scale);  // dK = scale * (P .* (dP - D))^T Q   (acc2)
          store32(base + 2 * p.dim, b0, 1.f);  // dV = P^T dO                        (acc1)
        } else {
          store32(base, a0, p.scale);          // dQ = scale * (P .* (dP - D)) K     (acc2)
        }
      }
      aph ^= 1;
      lap(4);
    }
    if (kAtbProf && p.prof && warp == 2 && lane == 0) {
      long long* q = p.prof + 16ll * (blockIdx.x + (DKV ? 0 : gridDim.x));  // dK/dV instance first, dQ instance behind
      for (int i = 0; i < 5; ++i) q[i] = pa[i];
      q[5] = clock64() - pa_t0;
      q[6] = (t1 - t0) * ncb;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ATB_TMEM);
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// Version 2 of the same kernel pair: ONE CTA per SM owning all 512 TMEM columns, X / Y TRIPLE-buffered, two MMA threads.
// In version 1 a tile pair is a serial chain inside its CTA (X, Y MMAs -> math -> accumulate MMAs -> next X, Y MMAs; P / dS
// alias X / Y, so the next X, Y cannot start before the accumulations have consumed them) and only the SM's second CTA fills
// the gaps: tools/attn_bwd_roles.py shows a math warp waiting for X, Y 22 % (dK/dV) / 47 % (dQ) of its life and the MMA thread
// waiting for P / dS 50-58 % of its.  Here tile n+3's X, Y go into the buffer tile n's accumulations free, so the X, Y of the
// next tiles are complete long before the math of tile n ends; the row operands (128 keys or queries) have two smem stages
// and the column tiles six.  One issuing thread needs ~100 cycles per MMA (waits, commits, descriptors): with 16 MMAs per tile
// pair a single issuer was the bound (first cut of this version: 241 us against 215 us for version 1 at B = 32), so the X, Y
// products and the accumulations have a thread each; "buffer free" travels between them as a commit barrier.
//   TMEM: X0 0 | Y0 64 | X1 128 | Y1 192 | X2 256 | Y2 320 | acc1 384 | acc2 448   (tile n uses buffer n % 3; P / dS alias X / Y)
// With the MMAs off the critical path the eight math warps were the bound (two per scheduler, latency-bound: 46-53 % of a
// tile pair in "ld + math"), so the math runs on SIXTEEN warps, four threads per row with 16 of the 64 tile columns each.
// An item's accumulators are written out AFTER the math of the next item's first tile: by then the last accumulation has
// completed, so nobody waits for it (the epilogue was 20-27 % of a math warp's life when it followed the last tile directly).
constexpr int ATB2_THREADS = 608;  // warp 0 TMA, warp 1 X/Y MMAs, warps 2..17 math (four threads per row), warp 18 accumulation MMAs
constexpr int ATB2_ACC_WARP = 18;
constexpr int ATB2_CSTAGES = 6, ATB2_RSTAGES = 2, ATB2_NB = 3;

template <bool DKV>
__global__ void __launch_bounds__(ATB2_THREADS, 1) attn_bwd2_kernel(const __grid_constant__ CUtensorMap tmap_qkv,
                                                                    const __grid_constant__ CUtensorMap tmap_do,
                                                                    const __grid_constant__ CUtensorMap tmap_qkv64,
                                                                    const __grid_constant__ CUtensorMap tmap_do64,
                                                                    const AttnBwdDev p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // smem: row operands R1 | R2 x 2 stages, column tiles (C1 | C2) x 4 stages, barriers, per-warp statistic rows
  auto sR = [&](int st) { return smem_base + st * 2 * ATB_RTILE; };
  auto sC = [&](int st) { return smem_base + ATB2_RSTAGES * 2 * ATB_RTILE + st * 2 * ATB_CTILE; };
  const uint32_t misc_off = ATB2_RSTAGES * 2 * ATB_RTILE + ATB2_CSTAGES * 2 * ATB_CTILE;
  const uint32_t bar_base = smem_base + misc_off;
  auto r_full = [&](int s_) { return bar_base + 8u * s_; };
  auto r_empty = [&](int s_) { return bar_base + 8u * (2 + s_); };
  auto c_full = [&](int s_) { return bar_base + 8u * (4 + s_); };
  auto c_empty = [&](int s_) { return bar_base + 8u * (10 + s_); };
  auto bar_xy = [&](int b_) { return bar_base + 8u * (16 + b_); };     // X, Y of the tile in buffer b_ complete
  auto bar_pd = [&](int b_) { return bar_base + 8u * (19 + b_); };     // P, dS written back
  auto buf_free = [&](int b_) { return bar_base + 8u * (22 + b_); };   // accumulations have consumed P, dS
  const uint32_t bar_acc = bar_base + 8u * 25, acc_empty = bar_base + 8u * 26;
  const uint32_t tmem_slot = bar_base + 8u * 27;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + misc_off + 8 * 27);
  float* s_stat = reinterpret_cast<float*>(smem_gen + misc_off + 256);  // [16 warps][lse * log2e 16 | D 16]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = (int)((long long)p.total_items * blockIdx.x / gridDim.x);
  const int t1 = (int)((long long)p.total_items * (blockIdx.x + 1) / gridDim.x);
  const int nrb = p.rblocks, ncb = p.cblocks;
  const int n_tiles = (t1 - t0) * ncb;
  constexpr uint32_t COL_A1 = 384, COL_A2 = 448;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_qkv);
    tma_prefetch_desc(&tmap_do);
    tma_prefetch_desc(&tmap_qkv64);
    tma_prefetch_desc(&tmap_do64);
    for (int i = 0; i < 16; ++i) mbar_init(bar_base + 8u * i, 1);
    for (int b_ = 0; b_ < ATB2_NB; ++b_) {
      mbar_init(bar_xy(b_), 1);
      mbar_init(bar_pd(b_), 16);  // one arrival per math warp
      mbar_init(buf_free(b_), 1);
    }
    mbar_init(bar_acc, 1);
    mbar_init(acc_empty, 16);
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
  griddep_sync();  // PDL: prologue overlapped the previous kernel; its results are visible from here on

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int cs = 0, n = 0;
      uint32_t cph = 0;
      for (int t = t0, itx = 0; t < t1; ++t, ++itx) {
        const int bh = t / nrb, ob = t - bh * nrb;  // outer block: 128 keys (DKV) or 128 queries
        const int b = bh / p.heads, h = bh - b * p.heads;
        const int row0 = b * p.n_tok;
        const int rs = itx & 1;
        mbar_wait(r_empty(rs), ((itx >> 1) & 1) ^ 1);
        mbar_expect_tx(r_full(rs), 2 * ATB_RTILE);
        if (DKV) {
          tma_load_2d(sR(rs), &tmap_qkv, r_full(rs), p.dim + h * 64, row0 + ob * 128);                  // K_j
          tma_load_2d(sR(rs) + ATB_RTILE, &tmap_qkv, r_full(rs), 2 * p.dim + h * 64, row0 + ob * 128);  // V_j
        } else {
          tma_load_2d(sR(rs), &tmap_qkv, r_full(rs), h * 64, row0 + ob * 128);                          // Q_i
          tma_load_2d(sR(rs) + ATB_RTILE, &tmap_do, r_full(rs), h * 64, row0 + ob * 128);               // dO_i
        }
        for (int ib = 0; ib < ncb; ++ib, ++n) {
          mbar_wait(c_empty(cs), cph ^ 1);
          mbar_expect_tx(c_full(cs), 2 * ATB_CTILE);
          const uint32_t c1 = sC(cs), c2 = c1 + ATB_CTILE;
          if (DKV) {
            tma_load_2d(c1, &tmap_qkv64, c_full(cs), h * 64, row0 + ib * 64);            // Q (64 queries)
            tma_load_2d(c2, &tmap_do64, c_full(cs), h * 64, row0 + ib * 64);             // dO
          } else {
            tma_load_2d(c1, &tmap_qkv64, c_full(cs), p.dim + h * 64, row0 + ib * 64);      // K (64 keys)
            tma_load_2d(c2, &tmap_qkv64, c_full(cs), 2 * p.dim + h * 64, row0 + ib * 64);  // V
          }
          if (++cs == ATB2_CSTAGES) { cs = 0; cph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer 1: X = R1 C1^T, Y = R2 C2^T of tile m into buffer m % 3 =====================
      const uint32_t idesc_xy = umma_idesc_bf16(128, 64);
      long long pm[2] = {0, 0};  // waiting (operands, buffer) | issuing
      const long long pm_t0 = kAtbProf ? clock64() : 0;
      long long pt = pm_t0;
      auto lap = [&](int i) { if (kAtbProf) { const long long c_ = clock64(); pm[i] += c_ - pt; pt = c_; } };
      int x_item = 0, x_ib = 0, bi = 0, cs = 0;
      uint32_t bph = 0, cph = 0;  // phase of this buffer's / stage's current use
      for (int m = 0; m < n_tiles; ++m) {
        const int rs = x_item & 1;
        if (m >= ATB2_NB) mbar_wait(buf_free(bi), bph ^ 1);  // accumulations of tile m - 3 have consumed this buffer's P, dS
        if (x_ib == 0) mbar_wait(r_full(rs), (x_item >> 1) & 1);
        mbar_wait(c_full(cs), cph);
        lap(0);
        tc_fence_after();
        const uint64_t dr1 = umma_desc_sw128(sR(rs)), dr2 = umma_desc_sw128(sR(rs) + ATB_RTILE);
        const uint64_t dc1 = umma_desc_sw128(sC(cs)), dc2 = umma_desc_sw128(sC(cs) + ATB_CTILE);
        const uint32_t tx = tmem_base + bi * 128, ty = tx + 64;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tx, dr1 + 2 * k, dc1 + 2 * k, idesc_xy, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(ty, dr2 + 2 * k, dc2 + 2 * k, idesc_xy, k != 0);
        umma_commit(bar_xy(bi));
        if (x_ib == ncb - 1) umma_commit(r_empty(rs));  // row operands of this item no longer needed
        if (++x_ib == ncb) { x_ib = 0; ++x_item; }
        if (++bi == ATB2_NB) { bi = 0; bph ^= 1; }
        if (++cs == ATB2_CSTAGES) { cs = 0; cph ^= 1; }
        lap(1);
      }
      if (kAtbProf && p.prof) {
        long long* q = p.prof + 16ll * (blockIdx.x + (DKV ? 0 : gridDim.x));
        q[8] = pm[0]; q[9] = pm[1]; q[12] = clock64() - pm_t0;
      }
    }
  } else if (warp == ATB2_ACC_WARP) {
    if (lane == 0) {
      // ===================== MMA issuer 2: acc1 += P C2 (dK/dV kernel), acc2 += dS C1 =====================
      const uint32_t idesc_acc = umma_idesc_bf16(128, 64, 0, 1);
      long long pm[2] = {0, 0};  // waiting for P, dS | issuing
      long long pt = kAtbProf ? clock64() : 0;
      auto lap = [&](int i) { if (kAtbProf) { const long long c_ = clock64(); pm[i] += c_ - pt; pt = c_; } };
      int a_ib = 0, bi = 0, cs = 0;
      uint32_t aph = 0, bph = 0;
      for (int m = 0; m < n_tiles; ++m) {
        mbar_wait(bar_pd(bi), bph);
        if (a_ib == 0) mbar_wait(acc_empty, aph ^ 1);  // previous item's accumulators drained
        lap(0);
        tc_fence_after();
        const uint64_t dm1 = umma_desc_sw128(sC(cs), 1024, 1024), dm2 = umma_desc_sw128(sC(cs) + ATB_CTILE, 1024, 1024);  // MN-major
        const uint32_t tx = tmem_base + bi * 128, ty = tx + 64;
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // K = 64 columns of this inner tile; 16 reduction rows = 2 KB per step
          // P / dS of column quarter k (one math thread's 16 columns) sit packed in the first 8 TMEM columns of that quarter
          const uint32_t acol = k * 16;
          if (DKV) umma_bf16_ts(tmem_base + COL_A1, tx + acol, dm2 + (2048 >> 4) * k, idesc_acc, (a_ib | k) != 0);
          umma_bf16_ts(tmem_base + COL_A2, ty + acol, dm1 + (2048 >> 4) * k, idesc_acc, (a_ib | k) != 0);
        }
        umma_commit(c_empty(cs));
        umma_commit(buf_free(bi));
        if (a_ib == ncb - 1) umma_commit(bar_acc);
        if (++a_ib == ncb) { a_ib = 0; aph ^= 1; }
        if (++bi == ATB2_NB) { bi = 0; bph ^= 1; }
        if (++cs == ATB2_CSTAGES) cs = 0;
        lap(1);
      }
      if (kAtbProf && p.prof) {
        long long* q = p.prof + 16ll * (blockIdx.x + (DKV ? 0 : gridDim.x));
        q[10] = pm[0]; q[11] = pm[1];
      }
    }
  } else if (warp < ATB2_ACC_WARP) {
    // ===================== softmax-backward math (16 warps, four threads per row) + epilogue (the first 8) =====================
    const int quad = warp & 3;
    const int cq = (warp - 2) >> 2;     // which 16 of the 64 tile / accumulator columns this thread owns
    const int r = quad * 32 + lane;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int n_tok = p.n_tok;
    uint32_t aph = 0, bph = 0;
    int bi = 0;  // TMEM buffer of the current tile (tile counter % 3) and the phase of its current use
    long long pa[6] = {0, 0, 0, 0, 0, 0};  // statistics staging | wait X, Y | ld + math | st + arrive | item epilogue
    const long long pa_t0 = kAtbProf ? clock64() : 0;
    long long pt = pa_t0;
    auto lap = [&](int i) { if (kAtbProf) { const long long c_ = clock64(); pa[i] += c_ - pt; pt = c_; } };
    float* my_stat = s_stat + (warp - 2) * 32;
    // ---- an item's accumulators -> bf16 rows of dqkv (this thread: its row's 16 columns; two neighbouring lanes store one
    // row's 32 bytes)
    auto write_item = [&](int b, int h, int ob) {
      mbar_wait(bar_acc, aph);
      tc_fence_after();
      uint32_t a0[16], b0[16];
      tmem_ld16(trow + COL_A2 + cq * 16, a0);
      if (DKV) tmem_ld16(trow + COL_A1 + cq * 16, b0);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      const int row2 = ob * 128 + quad * 32 + (lane & ~1);  // first of this lane pair's two rows
      __nv_bfloat16* base = p.dqkv + (long long)(b * n_tok + row2) * p.lddqkv + h * 64 + cq * 16 + (lane & 1) * 8;
      auto store16 = [&](__nv_bfloat16* dst, const uint32_t* v_, float mul) {
        uint4 c[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          c[j].x = pack_bf16x2(__uint_as_float(v_[8 * j + 0]) * mul, __uint_as_float(v_[8 * j + 1]) * mul);
          c[j].y = pack_bf16x2(__uint_as_float(v_[8 * j + 2]) * mul, __uint_as_float(v_[8 * j + 3]) * mul);
          c[j].z = pack_bf16x2(__uint_as_float(v_[8 * j + 4]) * mul, __uint_as_float(v_[8 * j + 5]) * mul);
          c[j].w = pack_bf16x2(__uint_as_float(v_[8 * j + 6]) * mul, __uint_as_float(v_[8 * j + 7]) * mul);
        }
        lane2_transpose_u4(c, lane);
#pragma unroll
        for (int j = 0; j < 2; ++j)
          if (row2 + j < n_tok) *reinterpret_cast<uint4*>(dst + (long long)j * p.lddqkv) = c[j];
      };
      if (DKV) {
        store16(base + p.dim, a0, p.scale);  // dK = scale * (P .* (dP - D))^T Q   (acc2)
        store16(base + 2 * p.dim, b0, 1.f);  // dV = P^T dO                        (acc1)
      } else {
        store16(base, a0, p.scale);          // dQ = scale * (P .* (dP - D)) K     (acc2)
      }
      aph ^= 1;
      lap(4);
    };
    bool pending = false;
    int pend_b = 0, pend_h = 0, pend_ob = 0;
    for (int t = t0; t < t1; ++t) {
      const int bh = t / nrb, ob = t - bh * nrb;
      const int b = bh / p.heads, h = bh - b * p.heads;
      const long long vec0 = (long long)bh * n_tok;
      const int orow = ob * 128 + r;  // key (DKV) or query index of this thread's row
      float lse_r = 0.f, dsum_r = 0.f;
      if (!DKV && orow < n_tok) {
        lse_r = p.lse[vec0 + orow] * 1.4426950408889634f;
        dsum_r = p.dsum[vec0 + orow];
      }
      // DKV: the statistics of this warp's 16 query columns of an inner tile (lanes 0..15 lse, lanes 16..31 D), fetched ONE
      // TILE AHEAD as a raw load — nothing consumes the value before the next tile, so the warp never stalls on it
      auto load_stat = [&](int ib) {
        const int q = min(ib * 64 + cq * 16 + (lane & 15), n_tok - 1);
        return __ldg((lane < 16 ? p.lse : p.dsum) + vec0 + q);
      };
      float stat_next = 0.f;
      if (DKV) stat_next = load_stat(0);
      for (int ib = 0; ib < ncb; ++ib) {
        const float* sl = my_stat;
        const float* sd = my_stat + 16;
        if (DKV) {
          __syncwarp();
          my_stat[lane] = lane < 16 ? stat_next * 1.4426950408889634f : stat_next;
          __syncwarp();
          if (ib + 1 < ncb) stat_next = load_stat(ib + 1);
        }
        lap(0);
        mbar_wait(bar_xy(bi), bph);
        lap(1);
        tc_fence_after();
        const uint32_t tx = trow + bi * 128 + cq * 16, ty = tx + 64;
        uint32_t x[16], y[16], pp[8], pd[8];
        tmem_ld16(tx, x);
        tmem_ld16(ty, y);
        tmem_ld_wait();
        // dS is formed WITHOUT the softmax scale (applied once per item to the accumulator in the epilogue), and tiles
        // that lie completely inside the sequence skip the per-element bounds predicates
        const bool interior = (ob * 128 + 127 < n_tok) && (ib * 64 + 63 < n_tok);  // CTA-uniform
        if (interior) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float pv[2], dv[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float l2 = DKV ? sl[2 * j + e] : lse_r;
              const float dd = DKV ? sd[2 * j + e] : dsum_r;
              const float pr = ex2a(__uint_as_float(x[2 * j + e]) * p.scale_log2e - l2);
              pv[e] = pr;
              dv[e] = pr * (__uint_as_float(y[2 * j + e]) - dd);
            }
            pp[j] = pack_bf16x2(pv[0], pv[1]);
            pd[j] = pack_bf16x2(dv[0], dv[1]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float pv[2], dv[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int icol = ib * 64 + cq * 16 + 2 * j + e;
              const bool ok = orow < n_tok && icol < n_tok;
              const float l2 = DKV ? sl[2 * j + e] : lse_r;
              const float dd = DKV ? sd[2 * j + e] : dsum_r;
              const float pr = ok ? ex2a(__uint_as_float(x[2 * j + e]) * p.scale_log2e - l2) : 0.f;
              pv[e] = pr;
              dv[e] = pr * (__uint_as_float(y[2 * j + e]) - dd);
            }
            pp[j] = pack_bf16x2(pv[0], pv[1]);
            pd[j] = pack_bf16x2(dv[0], dv[1]);
          }
        }
        lap(2);
        if (DKV) tmem_st8(tx, pp);
        tmem_st8(ty, pd);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_pd(bi));
        if (++bi == ATB2_NB) { bi = 0; bph ^= 1; }
        lap(3);
        if (ib == 0 && pending) {  // the previous item's accumulators: complete by now, and the accumulation thread is
          write_item(pend_b, pend_h, pend_ob);  // holding this item's first accumulation until they have been read
          pending = false;
        }
      }
      pend_b = b; pend_h = h; pend_ob = ob; pending = true;
    }
    if (pending) write_item(pend_b, pend_h, pend_ob);
    if (kAtbProf && p.prof && warp == 2 && lane == 0) {
      long long* q = p.prof + 16ll * (blockIdx.x + (DKV ? 0 : gridDim.x));
      for (int i = 0; i < 5; ++i) q[i] = pa[i];
      q[5] = clock64() - pa_t0;
      q[6] = n_tiles;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// D[b, h, q] = sum_d dO[q, h*64 + d] * O[q, h*64 + d]; one warp per (token row, 2 heads per pass)
__global__ void attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ dout, long long lddo,
                                     const __nv_bfloat16* __restrict__ out, long long ldo, float* __restrict__ dsum,
                                     int batch, int n_tok, int heads) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= (long long)batch * n_tok) return;
  const int b = (int)(row / n_tok), q = (int)(row - (long long)b * n_tok);
  // each lane covers 8 consecutive channels; 32 lanes = 256 channels = 4 heads per iteration
  for (int c0 = 0; c0 < heads * 64; c0 += 256) {
    const int c = c0 + lane * 8;
    float s = 0.f;
    if (c < heads * 64) {
      const uint4 a = *reinterpret_cast<const uint4*>(dout + row * lddo + c);
      const uint4 o = *reinterpret_cast<const uint4*>(out + row * ldo + c);
      const uint32_t* pa = &a.x;
      const uint32_t* po = &o.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 fa = unpack_bf16x2(pa[j]), fo = unpack_bf16x2(po[j]);
        s += fa.x * fo.x + fa.y * fo.y;
      }
    }
    // reduce over the 8 lanes of one head
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    const int h = c / 64;
    if ((lane & 7) == 0 && h < heads) dsum[((long long)b * heads + h) * n_tok + q] = s;
  }
}

}  // namespace mv

// qkv [B*N, 3D] (saved forward input), out [B*N, D] (saved forward output), dout [B*N, D], lse [B, heads, N]
// -> dqkv [B*N, 3D] (dq | dk | dv); dsum: fp32 workspace [B, heads, N].
extern "C" int mv_attn_bwd(const void* qkv, int64_t ldqkv, const void* out, int64_t ldo, const void* dout, int64_t lddo,
                           const float* lse, float* dsum, void* dqkv, int64_t lddqkv, int batch, int n_tok, int heads,
                           float scale, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(qkv && out && dout && lse && dsum && dqkv && batch > 0 && heads > 0 && n_tok > 0, "mv_attn_bwd: null/empty");
  MV_CHECK_ARG(ldo % 8 == 0 && lddo % 8 == 0 && lddqkv % 8 == 0, "mv_attn_bwd: leading dimensions must be multiples of 8");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const long long rows = (long long)batch * n_tok;
  MV_LAUNCH(attn_bwd_prep_kernel, (unsigned)((rows + 7) / 8), 256, 0, stream, 
      reinterpret_cast<const __nv_bfloat16*>(dout), lddo, reinterpret_cast<const __nv_bfloat16*>(out), ldo, dsum, batch,
      n_tok, heads);
  MV_CHECK_LAUNCH("attn_bwd_prep");
  AttnBwdDev p;
  p.n_tok = n_tok;
  p.heads = heads;
  p.dim = heads * 64;
  p.rblocks = (n_tok + 127) / 128;
  p.cblocks = (n_tok + 63) / 64;
  p.total_items = batch * heads * p.rblocks;
  p.scale = scale;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.lse = lse;
  p.dsum = dsum;
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  p.lddqkv = lddqkv;
  p.prof = g_attn_prof;
  const CUtensorMap* tq = get_tmap_2d_bf16(qkv, rows, 3ull * p.dim, ldqkv, 128);
  const CUtensorMap* td = get_tmap_2d_bf16(dout, rows, (uint64_t)p.dim, lddo, 128);
  const CUtensorMap* tq64 = get_tmap_2d_bf16(qkv, rows, 3ull * p.dim, ldqkv, 64);
  const CUtensorMap* td64 = get_tmap_2d_bf16(dout, rows, (uint64_t)p.dim, lddo, 64);
  if (!tq || !td || !tq64 || !td64) return MV_ERR_ARG;
  static const int ver_env = [] { const char* e = getenv("MV_ATTN_BWD_V"); return e ? atoi(e) : 2; }();  // 1: two CTAs per SM
  if (ver_env != 1) {
    const int smem2 = ATB2_RSTAGES * 2 * ATB_RTILE + ATB2_CSTAGES * 2 * ATB_CTILE + 256 + 16 * 32 * 4 + 1024;  // 164 KB
    static std::atomic<uint64_t> attr2{0};  // one bit per device
    if (first_use_on_device(attr2)) {
      cudaError_t e = cudaFuncSetAttribute(attn_bwd2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
      if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute(attn_bwd2): %s", cudaGetErrorString(e));
        return (int)e;
      }
    }
    int grid2 = device_sms() > 0 ? device_sms() : 148;
    if (grid2 > p.total_items) grid2 = p.total_items;
    MV_LAUNCH((attn_bwd2_kernel<true>), grid2, ATB2_THREADS, smem2, stream, *tq, *td, *tq64, *td64, p);
    MV_CHECK_LAUNCH("attn_bwd_dkv");
    MV_LAUNCH((attn_bwd2_kernel<false>), grid2, ATB2_THREADS, smem2, stream, *tq, *td, *tq64, *td64, p);
    MV_CHECK_LAUNCH("attn_bwd_dq");
    return MV_OK;
  }
  const int smem = 2 * ATB_RTILE + 4 * ATB_CTILE + 128 + 8 * 128 * 4 + 1024;
  static std::atomic<uint64_t> attr{0};  // one bit per device
  if (first_use_on_device(attr)) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attn_bwd): %s", cudaGetErrorString(e));
      return (int)e;
    }
  }
  int grid = 2 * (device_sms() > 0 ? device_sms() : 148);  // two co-resident CTAs per SM
  if (grid > p.total_items) grid = p.total_items;
  MV_LAUNCH((attn_bwd_kernel<true>), grid, ATB_THREADS, smem, stream, *tq, *td, *tq64, *td64, p);
  MV_CHECK_LAUNCH("attn_bwd_dkv");
  MV_LAUNCH((attn_bwd_kernel<false>), grid, ATB_THREADS, smem, stream, *tq, *td, *tq64, *td64, p);
  MV_CHECK_LAUNCH("attn_bwd_dq");
  return MV_OK;
}
