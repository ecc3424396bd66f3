// Multi-head self-attention forward (head_dim 64, any sequence length >= 16) on tcgen05 / TMEM.
//
// Replaces timm Attention.forward's F.scaled_dot_product_attention(q, k, v) (no mask, scale 1/8, dropout 0; the ViT is
// created at src/generators/foundation_models.py:53-57) including the reshape(B,N,3,H,64).permute(2,0,3,1,4) split
// and the transpose(1,2).reshape(B,N,C) merge: it reads q/k/v straight from the fused qkv rows [B*N, 3*D] and writes
// token-major O [B*N, D] (+ the log-sum-exp per row for the backward pass).
//
// Persistent kernel, one CTA per SM; work item = (image, head, 128-query tile). Keys/values stream through shared memory
// in blocks of kb <= 128 keys with an online softmax, so any N works (329 tokens at 256 px, 1301 at 512 px).
// A CTA runs TWO items at a time in two "slots", each with its own Q buffers, K/V ring, TMEM region and softmax
// warpgroup; one TMA thread and one MMA thread serve both slots in a fixed interleave
//        S_A(j+1) | PV_B(j) | S_B(j+1) | PV_A(j+1) | ...
// so the tensor core works for one slot while the other slot's warpgroup does its exponentials:
//   warp 0 (1 thread) : TMA producer — Q (2 buffers / slot) and K|V blocks (2 stages / slot), 128-byte swizzle
//   warp 1 (1 thread) : tcgen05.mma issuer — S = Q K_j^T (SS), O += P V_j (A = P from TMEM, V MN-major from smem)
//   warps 2..9 / 10..17: softmax warps of slot A / B, TWO threads per query row (same TMEM lane quadrant, the key block's
//                       columns split between them; row max / row sum exchanged through shared memory): running max with LAZY rescale of the
//                       O accumulator (only when the max grows by > 2^8), p = exp2(s*c - m), bf16 P written over S in
//                       TMEM, final O / l -> global.
// TMEM per slot: S/P at +0 (kb fp32 columns, P aliases the first kb/2), O at +128 (64 columns); slots at 0 and 256.
#include "mv_host.h"
#include "mv_ptx.cuh"

namespace mv {

constexpr int ATT_THREADS = 576;  // TMA warp, MMA warp, 2 slots x 8 softmax warps (two threads per query row)
constexpr int ATT_QBYTES = 128 * 128;       // one [128 x 64] bf16 tile
constexpr float ATT_RESCALE_THRESHOLD = 8.f;  // log2 units

struct AttnDev {
  int n_tok;     // tokens per image
  int kb;        // keys per block (multiple of 16, <= 128)
  int nb;        // key blocks per item
  int heads, dim;
  int q_tiles;
  int total_tiles;
  float scale_log2e;
  __nv_bfloat16* out;
  long long ldo;
  float* lse;  // [B, heads, n_tok] or null (natural log)
  long long* prof;  // diagnostics (MV_GEMM_PROFILE builds only): per CTA 16 x int64 cycle sums, see mv_attn_set_profile_buffer
};

#ifndef MV_GEMM_PROFILE
#define MV_GEMM_PROFILE 0
#endif
constexpr bool kAttProf = MV_GEMM_PROFILE != 0;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(ATT_THREADS, 1) attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q,
                                                                  const __grid_constant__ CUtensorMap tmap_kv,
                                                                  const AttnDev p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t kv_bytes = p.kb * 128;  // one K (or V) block
  // smem: per slot: Q[2] | KV ring [2] x (K | V, 16 KB reserved each)
  auto sQ = [&](int slot, int buf) { return smem_base + (slot * 2 + buf) * ATT_QBYTES; };
  auto sK = [&](int slot, int st) { return smem_base + 4 * ATT_QBYTES + ((slot * 2 + st) * 2) * ATT_QBYTES; };
  auto sV = [&](int slot, int st) { return sK(slot, st) + ATT_QBYTES; };
  const uint32_t misc_off = 12 * ATT_QBYTES;
  const uint32_t bar_base = smem_base + misc_off;
  // barriers per slot (11): q_full[2], q_empty[2], kv_full[2], kv_empty[2], s_full, p_full, o_full
  auto bar = [&](int slot, int idx) { return bar_base + 8u * (slot * 11 + idx); };
  const uint32_t tmem_slot = bar_base + 8u * 22;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + misc_off + 8 * 22);
  // row max (double buffered by key-block parity) and row sum exchanged between the two threads of a row:
  // xchg[slot][buffer 0,1 = max, 2 = sum][column half][128 rows]
  float* xchg = reinterpret_cast<float*>(smem_gen + misc_off + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = (int)((long long)p.total_tiles * blockIdx.x / gridDim.x);
  const int t1 = (int)((long long)p.total_tiles * (blockIdx.x + 1) / gridDim.x);
  const int cnt = t1 - t0;
  const int nb = p.nb;
  // slot s handles items t0 + s, t0 + s + 2, ...
  const int n_items[2] = {(cnt + 1) / 2, cnt / 2};

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
    for (int s = 0; s < 2; ++s) {
      for (int i = 0; i < 8; ++i) mbar_init(bar(s, i), 1);
      mbar_init(bar(s, 8), 1);    // s_full
      mbar_init(bar(s, 9), 256);  // p_full
      mbar_init(bar(s, 10), 1);   // o_full
    }
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
      const int G[2] = {n_items[0] * nb, n_items[1] * nb};
      const int gmax = G[0] > G[1] ? G[0] : G[1];
      for (int g = 0; g < gmax; ++g) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (g >= G[s]) continue;
          const int it = g / nb, j = g - it * nb;
          const int t = t0 + s + 2 * it;
          const int bh = t / p.q_tiles, qt = t - bh * p.q_tiles;
          const int b = bh / p.heads, h = bh - b * p.heads;
          const int row0 = b * p.n_tok;
          if (j == 0) {
            const int qb = it & 1;
            mbar_wait(bar(s, 2 + qb), ((it >> 1) & 1) ^ 1);
            mbar_expect_tx(bar(s, qb), ATT_QBYTES);
            tma_load_2d(sQ(s, qb), &tmap_q, bar(s, qb), h * 64, row0 + qt * 128);
          }
          const int st = g & 1;
          mbar_wait(bar(s, 6 + st), ((g >> 1) & 1) ^ 1);
          mbar_expect_tx(bar(s, 4 + st), 2 * kv_bytes);
          tma_load_2d(sK(s, st), &tmap_kv, bar(s, 4 + st), p.dim + h * 64, row0 + j * p.kb);
          tma_load_2d(sV(s, st), &tmap_kv, bar(s, 4 + st), 2 * p.dim + h * 64, row0 + j * p.kb);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      const int G[2] = {n_items[0] * nb, n_items[1] * nb};
      const int gmax = G[0] > G[1] ? G[0] : G[1];
      const uint32_t idesc_s = umma_idesc_bf16(128, p.kb);
      const uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);
      const int ksteps = p.kb / 16;
      const int csplit_m = (p.kb / 32 + 1) >> 1;
      long long pm_kv = 0, pm_p = 0;
      const long long pm_t0 = kAttProf ? clock64() : 0;
      auto issue_s = [&](int s, int g) {
        const int it = g / nb, j = g - it * nb;
        const int qb = it & 1, st = g & 1;
        const long long t_ = kAttProf ? clock64() : 0;
        if (j == 0) mbar_wait(bar(s, qb), (it >> 1) & 1);
        mbar_wait(bar(s, 4 + st), (g >> 1) & 1);
        if (kAttProf) pm_kv += clock64() - t_;
        tc_fence_after();
        const uint64_t dq = umma_desc_sw128(sQ(s, qb));
        const uint64_t dk = umma_desc_sw128(sK(s, st));
        const uint32_t d = tmem_base + s * 256;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(d, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        umma_commit(bar(s, 8));
        if (j == nb - 1) umma_commit(bar(s, 2 + qb));  // Q buffer free
      };
      auto issue_pv = [&](int s, int g) {
        const int it = g / nb, j = g - it * nb;
        const int st = g & 1;
        const long long t_ = kAttProf ? clock64() : 0;
        mbar_wait(bar(s, 9), g & 1);
        if (kAttProf) pm_p += clock64() - t_;
        tc_fence_after();
        const uint32_t d = tmem_base + s * 256 + 128;
        const uint32_t a = tmem_base + s * 256;
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t dv = umma_desc_sw128(sV(s, st) + k * 2048, 1024, 1024);
          // P (bf16 pairs, 8 TMEM columns per 16 keys): keys of the first csplit 32-key chunks start at column 0, the
          // rest (written by the row's second softmax thread) at column 32 csplit
          const uint32_t pcol = k < 2 * csplit_m ? k * 8 : 32 * csplit_m + (k - 2 * csplit_m) * 8;
          umma_bf16_ts(d, a + pcol, dv, idesc_pv, (j | k) != 0);
        }
        umma_commit(bar(s, 6 + st));  // K|V stage free
        if (j == nb - 1) umma_commit(bar(s, 10));
      };
      if (G[0] > 0) issue_s(0, 0);
      if (G[1] > 0) issue_s(1, 0);
      for (int g = 0; g < gmax; ++g) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (g < G[s]) {
            issue_pv(s, g);
            if (g + 1 < G[s]) issue_s(s, g + 1);
          }
        }
      }
      if (kAttProf && p.prof) {
        long long* q = p.prof + 16ll * blockIdx.x;
        q[8] = pm_kv; q[9] = pm_p; q[10] = clock64() - pm_t0;
      }
    }
  } else {
    // ===================== softmax warps =====================
    const int s = (warp - 2) >> 3;         // slot
    const int quad = warp & 3;             // TMEM lane quadrant this warp may access
    const int half = ((warp - 2) & 7) >> 2;  // which part of the key block's columns / of the 64 output dims
    const int r = quad * 32 + lane;        // query row within the tile == TMEM lane
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + s * 256;
    const int n_tok = p.n_tok, kb = p.kb;
    const int full32 = kb / 32;
    const bool tail16 = (kb & 31) != 0;
    // 32-column chunks [cbeg, cend) of the key block belong to this thread; the 16-column tail goes to half 1
    const int csplit = (full32 + 1) >> 1;
    const int cbeg = half ? csplit : 0, cend = half ? full32 : csplit;
    const bool my_tail = tail16 && half == 1;
    const float c = p.scale_log2e;
    float* xs = xchg + s * (3 * 2 * 128);
    auto slot_bar = [&]() { asm volatile("bar.sync %0, 256;" ::"r"(1 + s) : "memory"); };
    int g = 0;
    long long pa[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // wait S | pass 1 | exchange barrier | rescale | pass 2 | st wait + arrive | item epilogue
    const long long pa_t0 = kAttProf ? clock64() : 0;
    long long pt = pa_t0;
    auto lap = [&](int i) { if (kAttProf) { const long long n = clock64(); pa[i] += n - pt; pt = n; } };
    for (int it = 0; it < n_items[s]; ++it) {
      const int t = t0 + s + 2 * it;
      const int bh = t / p.q_tiles, qt = t - bh * p.q_tiles;
      const int b = bh / p.heads, h = bh - b * p.heads;
      float m_ref = -INFINITY, l = 0.f;  // l: this thread's share of the row sum
      for (int j = 0; j < nb; ++j, ++g) {
        mbar_wait(bar(s, 8), g & 1);
        lap(0);
        tc_fence_after();
        const int key0 = j * kb;
        const bool partial = key0 + kb > n_tok;  // block holds keys beyond the sequence: mask them
        // ---- pass 1: row max over this thread's columns
        float mx = -INFINITY;
#pragma unroll 1
        for (int cc = cbeg; cc < cend; ++cc) {
          uint32_t v[32];
          tmem_ld32(trow + cc * 32, v);
          tmem_ld_wait();
          if (!partial) {
#pragma unroll
            for (int q = 0; q < 32; ++q) mx = fmaxf(mx, __uint_as_float(v[q]));
          } else {
#pragma unroll
            for (int q = 0; q < 32; ++q)
              if (key0 + cc * 32 + q < n_tok) mx = fmaxf(mx, __uint_as_float(v[q]));
          }
        }
        if (my_tail) {
          uint32_t v[16];
          tmem_ld16(trow + full32 * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 16; ++q)
            if (key0 + full32 * 32 + q < n_tok) mx = fmaxf(mx, __uint_as_float(v[q]));
        }
        // exchange with the row's other thread (buffer = block parity: the partner may already be one block ahead)
        float* xm = xs + (g & 1) * 256;
        xm[half * 128 + r] = mx;
        lap(1);
        slot_bar();
        lap(2);
        mx = fmaxf(mx, xm[(half ^ 1) * 128 + r]);
        const float m_new = fmaxf(m_ref, mx * c);
        // ---- lazy rescale of the accumulator (warp-uniform decision: tcgen05.ld/st are warp collectives; both warps of
        // a quadrant see the same 32 rows, so they take the same branch); each thread rescales its 32 of the 64 O columns
        const bool grow = (j == 0) || (m_new - m_ref > ATT_RESCALE_THRESHOLD);
        if (__any_sync(0xffffffffu, grow)) {
          const float m_use = grow ? m_new : m_ref;
          if (j > 0) {
            const float alpha = grow ? ex2_approx(m_ref - m_new) : 1.f;
            l *= alpha;
            uint32_t o[32];
            tmem_ld32(trow + 128 + half * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 32; ++q) o[q] = __float_as_uint(__uint_as_float(o[q]) * alpha);
            tmem_st16(trow + 128 + half * 32, reinterpret_cast<uint32_t(&)[16]>(o[0]));
            tmem_st16(trow + 144 + half * 32, reinterpret_cast<uint32_t(&)[16]>(o[16]));
          }
          m_ref = m_use;
        }
        lap(3);
        // ---- pass 2: p = exp2(s*c - m_ref) -> bf16, written over S
        float sum = 0.f;
#pragma unroll 1
        for (int cc = cbeg; cc < cend; ++cc) {
          uint32_t v[32], pk[16];
          tmem_ld32(trow + cc * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            float e0 = ex2_approx(__uint_as_float(v[2 * q]) * c - m_ref);
            float e1 = ex2_approx(__uint_as_float(v[2 * q + 1]) * c - m_ref);
            if (partial) {
              if (key0 + cc * 32 + 2 * q >= n_tok) e0 = 0.f;
              if (key0 + cc * 32 + 2 * q + 1 >= n_tok) e1 = 0.f;
            }
            sum += e0 + e1;
            pk[q] = pack_bf16x2(e0, e1);
          }
          // P of chunk cc goes to the start of THIS thread's own S column range (half 0: 16 cc, half 1: 32 csplit +
          // 16 (cc - csplit)) — always columns this thread has already read, never the partner's unread scores
          tmem_st16(trow + (half ? 32 * csplit + 16 * (cc - csplit) : 16 * cc), pk);
        }
        if (my_tail) {
          uint32_t v[16], pk[16];
          tmem_ld16(trow + full32 * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int key = key0 + full32 * 32 + 2 * q;
            const float e0 = key < n_tok ? ex2_approx(__uint_as_float(v[2 * q]) * c - m_ref) : 0.f;
            const float e1 = key + 1 < n_tok ? ex2_approx(__uint_as_float(v[2 * q + 1]) * c - m_ref) : 0.f;
            sum += e0 + e1;
            pk[q] = pack_bf16x2(e0, e1);
          }
#pragma unroll
          for (int q = 8; q < 16; ++q) pk[q] = 0u;
          tmem_st16(trow + 32 * csplit + 16 * (full32 - csplit), pk);  // upper 8 columns: dead (already read) part of S
        }
        l += sum;
        lap(4);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar(s, 9));
        lap(5);
      }
      // ---- epilogue of the item: O / l -> bf16 rows; row sum = both threads' shares
      float* xl = xs + 2 * 256;
      xl[half * 128 + r] = l;
      slot_bar();
      const float l_row = l + xl[(half ^ 1) * 128 + r];
      mbar_wait(bar(s, 10), it & 1);
      tc_fence_after();
      uint32_t o[32];
      tmem_ld32(trow + 128 + half * 32, o);
      tmem_ld_wait();
      const int q = qt * 128 + r;
      const float inv = 1.f / l_row;
      uint4 oc[4];  // this thread's 32 output columns as four 16-byte chunks
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        oc[jj].x = pack_bf16x2(__uint_as_float(o[8 * jj + 0]) * inv, __uint_as_float(o[8 * jj + 1]) * inv);
        oc[jj].y = pack_bf16x2(__uint_as_float(o[8 * jj + 2]) * inv, __uint_as_float(o[8 * jj + 3]) * inv);
        oc[jj].z = pack_bf16x2(__uint_as_float(o[8 * jj + 4]) * inv, __uint_as_float(o[8 * jj + 5]) * inv);
        oc[jj].w = pack_bf16x2(__uint_as_float(o[8 * jj + 6]) * inv, __uint_as_float(o[8 * jj + 7]) * inv);
      }
      // four neighbouring lanes write one row's 64 bytes (profiles/r02_attention_fwd_experiments.txt: the row-per-lane
      // stores, 32 rows x 16 bytes per instruction, were 18 % of the kernel)
      lane4_transpose_u4(oc, lane);
      {
        const int qw = qt * 128 + quad * 32 + (lane & ~3);  // first of this lane group's four rows
        __nv_bfloat16* obase = p.out + (long long)(b * n_tok + qw) * p.ldo + h * 64 + half * 32 + (lane & 3) * 8;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          if (qw + jj < n_tok) *reinterpret_cast<uint4*>(obase + (long long)jj * p.ldo) = oc[jj];
      }
      if (q < n_tok && p.lse && half == 0)
        p.lse[((long long)b * p.heads + h) * n_tok + q] = (m_ref + log2f(l_row)) * 0.6931471805599453f;
      slot_bar();  // the sum buffer is reused by the next item
      lap(6);
    }
    if (kAttProf && p.prof && warp == 2 && lane == 0) {
      long long* q = p.prof + 16ll * blockIdx.x;
      for (int i = 0; i < 7; ++i) q[i] = pa[i];
      q[7] = clock64() - pa_t0;
      q[11] = n_items[0] + n_items[1];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

long long* g_attn_prof = nullptr;  // diagnostics only (mv_attn_set_profile_buffer); also read by attention_bwd.cu

}  // namespace mv

// Diagnostics (effective only in the MV_GEMM_PROFILE build): 16 int64 per CTA — softmax warp 2, lane 0: cycles [0] waiting for S,
// [1] max pass, [2] row-max exchange barrier, [3] lazy rescale, [4] exp pass, [5] TMEM store wait + arrive, [6] item epilogue,
// [7] lifetime; MMA thread: [8] waiting for Q / K|V, [9] waiting for P, [10] lifetime; [11] items of the CTA.
extern "C" void mv_attn_set_profile_buffer(void* buf) { mv::g_attn_prof = reinterpret_cast<long long*>(buf); }

// qkv bf16 [B*N, 3*D] rows = tokens (q | k | v, each heads x 64); out bf16 [B*N, D]; lse fp32 [B, heads, N] or NULL.
extern "C" int mv_attn_fwd(const void* qkv, int64_t ldqkv, void* out, int64_t ldo, float* lse, int batch, int n_tok,
                           int heads, float scale, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(qkv && out && batch > 0 && heads > 0, "mv_attn_fwd: null/empty");
  MV_CHECK_ARG(n_tok >= 16, "mv_attn_fwd: n_tok=%d must be >= 16", n_tok);
  MV_CHECK_ARG(ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "mv_attn_fwd: out alignment");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  AttnDev p;
  p.n_tok = n_tok;
  p.nb = (n_tok + 127) / 128;
  p.kb = ((n_tok + p.nb - 1) / p.nb + 15) / 16 * 16;  // balanced key blocks, multiple of 16, <= 128
  p.heads = heads;
  p.dim = heads * 64;
  p.q_tiles = (n_tok + 127) / 128;
  p.total_tiles = batch * heads * p.q_tiles;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.lse = lse;
  p.prof = g_attn_prof;
  const int smem = 12 * ATT_QBYTES + 256 + 2 * 3 * 2 * 128 * 4 + 1024;  // tiles, barriers, row-statistic exchange
  const uint64_t rows = (uint64_t)batch * n_tok;
  const CUtensorMap* tq = get_tmap_2d_bf16(qkv, rows, 3ull * p.dim, ldqkv, 128);
  const CUtensorMap* tkv = get_tmap_2d_bf16(qkv, rows, 3ull * p.dim, ldqkv, p.kb);
  if (!tq || !tkv) return MV_ERR_ARG;
  static std::atomic<uint64_t> attr{0};  // one bit per device
  if (first_use_on_device(attr)) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(attn_fwd): %s", cudaGetErrorString(e));
      return (int)e;
    }
  }
  int grid = device_sms() > 0 ? device_sms() : 148;
  if (grid > p.total_tiles) grid = p.total_tiles;
  MV_LAUNCH(attn_fwd_kernel, grid, ATT_THREADS, smem, stream, *tq, *tkv, p);
  MV_CHECK_LAUNCH("attn_fwd");
  return MV_OK;
}
