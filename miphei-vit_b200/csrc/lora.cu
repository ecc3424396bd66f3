// LoRA gradient path of QkvWithLoRA (src/generators/lora.py:16-18,29-33) for one transformer block.
//
// Forward (engine): T = xn [A_q | A_v]  (M x 16, stored in the 16 extension columns of the LayerNorm output) and the
// QKV GEMM runs over K = D + 16 with the weight rows extended by alpha*B_q^T (q rows) / alpha*B_v^T (v rows).
// Backward: dT = [dQ (alpha B_q)^T | dV (alpha B_v)^T] (16 extension columns of the dQKV buffer, produced by a skinny
// tensor-core GEMM), and here
//     dA_q = xn^T dT[:, 0:8]      dA_v = xn^T dT[:, 8:16]          (reduction over the M tokens)
//     dB_q = alpha T[:, 0:8]^T dQ dB_v = alpha T[:, 8:16]^T dV
// All four are [16 x D] products whose contraction runs over the M tokens, with a 16-row left operand: exactly one
// mma.sync m16n8k16 tile high.  lora_partials_kernel streams the token rows through shared memory ONCE (cp.async, two
// stages), takes both operands with ldmatrix.trans (the token index is the slow one of both) and leaves per-slice partial
// sums in a workspace; lora_reduce_kernel adds the slices and writes the four gradients in parameter layout.  (Round 1
// ran this as two transposes, a memset, three split-K tcgen05 GEMMs whose 128-row A tile held 16 real rows, and an unpack:
// six launches, 72 us per block at M = 10528; widths the kernel is not instantiated for still take that path.)
#include "mv_host.h"
#include "mv_ptx.cuh"
#include <stdlib.h>
#include <string.h>

namespace mv {

// src0/src1: bf16 [M, 16] slices (row pitch ld0/ld1) -> dst0/dst1: bf16 [16, ldt] (row r = column r of the source)
__global__ void transpose16_kernel(const __nv_bfloat16* __restrict__ src0, long long ld0,
                                   const __nv_bfloat16* __restrict__ src1, long long ld1, __nv_bfloat16* __restrict__ dst0,
                                   __nv_bfloat16* __restrict__ dst1, long long ldt, int M) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const __nv_bfloat16* s = blockIdx.y == 0 ? src0 : src1;
  const long long ld = blockIdx.y == 0 ? ld0 : ld1;
  __nv_bfloat16* d = blockIdx.y == 0 ? dst0 : dst1;
  const uint4 a = *reinterpret_cast<const uint4*>(s + (long long)m * ld);
  const uint4 b = *reinterpret_cast<const uint4*>(s + (long long)m * ld + 8);
  const __nv_bfloat16* pa = reinterpret_cast<const __nv_bfloat16*>(&a);
  const __nv_bfloat16* pb = reinterpret_cast<const __nv_bfloat16*>(&b);
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    d[(long long)r * ldt + m] = pa[r];
    d[(long long)(8 + r) * ldt + m] = pb[r];
  }
}

// g_a: fp32 [16, D] = dAcat^T; g_b: fp32 [16, 3D] -> dA_q [D, 8], dA_v [D, 8], dB_q [8, D], dB_v [8, D]
__global__ void lora_unpack_kernel(const float* __restrict__ g_a, const float* __restrict__ g_b, float* __restrict__ dAq,
                                   float* __restrict__ dAv, float* __restrict__ dBq, float* __restrict__ dBv, int D,
                                   float alpha) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // 0 .. 8*D
  if (i >= 8 * D) return;
  const int r = i / D, d = i - r * D;
  dAq[d * 8 + r] = g_a[(long long)r * D + d];
  dAv[d * 8 + r] = g_a[(long long)(8 + r) * D + d];
  dBq[(long long)r * D + d] = alpha * g_b[(long long)r * 3 * D + d];
  dBv[(long long)r * D + d] = alpha * g_b[(long long)(8 + r) * 3 * D + 2 * D + d];
}

// LoRA operand refresh for ALL blocks in one launch.  src: flat fp32 [depth, 4, 8*D] = per block (A_q [D,8], B_q [8,D],
// A_v [D,8], B_v [8,D]) in parameter layout; ptrs: int64 [depth, 4] device pointers to the block's bf16 operands
//   acat [16, D]            = [A_q | A_v]^T                                     (B operand of T = xn [A_q | A_v])
//   wqkv_ext [3D, ldw]      columns D..D+15: alpha B_q^T on the q rows, alpha B_v^T on the v rows (K-extended QKV weight)
//   wqkv_bwd_ext [D, ldb]   columns 3D..3D+15: [A_q | A_v]                     (K-extended dX weight; may be null)
//   bcat [16, 3D]           = [alpha B_q ; 0 ; alpha B_v] blocks                (B operand of dT; may be null)
__global__ void lora_refresh_kernel(const float* __restrict__ src, const long long* __restrict__ ptrs, int D, float alpha,
                                    long long ldw, long long ldb) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  const int blk = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 8 * D) return;
  const int r = e / D, d = e - r * D;
  const float* s = src + (long long)blk * 32 * D;
  const float aq = s[d * 8 + r], av = s[16ll * D + d * 8 + r];
  const float bq = alpha * s[8ll * D + (long long)r * D + d], bv = alpha * s[24ll * D + (long long)r * D + d];
  __nv_bfloat16* acat = reinterpret_cast<__nv_bfloat16*>(ptrs[blk * 4 + 0]);
  __nv_bfloat16* wext = reinterpret_cast<__nv_bfloat16*>(ptrs[blk * 4 + 1]);
  __nv_bfloat16* wbwd = reinterpret_cast<__nv_bfloat16*>(ptrs[blk * 4 + 2]);
  __nv_bfloat16* bcat = reinterpret_cast<__nv_bfloat16*>(ptrs[blk * 4 + 3]);
  acat[(long long)r * D + d] = __float2bfloat16(aq);
  acat[(long long)(8 + r) * D + d] = __float2bfloat16(av);
  wext[(long long)d * ldw + D + r] = __float2bfloat16(bq);
  wext[(long long)(2 * D + d) * ldw + D + 8 + r] = __float2bfloat16(bv);
  if (wbwd) {
    wbwd[(long long)d * ldb + 3 * D + r] = __float2bfloat16(aq);
    wbwd[(long long)d * ldb + 3 * D + 8 + r] = __float2bfloat16(av);
  }
  if (bcat) {
    bcat[(long long)r * 3 * D + d] = __float2bfloat16(bq);
    bcat[(long long)(8 + r) * 3 * D + 2 * D + d] = __float2bfloat16(bv);
  }
}


// ------------------------------------------------------------------ mma.sync path
constexpr int LG_THREADS = 256;   // 8 warps, each owning D / 8 output columns
constexpr int LG_STAGES = 3;   // two 16-token stages in flight while one is consumed (one CTA per SM)
constexpr int LG_SPITCH = 48;     // bytes per token row of the small [16 tokens x 16] tile (32 used; conflict-free ldmatrix)

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// grid (slices, 3): z = 0: dT^T xn, z = 1: T^T dQ, z = 2: T^T dV  ->  part[slice][z][16][D] fp32 (every element written).
// NT = D / 64: n8 tiles per warp.
template <int NT>
__global__ void __launch_bounds__(LG_THREADS) lora_partials_kernel(const __nv_bfloat16* __restrict__ xe, long long ldx,
                                                                   const __nv_bfloat16* __restrict__ qe, long long ldq, int M,
                                                                   int tok_per_slice, float* __restrict__ part) {
  extern __shared__ uint8_t lg_smem[];
  constexpr int D = NT * 64;
  constexpr int BPITCH = D * 2 + 16;  // bytes per token row of the big tile: 16-byte skew -> conflict-free ldmatrix
  constexpr int STAGE = 16 * BPITCH + 16 * LG_SPITCH;
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  const int z = blockIdx.y, slice = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t_begin = slice * tok_per_slice;
  const int t_end = min(M, t_begin + tok_per_slice);
  const __nv_bfloat16* big = z == 0 ? xe : (z == 1 ? qe : qe + 2 * D);
  const long long ldbig = z == 0 ? ldx : ldq;
  const __nv_bfloat16* sml = z == 0 ? qe + 3 * D : xe + D;
  const long long ldsml = z == 0 ? ldq : ldx;
  const uint32_t sbase = smem_u32(lg_smem);
  const int nsteps = (t_end - t_begin + 15) / 16;

  auto load_stage = [&](int st, int step) {
    const int tok0 = t_begin + step * 16;
    const uint32_t sb = sbase + st * STAGE, ss = sb + 16 * BPITCH;
    constexpr int CH = D / 8;  // 16-byte chunks per token row
    for (int i = tid; i < 16 * CH; i += LG_THREADS) {
      const int r = i / CH, c = i - r * CH;
      const int tok = tok0 + r;
      const bool ok = tok < t_end;  // rows beyond the slice are zero-filled: they add nothing
      cp_async_16(sb + r * BPITCH + c * 16, big + (long long)(ok ? tok : t_begin) * ldbig + c * 8, ok ? 16 : 0);
    }
    if (tid < 32) {
      const int r = tid >> 1, c = tid & 1;
      const int tok = tok0 + r;
      const bool ok = tok < t_end;
      cp_async_16(ss + r * LG_SPITCH + c * 16, sml + (long long)(ok ? tok : t_begin) * ldsml + c * 8, ok ? 16 : 0);
    }
    cp_async_commit();
  };

  float acc[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }

  // prologue: LG_STAGES - 1 stages in flight (empty commit groups keep the group count uniform at the tail)
#pragma unroll
  for (int i = 0; i < LG_STAGES - 1; ++i) {
    if (i < nsteps) load_stage(i, i); else cp_async_commit();
  }
  for (int step = 0; step < nsteps; ++step) {
    const int st = step % LG_STAGES;
    cp_async_wait<LG_STAGES - 2>();  // this thread's copies of stage `step` have landed ...
    __syncthreads();                 // ... and everybody's; every warp has also finished reading the stage refilled below
    if (step + LG_STAGES - 1 < nsteps) load_stage((step + LG_STAGES - 1) % LG_STAGES, step + LG_STAGES - 1);
    else cp_async_commit();
    const uint32_t sb = sbase + st * STAGE, ss = sb + 16 * BPITCH;
    // A (m = 16 LoRA columns, k = 16 tokens) from the token-major small tile: four transposed 8 x 8 blocks
    uint32_t a[4];
    {
      const int idx = lane >> 3, row = lane & 7;
      ldmatrix_x4_trans(ss + ((idx >> 1) * 8 + row) * LG_SPITCH + (idx & 1) * 16, a);
    }
    const int col0 = warp * (NT * 8);
#pragma unroll
    for (int j = 0; j < NT; j += 2) {
      // B (k = 16 tokens, n = 8 columns) x 2 column tiles: blocks (k 0-7, tile j), (k 8-15, tile j), (k 0-7, j+1), (k 8-15, j+1)
      uint32_t b[4];
      const int idx = lane >> 3, row = lane & 7;
      ldmatrix_x4_trans(sb + ((idx & 1) * 8 + row) * BPITCH + (col0 + (j + (idx >> 1)) * 8) * 2, b);
      mma_bf16_16816(acc[j], a, b[0], b[1]);
      mma_bf16_16816(acc[j + 1], a, b[2], b[3]);
    }
  }
  float* out = part + ((long long)(slice * 3 + z) * 16) * D;
  const int r0 = lane >> 2, c0 = warp * (NT * 8) + (lane & 3) * 2;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    *reinterpret_cast<float2*>(out + (long long)r0 * D + c0 + j * 8) = make_float2(acc[j][0], acc[j][1]);
    *reinterpret_cast<float2*>(out + (long long)(r0 + 8) * D + c0 + j * 8) = make_float2(acc[j][2], acc[j][3]);
  }
}

// part [slices][3][16][D] -> dA_q [D, 8], dA_v [D, 8], dB_q [8, D], dB_v [8, D].  CTA = (r, 32 consecutive d): its 8 warps
// each sum every 8th slice (all loads independent), then add up through shared memory.
__global__ void __launch_bounds__(256) lora_reduce_kernel(const float* __restrict__ part, int slices, int D, float alpha,
                                                          float* __restrict__ dAq, float* __restrict__ dAv,
                                                          float* __restrict__ dBq, float* __restrict__ dBv) {
  __shared__ float red[8][4][32];
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.y, d = blockIdx.x * 32 + lane;
  const long long zs = 16ll * D, ss = 3 * zs;
  const float* p0 = part + (long long)r * D + d;
  const long long o_av = 8ll * D, o_bq = zs, o_bv = 2 * zs + 8ll * D;
  float aq = 0.f, av = 0.f, bq = 0.f, bv = 0.f;
  if (d < D) {
#pragma unroll 4
    for (int s = warp; s < slices; s += 8) {
      const float* p = p0 + (long long)s * ss;
      aq += __ldcg(p); av += __ldcg(p + o_av); bq += __ldcg(p + o_bq); bv += __ldcg(p + o_bv);
    }
  }
  red[warp][0][lane] = aq; red[warp][1][lane] = av; red[warp][2][lane] = bq; red[warp][3][lane] = bv;
  __syncthreads();
  if (warp < 4 && d < D) {  // warp w finishes quantity w
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][warp][lane];
    if (warp == 0) dAq[d * 8 + r] = t;
    else if (warp == 1) dAv[d * 8 + r] = t;
    else if (warp == 2) dBq[(long long)r * D + d] = alpha * t;
    else dBv[(long long)r * D + d] = alpha * t;
  }
}

// tokens per slice (multiple of 16) so that slices x 3 CTAs fill the SMs once (150 KB of stages: one CTA per SM)
static int lora_tok_per_slice(int m) {
  const int sms = device_sms() > 0 ? device_sms() : 148;
  int slices = sms / 3;
  if (slices < 1) slices = 1;
  int tps = ((m + slices - 1) / slices + 15) / 16 * 16;
  if (tps < 16) tps = 16;
  return tps;
}
static bool lora_mma_width(int d) {
  switch (d / 64) { case 2: case 4: case 6: case 8: case 12: case 16: case 24: return d % 64 == 0; default: return false; }
}

}  // namespace mv

extern "C" int64_t mv_lora_grads_workspace_bytes(int m, int d) {
  const int64_t ldt = (m + 7) / 8 * 8;
  const int64_t legacy = 2 * 16 * ldt * 2 /*T^T, dT^T bf16*/ + 16ll * d * 4 + 16ll * 3 * d * 4 + 256;
  if (!mv::lora_mma_width(d) || m <= 0) return legacy;
  const int tps = mv::lora_tok_per_slice(m);
  const int64_t slices = (m + tps - 1) / tps;
  const int64_t part = slices * 3 * 16 * d * 4;  // per-slice partial sums
  return part > legacy ? part : legacy;
}

// xn_ext bf16 [M, >= D+16] (LayerNorm output, T in columns D..D+15), dqkv_ext bf16 [M, >= 3D+16] (dT in columns
// 3D..3D+15).  Outputs are written (not accumulated).
extern "C" int mv_lora_grads(const void* xn_ext, int64_t ldx, const void* dqkv_ext, int64_t ldq, int m, int d, float alpha,
                             float* dA_q, float* dA_v, float* dB_q, float* dB_v, void* workspace, int64_t workspace_bytes,
                             void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(xn_ext && dqkv_ext && dA_q && dA_v && dB_q && dB_v && workspace && m > 0 && d > 0, "mv_lora_grads: null/empty");
  MV_CHECK_ARG(workspace_bytes >= mv_lora_grads_workspace_bytes(m, d), "mv_lora_grads: workspace too small");
  MV_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && ldx % 8 == 0 && ldq % 8 == 0 && d % 8 == 0,
               "mv_lora_grads: alignment");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  static const int mma_env = [] { const char* e = getenv("MV_LORA_MMA"); return e ? atoi(e) : 1; }();  // 0: round-1 path
  if (mma_env != 0 && lora_mma_width(d)) {
    const int tps = lora_tok_per_slice(m);
    const int slices = (m + tps - 1) / tps;
    float* part = reinterpret_cast<float*>(workspace);
    const __nv_bfloat16* xe_ = reinterpret_cast<const __nv_bfloat16*>(xn_ext);
    const __nv_bfloat16* qe_ = reinterpret_cast<const __nv_bfloat16*>(dqkv_ext);
    dim3 grid(slices, 3);
#define MV_LORA_LAUNCH(NT_)                                                                                          \
  case NT_: {                                                                                                        \
    constexpr int smem_ = LG_STAGES * (16 * (NT_ * 64 * 2 + 16) + 16 * LG_SPITCH);                                   \
    static std::atomic<uint64_t> attr_{0};                                                                           \
    if (first_use_on_device(attr_)) {                                                                                \
      cudaError_t e_ = cudaFuncSetAttribute(lora_partials_kernel<NT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_); \
      if (e_ != cudaSuccess) { set_error("cudaFuncSetAttribute(lora_partials): %s", cudaGetErrorString(e_)); return (int)e_; } \
    }                                                                                                                \
    MV_LAUNCH(lora_partials_kernel<NT_>, grid, LG_THREADS, smem_, stream, xe_, (long long)ldx, qe_, (long long)ldq, m, tps, part); \
  } break;
    switch (d / 64) {
      MV_LORA_LAUNCH(2) MV_LORA_LAUNCH(4) MV_LORA_LAUNCH(6) MV_LORA_LAUNCH(8) MV_LORA_LAUNCH(12) MV_LORA_LAUNCH(16) MV_LORA_LAUNCH(24)
      default: break;
    }
#undef MV_LORA_LAUNCH
    MV_CHECK_LAUNCH("lora_partials");
    MV_LAUNCH(lora_reduce_kernel, dim3((d + 31) / 32, 8), 256, 0, stream, part, slices, d, alpha, dA_q, dA_v, dB_q, dB_v);
    MV_CHECK_LAUNCH("lora_reduce");
    return MV_OK;
  }
  const int64_t ldt = (m + 7) / 8 * 8;
  __nv_bfloat16* tT = reinterpret_cast<__nv_bfloat16*>(workspace);
  __nv_bfloat16* dtT = tT + 16 * ldt;
  float* g_a = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + ((2 * 16 * ldt * 2 + 255) / 256) * 256);
  float* g_b = g_a + 16ll * d;
  const __nv_bfloat16* xe = reinterpret_cast<const __nv_bfloat16*>(xn_ext);
  const __nv_bfloat16* qe = reinterpret_cast<const __nv_bfloat16*>(dqkv_ext);
  dim3 grid((m + 255) / 256, 2);
  MV_LAUNCH(transpose16_kernel, grid, 256, 0, stream, xe + d, ldx, qe + 3ll * d, ldq, tT, dtT, ldt, m);
  MV_CHECK_LAUNCH("transpose16");
  cudaError_t e = cudaMemsetAsync(g_a, 0, (16ll * d + 16ll * 3 * d) * 4, stream);
  if (e != cudaSuccess) {
    set_error("cudaMemsetAsync: %s", cudaGetErrorString(e));
    return (int)e;
  }
  mv_gemm_args a;
  memset(&a, 0, sizeof(a));
  a.mode = MV_GEMM_NN_ATOMIC;
  a.out_f32 = 1;
  a.m = 16;
  a.k = m;
  // dAcat^T [16, D] = dT^T [16, M] . xn [M, D]
  a.a = dtT; a.lda = ldt; a.b = xn_ext; a.ldb = ldx; a.n = d; a.out = g_a; a.ldo = d;
  int rc = mv_gemm_bf16(&a, stream_);
  if (rc) return rc;
  // dBcat [16, 3D] = T^T [16, M] . dQKV [M, 3D]: only the q and v column thirds are used (B_k does not exist), so the
  // dK third of dQKV is never streamed
  a.a = tT; a.b = dqkv_ext; a.ldb = ldq; a.n = d; a.out = g_b; a.ldo = 3 * d;
  rc = mv_gemm_bf16(&a, stream_);
  if (rc) return rc;
  a.b = qe + 2ll * d; a.out = g_b + 2ll * d;
  rc = mv_gemm_bf16(&a, stream_);
  if (rc) return rc;
  MV_LAUNCH(lora_unpack_kernel, (8 * d + 255) / 256, 256, 0, stream, g_a, g_b, dA_q, dA_v, dB_q, dB_v, d, alpha);
  MV_CHECK_LAUNCH("lora_unpack");
  return MV_OK;
}

extern "C" int mv_lora_refresh(const float* lora_flat, const int64_t* ptrs, int depth, int d, float alpha, int64_t ldw,
                               int64_t ldb, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(lora_flat && ptrs && depth > 0 && d > 0 && d % 8 == 0, "mv_lora_refresh: null/empty");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  dim3 grid((8 * d + 255) / 256, depth);
  MV_LAUNCH(lora_refresh_kernel, grid, 256, 0, stream, lora_flat, reinterpret_cast<const long long*>(ptrs), d, alpha,
            (long long)ldw, (long long)ldb);
  MV_CHECK_LAUNCH("lora_refresh");
  return MV_OK;
}
