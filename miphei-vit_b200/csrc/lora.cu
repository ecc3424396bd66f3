// LoRA gradient path of QkvWithLoRA (src/generators/lora.py:16-18,29-33) for one transformer block.
//
// Forward (engine): T = xn [A_q | A_v]  (M x 16, stored in the 16 extension columns of the LayerNorm output) and the
// QKV GEMM runs over K = D + 16 with the weight rows extended by alpha*B_q^T (q rows) / alpha*B_v^T (v rows).
// Backward: dT = [dQ (alpha B_q)^T | dV (alpha B_v)^T] (16 extension columns of the dQKV buffer, produced by a skinny
// tensor-core GEMM), and here
//     dA_q = xn^T dT[:, 0:8]      dA_v = xn^T dT[:, 8:16]          (reduction over the M tokens)
//     dB_q = alpha T[:, 0:8]^T dQ dB_v = alpha T[:, 8:16]^T dV
// computed as two split-K tensor-core GEMMs (MV_GEMM_NN_ATOMIC) on the transposed skinny operands.
#include "mv_host.h"
#include "mv_ptx.cuh"
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

}  // namespace mv

extern "C" int64_t mv_lora_grads_workspace_bytes(int m, int d) {
  const int64_t ldt = (m + 7) / 8 * 8;
  return 2 * 16 * ldt * 2 /*T^T, dT^T bf16*/ + 16ll * d * 4 + 16ll * 3 * d * 4 + 256;
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
