// Train-mode BatchNorm statistics of the 16 SegmentationHead gates (AttentionBlock.psi[0..1], src/generators/unet.py:
// 407-422) in closed form.  The BatchNorm input of hidden unit j is a_j = W1_j . f + b1_j, a LINEAR function of the
// 32-channel full-resolution map f, so its batch statistics follow from the first two moments of f:
//
//   mean_j = W1_j . E[f] + b1_j            var_j = W1_j Cov(f) W1_j^T,   Cov(f) = E[f f^T] - E[f] E[f]^T
//
// mv_gram32 reads f once (64 B / pixel) and accumulates  G = f^T f  [32, 32]  and  1^T f  [32]  on the warp-level tensor
// cores (mma.sync m16n8k16, fp32 accumulate) — instead of materialising / reducing a [pixels, 256] activation map.
// mv_heads_bn_from_gram turns (G, 1^T f) into the folded (scale, shift) of the gate kernel, the saved (mean, rstd) and the
// running-statistics update (torch.nn.BatchNorm2d semantics: momentum, unbiased variance).
#include "mv_host.h"
#include "mv_ptx.cuh"

namespace mv {

constexpr int GRAM_THREADS = 256;
constexpr int GRAM_CHUNK = 64;     // pixels per warp iteration
constexpr int GRAM_PITCH = 40;     // bf16 elements per staged pixel row (80 B: conflict-free ldmatrix rows)

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
template <bool F16>
__device__ __forceinline__ void mma_bf16_16816(float* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
  if (F16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// gram: fp32 [40, 32], pre-zeroed; rows 0..31 += f^T f, row 32 += 1^T f  (the layout the heads backward consumes)
template <bool F16>
__global__ void __launch_bounds__(GRAM_THREADS) gram32_kernel(const __nv_bfloat16* __restrict__ f, long long ldf, long long M,
                                                              float* __restrict__ gram) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  __shared__ __align__(16) __nv_bfloat16 stage[GRAM_THREADS / 32][GRAM_CHUNK * GRAM_PITCH];
  __shared__ float red[33 * 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 33 * 32; i += GRAM_THREADS) red[i] = 0.f;
  __syncthreads();
  __nv_bfloat16* st = stage[warp];
  const uint32_t st_addr = smem_u32(st);
  // acc[mt][nt][4]: channels 16*mt + {g, g+8} x channels 8*nt + 2t + {0,1}
  float acc[2][4][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
  float csum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) csum[j] = 0.f;

  const long long nchunks = (M + GRAM_CHUNK - 1) / GRAM_CHUNK;
  const long long wid = (long long)blockIdx.x * (GRAM_THREADS / 32) + warp;
  const long long wstep = (long long)gridDim.x * (GRAM_THREADS / 32);
  const int seg = lane & 3;   // 16-byte segment (8 channels) of a pixel row
  const int prow = lane >> 2; // pixel within a group of 8
  for (long long ch = wid; ch < nchunks; ch += wstep) {
    const long long p0 = ch * GRAM_CHUNK;
    uint4 v[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const long long p = p0 + it * 8 + prow;
      v[it] = make_uint4(0, 0, 0, 0);
      if (p < M) v[it] = *reinterpret_cast<const uint4*>(f + p * ldf + seg * 8);
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      *reinterpret_cast<uint4*>(st + (it * 8 + prow) * GRAM_PITCH + seg * 8) = v[it];
      const uint32_t w[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 x = unpack16x2<F16>(w[j]);
        csum[2 * j] += x.x;
        csum[2 * j + 1] += x.y;
      }
    }
    __syncwarp();
#pragma unroll
    for (int ks = 0; ks < GRAM_CHUNK / 16; ++ks) {
      // transposed 8x8 loads of the staged [pixel][channel] tile: matrix q = (pixels 8*(q>>1).., channels 8*(q&1)..)
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int q = lane >> 3;
        const int row = ks * 16 + (lane & 7) + 8 * (q >> 1);
        const int col = mt * 16 + 8 * (q & 1);
        ldmatrix_x4_trans(st_addr + (row * GRAM_PITCH + col) * 2, a[mt][0], a[mt][1], a[mt][2], a[mt][3]);
      }
      // the B fragment of channel block nt (k = pixel, n = channel) is the same data: regs {0, 2} of the A tile holding
      // channels 8*nt.. (even nt) or regs {1, 3} (odd nt)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          mma_bf16_16816<F16>(acc[mt][nt], a[mt][0], a[mt][1], a[mt][2], a[mt][3], a[nt >> 1][nt & 1], a[nt >> 1][2 + (nt & 1)]);
    }
    __syncwarp();
  }
  // fold the warps of this CTA in shared memory, then one global atomic per entry per CTA
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int r0 = mt * 16 + g, c0 = nt * 8 + 2 * t;
      atomicAdd(&red[r0 * 32 + c0], acc[mt][nt][0]);
      atomicAdd(&red[r0 * 32 + c0 + 1], acc[mt][nt][1]);
      atomicAdd(&red[(r0 + 8) * 32 + c0], acc[mt][nt][2]);
      atomicAdd(&red[(r0 + 8) * 32 + c0 + 1], acc[mt][nt][3]);
    }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float s = csum[j];
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 8);
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    if (lane < 4) atomicAdd(&red[32 * 32 + seg * 8 + j], s);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 33 * 32; i += GRAM_THREADS) {
    const float x = red[i];
    if (x != 0.f) atomicAdd(gram + i, x);
  }
}

// one thread per hidden unit j; all moment arithmetic in double (256 x 32 x 32 multiply-adds in total)
__global__ void heads_bn_from_gram_kernel(const float* __restrict__ gram, double count, const float* __restrict__ w1,
                                          const float* __restrict__ b1, const float* __restrict__ gamma,
                                          const float* __restrict__ beta, float* __restrict__ running_mean,
                                          float* __restrict__ running_var, float momentum, float eps, int C,
                                          float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                          float* __restrict__ rstd_out, int w1_fmt) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  __shared__ double cov[32 * 32];
  __shared__ double mf[32];
  if (threadIdx.x < 32) mf[threadIdx.x] = (double)gram[32 * 32 + threadIdx.x] / count;
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) cov[i] = (double)gram[i] / count - mf[i >> 5] * mf[i & 31];
  __syncthreads();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= C) return;
  double w[32];
#pragma unroll
  for (int a = 0; a < 32; ++a) {
    // the statistics must describe what the gate GEMM computes: W1 rounded to that GEMM's 16-bit operand format
    const float wv = w1[j * 32 + a];
    w[a] = w1_fmt == 1 ? (double)__bfloat162float(__float2bfloat16(wv)) : w1_fmt == 2 ? (double)__half2float(__float2half_rn(wv)) : (double)wv;
  }
  double lin = 0.0, var = 0.0;
#pragma unroll 4
  for (int a = 0; a < 32; ++a) {
    lin += w[a] * mf[a];
    double r = 0.0;
#pragma unroll
    for (int b = 0; b < 32; ++b) r += cov[a * 32 + b] * w[b];
    var += w[a] * r;
  }
  if (var < 0.0) var = 0.0;
  const double mean = lin + (double)b1[j];  // statistics of (W1 f + b1); the gate GEMM applies (scale, shift) to W1 f
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[j] * rstd;
  scale[j] = sc;
  shift[j] = beta[j] - (float)lin * sc;
  mean_out[j] = (float)mean;
  rstd_out[j] = rstd;
  if (running_mean) {
    running_mean[j] = (1.f - momentum) * running_mean[j] + momentum * (float)mean;
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[j] = (1.f - momentum) * running_var[j] + momentum * (float)unbiased;
  }
}

// Closed-form gradients of the 16 SegmentationHead gates (AttentionBlock.psi = conv1x1 -> BatchNorm(batch statistics) ->
// ReLU -> conv1x1 -> sigmoid, src/generators/unet.py:407-422) from three small moment matrices, one CTA:
//   E  [40, 256]: rows 0..31 = f^T e, row 32 = 1^T e   (e = gate-unit gradient masked by the ReLU, bf16 [M, 256])
//   FF [40, 32] : rows 0..31 = f^T f, row 32 = 1^T f
// with a = W1 f + b1, ahat = (a - mean) rstd, the BatchNorm backward folds into
//   dW1 = gr (w2 EF - S1/n F1 - S2/n XF),  d gamma = S2,  d beta = S1,  d w2 = scale A + shift E1
//   d f  = dt W3t + e Ca - f Mx - (K0 + K1)          (Ca, Mx, K0 + K1 are written as the operands of the two GEMMs
//                                                     that form d f: CaT bf16 [32, 256], MxN fp16/bf16 [32, 64] (times a power
//                                                     of two, 1 / that in mx_scale [32]), kshift [32])
// Everything is fp32; unit j = 16 * head + hidden index; units >= n_units are skipped.
__global__ void __launch_bounds__(256) heads_bwd_algebra_kernel(
    const float* __restrict__ E, const float* __restrict__ FF, const float* __restrict__ W1, const float* __restrict__ b1,
    const float* __restrict__ gam, const float* __restrict__ w2, const float* __restrict__ fin, float n, int n_units,
    float* __restrict__ dW1, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dw2,
    __nv_bfloat16* __restrict__ CaT, void* __restrict__ MxN, int mx_f16, float* __restrict__ mx_scale,
    float* __restrict__ kshift) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  __shared__ float F2s[32 * 33], F1s[32], Ws[256 * 33], kr[256], K01[32];
  const int j = threadIdx.x;
  for (int i = j; i < 32 * 32; i += 256) F2s[(i >> 5) * 33 + (i & 31)] = FF[i];
  if (j < 32) { F1s[j] = FF[32 * 32 + j]; K01[j] = 0.f; }
  const bool on = j < n_units;
  for (int c = 0; c < 32; ++c) Ws[j * 33 + c] = on ? W1[j * 32 + c] : 0.f;
  __syncthreads();
  const float* scale_g = fin;
  const float* shift_g = fin + 256;
  const float* mean = fin + 512;
  const float* rstd = fin + 768;
  float k2r = 0.f;
  if (on) {
    const float rs = rstd[j], bm = b1[j] - mean[j], g = gam[j] * rs, ww = w2[j];
    const float E1 = E[32 * 256 + j];
    float A = 0.f;
    for (int c = 0; c < 32; ++c) A += Ws[j * 33 + c] * E[c * 256 + j];
    const float S1 = ww * E1;
    const float S2 = ww * rs * (A + bm * E1);
    dbeta[j] = S1;
    dgamma[j] = S2;
    dw2[j] = scale_g[j] * A + shift_g[j] * E1;
    const float s1n = S1 / n, s2n = S2 / n;
    for (int c = 0; c < 32; ++c) {
      float wf = 0.f;
      for (int a = 0; a < 32; ++a) wf += Ws[j * 33 + a] * F2s[a * 33 + c];
      const float xf = rs * (wf + bm * F1s[c]);
      dW1[j * 32 + c] = g * (ww * E[c * 256 + j] - s1n * F1s[c] - s2n * xf);
      CaT[c * 256 + j] = __float2bfloat16(g * ww * Ws[j * 33 + c]);
    }
    const float k2 = g * s2n;
    k2r = k2 * rs;
    const float k01 = g * s1n + k2r * bm;  // K0 + K1 share the factor W1[j, c]
    for (int c = 0; c < 32; ++c) atomicAdd(&K01[c], k01 * Ws[j * 33 + c]);
  } else {
    for (int c = 0; c < 32; ++c) CaT[c * 256 + j] = __float2bfloat16(0.f);
  }
  kr[j] = k2r;
  __syncthreads();
  if (j < 32) kshift[j] = -K01[j];
  // MxN[r, c] = -Mx[c, r] = -sum_j W1[j, c] k2r_j W1[j, r]   (B operand [N = 32, K = 32 of pitch 64] of  dy = f (-Mx) + ...)
  // Gradient-sized values (1e-10 and below) underflow fp16: the matrix is stored times a power of two that brings its
  // largest entry into [1, 2) and the GEMM epilogue multiplies by the reciprocal (mx_scale, one value per output column).
  __shared__ float Mxs[32 * 32];
  __shared__ unsigned int mx_max_bits;
  if (j == 0) mx_max_bits = 0u;
  __syncthreads();
  for (int i = j; i < 32 * 32; i += 256) {
    const int r = i >> 5, c = i & 31;
    float acc = 0.f;
    for (int u = 0; u < n_units; ++u) acc += Ws[u * 33 + c] * kr[u] * Ws[u * 33 + r];
    Mxs[i] = -acc;
    atomicMax(&mx_max_bits, __float_as_uint(fabsf(acc)));  // non-negative floats order like their bit patterns
  }
  __syncthreads();
  const float mx_max = __uint_as_float(mx_max_bits);
  int ex = 0;
  if (mx_max > 0.f && mx_max < 3.0e38f) (void)frexpf(mx_max, &ex);  // mx_max = m * 2^ex, m in [0.5, 1)
  const float up = mx_max > 0.f ? ldexpf(1.f, 1 - ex) : 1.f;        // mx_max * up in [1, 2)
  if (j < 32) mx_scale[j] = 1.f / up;
  for (int i = j; i < 32 * 64; i += 256) {
    const int r = i >> 6, c = i & 63;
    const float v = c < 32 ? Mxs[r * 32 + c] * up : 0.f;
    if (mx_f16) reinterpret_cast<__half*>(MxN)[i] = __float2half_rn(v);
    else reinterpret_cast<__nv_bfloat16*>(MxN)[i] = __float2bfloat16(v);
  }
}

}  // namespace mv

extern "C" int mv_gram32(const void* f, int64_t ldf, int64_t m, int f16, float* gram, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(f && gram && m > 0, "mv_gram32: null/empty");
  MV_CHECK_ARG(ldf >= 32 && ldf % 8 == 0 && (reinterpret_cast<uintptr_t>(f) & 15) == 0, "mv_gram32: rows of 32 bf16, 16-byte aligned");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const long long nchunks = (m + GRAM_CHUNK - 1) / GRAM_CHUNK;
  const int sms = device_sms() > 0 ? device_sms() : 148;
  long long grid = (nchunks + (GRAM_THREADS / 32) - 1) / (GRAM_THREADS / 32);
  if (grid > 4ll * sms) grid = 4ll * sms;
  if (f16) MV_LAUNCH(gram32_kernel<true>, (unsigned)grid, GRAM_THREADS, 0, stream, reinterpret_cast<const __nv_bfloat16*>(f), ldf, m, gram);
  else MV_LAUNCH(gram32_kernel<false>, (unsigned)grid, GRAM_THREADS, 0, stream, reinterpret_cast<const __nv_bfloat16*>(f), ldf, m, gram);
  MV_CHECK_LAUNCH("gram32");
  return MV_OK;
}

extern "C" int mv_heads_bn_from_gram(const float* gram, double count, const float* w1, const float* b1, const float* gamma,
                                     const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                                     int c, float* scale, float* shift, float* mean, float* rstd, int w1_fmt,
                                     void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(gram && w1 && b1 && gamma && beta && scale && shift && mean && rstd && c > 0 && count > 0,
               "mv_heads_bn_from_gram: null/empty");
  MV_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "mv_heads_bn_from_gram: running stats go together");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MV_LAUNCH(heads_bn_from_gram_kernel, (c + 63) / 64, 64, 0, stream, gram, count, w1, b1, gamma, beta, running_mean, running_var,
                                                             momentum, eps, c, scale, shift, mean, rstd, w1_fmt);
  MV_CHECK_LAUNCH("heads_bn_from_gram");
  return MV_OK;
}

extern "C" int mv_heads_bwd_algebra(const float* E, const float* FF, const float* w1, const float* b1, const float* gamma,
                                    const float* w2, const float* fin, double count, int n_units, float* dw1, float* dgamma,
                                    float* dbeta, float* dw2, void* ca_t_bf16, void* mx_n, int mx_f16, float* mx_scale,
                                    float* kshift, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(E && FF && w1 && b1 && gamma && w2 && fin && dw1 && dgamma && dbeta && dw2 && ca_t_bf16 && mx_n && mx_scale && kshift,
               "mv_heads_bwd_algebra: null pointer");
  MV_CHECK_ARG(n_units > 0 && n_units <= 256 && count > 0, "mv_heads_bwd_algebra: n_units in 1..256");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MV_LAUNCH(heads_bwd_algebra_kernel, 1, 256, 0, stream, E, FF, w1, b1, gamma, w2, fin, (float)count, n_units, dw1, dgamma,
            dbeta, dw2, reinterpret_cast<__nv_bfloat16*>(ca_t_bf16), mx_n, mx_f16, mx_scale, kshift);
  MV_CHECK_LAUNCH("heads_bwd_algebra");
  return MV_OK;
}
