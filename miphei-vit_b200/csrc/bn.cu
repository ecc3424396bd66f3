// Train-mode BatchNorm2d (+ReLU) around the implicit-GEMM convolutions (Basic_Conv3x3, src/generators/mipheivit.py:
// 20-41; AttentionBlock.psi[1], src/generators/unet.py:407-422), NHWC bf16 activations viewed as [M = B*H*W, C] rows.
//
//   forward : the conv GEMM epilogue accumulates per-channel (sum, sum of squares) of the stored raw output z
//             -> mv_bn_finalize: batch mean / biased variance -> (scale, shift) for y = relu(z*scale + shift) and the
//                running-statistics update (momentum 0.1, unbiased variance; torch.nn.BatchNorm2d semantics)
//             -> mv_bn_relu_apply: y = relu(z*scale + shift)                                     (HBM-bound, 16-byte vectors)
//   backward: dzh = dy * [y > 0];  S1 = sum dzh, S2 = sum dzh * xhat  (mv_bn_relu_bwd_stats, column reduction)
//             dz = gamma * rstd * (dzh - S1/n - xhat * S2/n);  dgamma = S2, dbeta = S1           (mv_bn_relu_bwd_apply)
#include "mv_host.h"
#include "mv_ptx.cuh"

namespace mv {

__global__ void bn_finalize_kernel(const float* __restrict__ colstats, float count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, const float* __restrict__ pre_bias,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
                                   float eps, int C, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mean = colstats[c] / count;
  const float var = fmaxf(colstats[C + c] / count - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  const float sc = gamma[c] * rstd;
  // the statistics were taken on (acc + pre_bias); the GEMM that applies (scale, shift) sees acc only
  const float pb = pre_bias ? pre_bias[c] : 0.f;
  scale[c] = sc;
  shift[c] = beta[c] + (pb - mean) * sc;
  mean_out[c] = mean;
  rstd_out[c] = rstd;
  if (running_mean) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    const float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
  }
}

// y > 0 on the raw 16-bit pattern: sign clear and not zero — the same test for bf16 and fp16 storage (ReLU outputs
// are never NaN / inf), so the backward kernels take either format of the saved activation
__device__ __forceinline__ bool pos16(uint32_t bits) { return bits != 0u && (bits & 0x8000u) == 0u; }

// 8 consecutive channels of row-major z (bf16 or fp32) at vector index i
template <bool ZF32>
__device__ __forceinline__ void load_z8(const void* z, long long i, float* f) {
  if (ZF32) {
    const float4 a = reinterpret_cast<const float4*>(z)[2 * i], b = reinterpret_cast<const float4*>(z)[2 * i + 1];
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    const uint4 u = reinterpret_cast<const uint4*>(z)[i];
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), d = unpack_bf16x2(u.z), e = unpack_bf16x2(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = d.x; f[5] = d.y; f[6] = e.x; f[7] = e.y;
  }
}

// one thread = 8 consecutive channels of one row
template <bool ZF32, bool F16>
__global__ void bn_relu_apply_kernel(const void* __restrict__ z, const float* __restrict__ scale,
                                     const float* __restrict__ shift, __nv_bfloat16* __restrict__ y, long long M, int C) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  const int cg = C / 8;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * cg) return;
  const int c = (int)(i % cg) * 8;
  float f[8];
  load_z8<ZF32>(z, i, f);
  const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c)), s1 = __ldg(reinterpret_cast<const float4*>(scale + c + 4));
  const float4 h0 = __ldg(reinterpret_cast<const float4*>(shift + c)), h1 = __ldg(reinterpret_cast<const float4*>(shift + c + 4));
  uint4 o;
  o.x = pack16x2<F16>(fmaxf(f[0] * s0.x + h0.x, 0.f), fmaxf(f[1] * s0.y + h0.y, 0.f));
  o.y = pack16x2<F16>(fmaxf(f[2] * s0.z + h0.z, 0.f), fmaxf(f[3] * s0.w + h0.w, 0.f));
  o.z = pack16x2<F16>(fmaxf(f[4] * s1.x + h1.x, 0.f), fmaxf(f[5] * s1.y + h1.y, 0.f));
  o.w = pack16x2<F16>(fmaxf(f[6] * s1.z + h1.z, 0.f), fmaxf(f[7] * s1.w + h1.w, 0.f));
  reinterpret_cast<uint4*>(y)[i] = o;
}

constexpr int BNB_THREADS = 256;
constexpr int BNB_ROWS = 512;  // rows per block

// sums[0..C) += sum_rows dzh, sums[C..2C) += sum_rows dzh * xhat.  blockDim.x = 256 = (C/8) column groups x row lanes
template <bool ZF32>
__global__ void __launch_bounds__(BNB_THREADS) bn_relu_bwd_stats_kernel(
    const __nv_bfloat16* __restrict__ dy, long long lddy, const __nv_bfloat16* __restrict__ y,
    const void* __restrict__ z, const float* __restrict__ mean, const float* __restrict__ rstd,
    float* __restrict__ sums, long long M, int C) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  extern __shared__ float red[];  // [2][256][8]
  const int cg = C / 8;
  const int cgi = threadIdx.x % cg;
  const int rl = threadIdx.x / cg;
  const int rlanes = BNB_THREADS / cg;
  const int c = cgi * 8;
  float s1[8], s2[8], mu[8], rs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; mu[j] = mean[c + j]; rs[j] = rstd[c + j]; }
  const long long r0 = (long long)blockIdx.x * BNB_ROWS;
  const long long r1 = r0 + BNB_ROWS < M ? r0 + BNB_ROWS : M;
  if (rl < rlanes) {
    for (long long r = r0 + rl; r < r1; r += rlanes) {
      const uint4 ud = *reinterpret_cast<const uint4*>(dy + r * lddy + c);
      const uint4 uy = *reinterpret_cast<const uint4*>(y + r * C + c);
      float fz[8];
      load_z8<ZF32>(z, (r * C + c) / 8, fz);
      const uint32_t* pd = &ud.x; const uint32_t* py = &uy.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 fd = unpack_bf16x2(pd[j]);
        const float g0 = pos16(py[j] & 0xFFFFu) ? fd.x : 0.f, g1 = pos16(py[j] >> 16) ? fd.y : 0.f;
        s1[2 * j] += g0; s1[2 * j + 1] += g1;
        s2[2 * j] += g0 * (fz[2 * j] - mu[2 * j]) * rs[2 * j];
        s2[2 * j + 1] += g1 * (fz[2 * j + 1] - mu[2 * j + 1]) * rs[2 * j + 1];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[threadIdx.x * 8 + j] = s1[j];
    red[(BNB_THREADS + threadIdx.x) * 8 + j] = s2[j];
  }
  __syncthreads();
  // thread t < C reduces column t over the row lanes
  for (int col = threadIdx.x; col < C; col += blockDim.x) {
    const int g = col / 8, j = col % 8;
    float a = 0.f, b = 0.f;
    for (int l = 0; l < rlanes; ++l) {
      a += red[(l * cg + g) * 8 + j];
      b += red[(BNB_THREADS + l * cg + g) * 8 + j];
    }
    atomicAdd(sums + col, a);
    atomicAdd(sums + C + col, b);
  }
}

template <bool ZF32>
__global__ void bn_relu_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy,
                                         const __nv_bfloat16* __restrict__ y, const void* __restrict__ z,
                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                         const float* __restrict__ gamma, const float* __restrict__ sums, float inv_n,
                                         __nv_bfloat16* __restrict__ dz, long long M, int C) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  const int cg = C / 8;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * cg) return;
  const int c = (int)(i % cg) * 8;
  const long long r = i / cg;
  const uint4 ud = *reinterpret_cast<const uint4*>(dy + r * lddy + c);
  const uint4 uy = reinterpret_cast<const uint4*>(y)[i];
  float fz[8];
  load_z8<ZF32>(z, i, fz);
  const uint32_t* pd = &ud.x; const uint32_t* py = &uy.x;
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 fd = unpack_bf16x2(pd[j]);
    float r2[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int cc = c + 2 * j + e;
      const float g = pos16(e ? py[j] >> 16 : py[j] & 0xFFFFu) ? (e ? fd.y : fd.x) : 0.f;
      const float xh = (fz[2 * j + e] - mean[cc]) * rstd[cc];
      r2[e] = gamma[cc] * rstd[cc] * (g - sums[cc] * inv_n - xh * sums[C + cc] * inv_n);
    }
    o[j] = pack_bf16x2(r2[0], r2[1]);
  }
  reinterpret_cast<uint4*>(dz)[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

}  // namespace mv

extern "C" int mv_bn_finalize(const float* colstats, double count, const float* gamma, const float* beta,
                              const float* pre_bias, float* running_mean, float* running_var, float momentum, float eps,
                              int c, float* scale, float* shift, float* mean, float* rstd, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(colstats && gamma && beta && scale && shift && mean && rstd && c > 0 && count > 0, "mv_bn_finalize: null/empty");
  MV_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "mv_bn_finalize: running stats go together");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MV_LAUNCH(bn_finalize_kernel, (c + 127) / 128, 128, 0, stream, colstats, (float)count, gamma, beta, pre_bias, running_mean,
                                                          running_var, momentum, eps, c, scale, shift, mean, rstd);
  MV_CHECK_LAUNCH("bn_finalize");
  return MV_OK;
}

extern "C" int mv_bn_relu_apply(const void* z, int z_f32, const float* scale, const float* shift, void* y, int y_f16,
                                int64_t m, int c, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(z && scale && shift && y && m > 0 && c % 8 == 0, "mv_bn_relu_apply: shape");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const long long total = m * (c / 8);
  const unsigned agrid = (unsigned)((total + 255) / 256);
  __nv_bfloat16* yy = reinterpret_cast<__nv_bfloat16*>(y);
  if (z_f32 && y_f16) MV_LAUNCH((bn_relu_apply_kernel<true, true>), agrid, 256, 0, stream, z, scale, shift, yy, m, c);
  else if (z_f32) MV_LAUNCH((bn_relu_apply_kernel<true, false>), agrid, 256, 0, stream, z, scale, shift, yy, m, c);
  else if (y_f16) MV_LAUNCH((bn_relu_apply_kernel<false, true>), agrid, 256, 0, stream, z, scale, shift, yy, m, c);
  else MV_LAUNCH((bn_relu_apply_kernel<false, false>), agrid, 256, 0, stream, z, scale, shift, yy, m, c);
  MV_CHECK_LAUNCH("bn_relu_apply");
  return MV_OK;
}

// sums: fp32 [2, C], zeroed by this call
extern "C" int mv_bn_relu_bwd(const void* dy, int64_t lddy, const void* y, const void* z, int z_f32, const float* mean,
                              const float* rstd, const float* gamma, float* sums, void* dz, int64_t m, int c,
                              void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(dy && y && z && mean && rstd && gamma && sums && dz && m > 0, "mv_bn_relu_bwd: null/empty");
  MV_CHECK_ARG(c % 8 == 0 && c <= 2048 && lddy % 8 == 0, "mv_bn_relu_bwd: C %% 8 == 0, C <= 2048 (C=%d)", c);
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  cudaError_t e = cudaMemsetAsync(sums, 0, 2 * c * sizeof(float), stream);
  if (e != cudaSuccess) {
    set_error("cudaMemsetAsync: %s", cudaGetErrorString(e));
    return (int)e;
  }
  const int smem = 2 * BNB_THREADS * 8 * 4;
  const unsigned sgrid = (unsigned)((m + BNB_ROWS - 1) / BNB_ROWS);
  if (z_f32) MV_LAUNCH((bn_relu_bwd_stats_kernel<true>), sgrid, BNB_THREADS, smem, stream, 
      reinterpret_cast<const __nv_bfloat16*>(dy), lddy, reinterpret_cast<const __nv_bfloat16*>(y), z, mean, rstd, sums, m, c);
  else MV_LAUNCH((bn_relu_bwd_stats_kernel<false>), sgrid, BNB_THREADS, smem, stream, 
      reinterpret_cast<const __nv_bfloat16*>(dy), lddy, reinterpret_cast<const __nv_bfloat16*>(y), z, mean, rstd, sums, m, c);
  MV_CHECK_LAUNCH("bn_relu_bwd_stats");
  const long long total = m * (c / 8);
  if (z_f32) MV_LAUNCH((bn_relu_bwd_apply_kernel<true>), (unsigned)((total + 255) / 256), 256, 0, stream, 
      reinterpret_cast<const __nv_bfloat16*>(dy), lddy, reinterpret_cast<const __nv_bfloat16*>(y), z, mean, rstd, gamma, sums,
      (float)(1.0 / (double)m), reinterpret_cast<__nv_bfloat16*>(dz), m, c);
  else MV_LAUNCH((bn_relu_bwd_apply_kernel<false>), (unsigned)((total + 255) / 256), 256, 0, stream, 
      reinterpret_cast<const __nv_bfloat16*>(dy), lddy, reinterpret_cast<const __nv_bfloat16*>(y), z, mean, rstd, gamma, sums,
      (float)(1.0 / (double)m), reinterpret_cast<__nv_bfloat16*>(dz), m, c);
  MV_CHECK_LAUNCH("bn_relu_bwd_apply");
  return MV_OK;
}
