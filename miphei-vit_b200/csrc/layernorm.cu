// LayerNorm over the channel dimension of token rows (timm Block.norm1 / norm2 and the final norm; eps 1e-6,
// created at src/generators/foundation_models.py:53-57).  HBM-bound: one warp per row, the row lives in registers,
// two-pass statistics, float4 loads, 16-byte bf16 stores, warp-shuffle reductions.
//
//   forward : x fp32 [M, D] (residual stream) -> y bf16 [M, ldy] (+ optional mean / rstd for the backward pass)
//   backward: dx_out = dres + LN'(x; w) . dy     (frozen affine parameters: no dgamma / dbeta)
//             dres is the gradient flowing along the residual branch, dy the gradient of the LN output.
#include <stdlib.h>

#include "mv_host.h"
#include "mv_ptx.cuh"

namespace mv {

constexpr int LN_WARPS = 8;

template <int V, bool F16 = false>  // V float4 per lane: D = 128 * V; F16: y stored as fp16 (training decoder input)
__global__ void __launch_bounds__(LN_WARPS * 32) layernorm_fwd_kernel(const float* __restrict__ x, long long ldx,
                                                                      const float* __restrict__ w,
                                                                      const float* __restrict__ b,
                                                                      __nv_bfloat16* __restrict__ y, long long ldy,
                                                                      float* __restrict__ mean_out,
                                                                      float* __restrict__ rstd_out, int M, float eps) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  const int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  constexpr int D = 128 * V;
  const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * ldx);
  float4 v[V];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    v[i] = xr[i * 32 + lane];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float a = v[i].x - mean, c = v[i].y - mean, d = v[i].z - mean, e = v[i].w - mean;
    q += (a * a + c * c) + (d * d + e * e);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + eps);
  if (lane == 0 && mean_out) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
  const float4* wr = reinterpret_cast<const float4*>(w);
  const float4* br = reinterpret_cast<const float4*>(b);
  uint2* yr = reinterpret_cast<uint2*>(y + (long long)row * ldy);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float4 ww = __ldg(wr + i * 32 + lane), bb = __ldg(br + i * 32 + lane);
    uint2 o;
    o.x = pack16x2<F16>((v[i].x - mean) * rstd * ww.x + bb.x, (v[i].y - mean) * rstd * ww.y + bb.y);
    o.y = pack16x2<F16>((v[i].z - mean) * rstd * ww.z + bb.z, (v[i].w - mean) * rstd * ww.w + bb.w);
    yr[i * 32 + lane] = o;
  }
}

// Streaming variant of the forward: persistent CTAs (one per SM, LNS_WARPS warps), every warp walks rows w, w + W, ... and
// keeps the NEXT row in flight as cp.async copies into its own shared-memory buffer while it normalises the current one.
// The register-resident kernel above is two discrete latency-bound waves at M = 5264 (658 CTAs on 444 slots: 12 us for
// 48 MB); total on-chip storage cannot hold all rows at once, so the loads have to be pipelined against the math instead.
constexpr int LNS_WARPS = 16;
__device__ __forceinline__ void ln_cp_async_16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
template <int V, bool F16>
__global__ void __launch_bounds__(LNS_WARPS * 32, 1) layernorm_fwd_stream_kernel(const float* __restrict__ x, long long ldx,
                                                                                const float* __restrict__ w,
                                                                                const float* __restrict__ b,
                                                                                __nv_bfloat16* __restrict__ y, long long ldy,
                                                                                float* __restrict__ mean_out,
                                                                                float* __restrict__ rstd_out, int M, float eps) {
  extern __shared__ __align__(16) uint8_t lns_smem[];
  constexpr int D = 128 * V;
  float4* sw = reinterpret_cast<float4*>(lns_smem);           // [D / 4] weight
  float4* sb = sw + D / 4;                                      // [D / 4] bias
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4* rowbuf = sb + D / 4 + warp * (2 * D / 4);            // this warp's two row buffers
  // parameters do not depend on the previous kernel: stage them before the PDL wait
  for (int i = threadIdx.x; i < D / 4; i += LNS_WARPS * 32) {
    sw[i] = __ldg(reinterpret_cast<const float4*>(w) + i);
    sb[i] = __ldg(reinterpret_cast<const float4*>(b) + i);
  }
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  __syncthreads();
  const int W = gridDim.x * LNS_WARPS;
  int row = blockIdx.x * LNS_WARPS + warp;
  auto fetch = [&](int r, int buf) {
    const float4* src = reinterpret_cast<const float4*>(x + (long long)r * ldx);
    const uint32_t dst = smem_u32(rowbuf + buf * (D / 4));
#pragma unroll
    for (int i = 0; i < V; ++i) ln_cp_async_16(dst + (i * 32 + lane) * 16, src + i * 32 + lane);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (row < M) fetch(row, 0);
  int buf = 0;
  for (; row < M; row += W, buf ^= 1) {
    if (row + W < M) {
      fetch(row + W, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    // every lane reads back exactly the 16-byte chunks it copied itself: no warp synchronisation needed
    const float4* xr = rowbuf + buf * (D / 4);
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      v[i] = xr[i * 32 + lane];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) * (1.f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float a = v[i].x - mean, c = v[i].y - mean, d = v[i].z - mean, e = v[i].w - mean;
      q += (a * a + c * c) + (d * d + e * e);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + eps);
    if (lane == 0 && mean_out) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
    uint2* yr = reinterpret_cast<uint2*>(y + (long long)row * ldy);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float4 ww = sw[i * 32 + lane], bb = sb[i * 32 + lane];
      uint2 o;
      o.x = pack16x2<F16>((v[i].x - mean) * rstd * ww.x + bb.x, (v[i].y - mean) * rstd * ww.y + bb.y);
      o.y = pack16x2<F16>((v[i].z - mean) * rstd * ww.z + bb.z, (v[i].w - mean) * rstd * ww.w + bb.w);
      yr[i * 32 + lane] = o;
    }
  }
}

// dx = dres + rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * w,  xhat = (x - mean) * rstd
template <int V, bool DY_F32>
__global__ void __launch_bounds__(LN_WARPS * 32) layernorm_bwd_kernel(const float* __restrict__ x, long long ldx,
                                                                      const float* __restrict__ w,
                                                                      const void* __restrict__ dy, long long lddy,
                                                                      const float* __restrict__ dres, long long lddres,
                                                                      float* __restrict__ dx, long long lddx,
                                                                      __nv_bfloat16* __restrict__ dx_bf16,
                                                                      long long lddxb, int M, float eps) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  const int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  constexpr int D = 128 * V;
  const float4* xr = reinterpret_cast<const float4*>(x + (long long)row * ldx);
  float4 v[V], g[V];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    v[i] = xr[i * 32 + lane];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / D) + eps);
  const float4* wr = reinterpret_cast<const float4*>(w);
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    float4 d;
    if (DY_F32) {
      d = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dy) + (long long)row * lddy)[i * 32 + lane];
    } else {
      const uint2 u = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(dy) + (long long)row * lddy)[i * 32 + lane];
      const float2 a = unpack_bf16x2(u.x), c = unpack_bf16x2(u.y);
      d = make_float4(a.x, a.y, c.x, c.y);
    }
    const float4 ww = __ldg(wr + i * 32 + lane);
    g[i] = make_float4(d.x * ww.x, d.y * ww.y, d.z * ww.z, d.w * ww.w);
    v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;  // xhat
    sg += (g[i].x + g[i].y) + (g[i].z + g[i].w);
    sgx += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
  }
  const float mg = warp_sum(sg) * (1.f / D);
  const float mgx = warp_sum(sgx) * (1.f / D);
  float4* dxr = reinterpret_cast<float4*>(dx + (long long)row * lddx);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (dres) r = reinterpret_cast<const float4*>(dres + (long long)row * lddres)[i * 32 + lane];
    float4 o;
    o.x = r.x + rstd * (g[i].x - mg - v[i].x * mgx);
    o.y = r.y + rstd * (g[i].y - mg - v[i].y * mgx);
    o.z = r.z + rstd * (g[i].z - mg - v[i].z * mgx);
    o.w = r.w + rstd * (g[i].w - mg - v[i].w * mgx);
    dxr[i * 32 + lane] = o;
    if (dx_bf16) {
      uint2 u;
      u.x = pack_bf16x2(o.x, o.y);
      u.y = pack_bf16x2(o.z, o.w);
      reinterpret_cast<uint2*>(dx_bf16 + (long long)row * lddxb)[i * 32 + lane] = u;
    }
  }
}

}  // namespace mv

#define MV_LN_DISPATCH(V_, CALL) \
  switch (V_) {                  \
    case 1: { constexpr int V = 1; CALL; } break;   \
    case 2: { constexpr int V = 2; CALL; } break;   \
    case 3: { constexpr int V = 3; CALL; } break;   \
    case 4: { constexpr int V = 4; CALL; } break;   \
    case 6: { constexpr int V = 6; CALL; } break;   \
    case 8: { constexpr int V = 8; CALL; } break;   \
    case 12: { constexpr int V = 12; CALL; } break; \
    case 16: { constexpr int V = 16; CALL; } break; \
    default: mv::set_error("layernorm: D=%d unsupported (need D/128 in {1,2,3,4,6,8,12,16})", d); return MV_ERR_ARG; \
  }

extern "C" int mv_layernorm_fwd(const float* x, int64_t ldx, const float* w, const float* b, void* y, int64_t ldy,
                                int y_f16, float* mean, float* rstd, int m, int d, float eps, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(x && w && b && y && m > 0, "mv_layernorm_fwd: null/empty");
  MV_CHECK_ARG(d % 128 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "mv_layernorm_fwd: D %% 128, ldx %% 4, ldy %% 4");
  MV_CHECK_ARG((mean == nullptr) == (rstd == nullptr), "mv_layernorm_fwd: mean and rstd go together");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  // streaming variant from two rows per streaming warp on (CUDA graph of 40 launches over rotating buffers, D = 1536:
  // M = 5264 10.9 -> 9.7 us, M = 10528 17.5 -> 16.7 us; M = 2632 4.9 -> 5.7 us, so not below that);
  // MV_LN_STREAM=0 disables it, =2 forces it for every M
  static const int stream_env = [] { const char* e = getenv("MV_LN_STREAM"); return e ? atoi(e) : 1; }();
  const int sms = device_sms() > 0 ? device_sms() : 148;
  const int smem_s = (2 + 2 * LNS_WARPS) * d * 4;
  if (!y_f16 && stream_env != 0 && smem_s <= 220 * 1024 && ldx % 4 == 0 &&
      (stream_env == 2 || m >= 2 * sms * LNS_WARPS)) {
    int grid_s = (m + LNS_WARPS - 1) / LNS_WARPS;
    if (grid_s > sms) grid_s = sms;
    MV_LN_DISPATCH(d / 128, {
      static std::atomic<uint64_t> attr_s{0};  // one bit per device; one flag per instantiation (the macro repeats this block)
      if (first_use_on_device(attr_s)) {
        cudaError_t e_ = cudaFuncSetAttribute(layernorm_fwd_stream_kernel<V, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e_ != cudaSuccess) { set_error("cudaFuncSetAttribute(layernorm_fwd_stream): %s", cudaGetErrorString(e_)); return (int)e_; }
      }
      MV_LAUNCH((layernorm_fwd_stream_kernel<V, false>), grid_s, LNS_WARPS * 32, smem_s, stream, x, ldx, w, b,
                reinterpret_cast<__nv_bfloat16*>(y), ldy, mean, rstd, m, eps);
    });
    MV_CHECK_LAUNCH("layernorm_fwd_stream");
    return MV_OK;
  }
  const int grid = (m + LN_WARPS - 1) / LN_WARPS;
  if (y_f16) {
    MV_LN_DISPATCH(d / 128, (MV_LAUNCH((layernorm_fwd_kernel<V, true>), grid, LN_WARPS * 32, 0, stream,
                                x, ldx, w, b, reinterpret_cast<__nv_bfloat16*>(y), ldy, mean, rstd, m, eps)));
  } else {
    MV_LN_DISPATCH(d / 128, (MV_LAUNCH((layernorm_fwd_kernel<V>), grid, LN_WARPS * 32, 0, stream,
                                x, ldx, w, b, reinterpret_cast<__nv_bfloat16*>(y), ldy, mean, rstd, m, eps)));
  }
  MV_CHECK_LAUNCH("layernorm_fwd");
  return MV_OK;
}

extern "C" int mv_layernorm_bwd(const float* x, int64_t ldx, const float* w, const void* dy, int64_t lddy, int dy_f32,
                                const float* dres, int64_t lddres, float* dx, int64_t lddx, void* dx_bf16,
                                int64_t lddxb, int m, int d, float eps, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(x && w && dy && dx && m > 0, "mv_layernorm_bwd: null/empty");
  MV_CHECK_ARG(d % 128 == 0 && ldx % 4 == 0 && lddy % 4 == 0 && lddx % 4 == 0 && lddres % 4 == 0 && lddxb % 4 == 0,
               "mv_layernorm_bwd: alignment");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const int grid = (m + LN_WARPS - 1) / LN_WARPS;
  __nv_bfloat16* dxb = reinterpret_cast<__nv_bfloat16*>(dx_bf16);
  if (dy_f32) {
    MV_LN_DISPATCH(d / 128, (MV_LAUNCH((layernorm_bwd_kernel<V, true>), grid, LN_WARPS * 32, 0, stream, 
                                x, ldx, w, dy, lddy, dres, lddres, dx, lddx, dxb, lddxb, m, eps)));
  } else {
    MV_LN_DISPATCH(d / 128, (MV_LAUNCH((layernorm_bwd_kernel<V, false>), grid, LN_WARPS * 32, 0, stream, 
                                x, ldx, w, dy, lddy, dres, lddres, dx, lddx, dxb, lddxb, m, eps)));
  }
  MV_CHECK_LAUNCH("layernorm_bwd");
  return MV_OK;
}
