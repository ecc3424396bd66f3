// Memory-bound glue kernels around the GEMMs (all HBM-bound: coalesced, 16-byte vectors where the layout allows).
//
//   mv_prep_input        x fp32 NCHW [B,3,S,S] -> NHWC bf16 image padded to 8 channels (D0 of ConvStream,
//                        mipheivit.py:66-73) and the patch matrix [B*g*g, 592] of timm PatchEmbed (conv k14 s14 as a
//                        GEMM: K = (c, ky, kx) = 588, zero-padded to 592 for 16-byte rows)
//   mv_tokens_to_map     Encoder.forward tail, mipheivit.py:158-162: drop the prefix tokens, view as a g x g map,
//                        F.interpolate(scale_factor=(t/g, t/g), mode="bicubic") -> NHWC bf16 [B,t,t,D]
//   mv_upsample2x        Fusion_Block's F.interpolate(scale_factor=2, bilinear, align_corners=False), NHWC bf16
#include "mv_host.h"
#include "mv_ptx.cuh"

namespace mv {

// Input pixel (b, c, y, x): fp32 NCHW (already normalised, the reference's DataLoader output) or raw uint8 NHWC H&E tiles
// normalised here with the H-Optimus statistics of src/dataset.py:600-601, v = u8 * scale[c] + bias[c].
struct InNorm {
  float scale[3], bias[3];
};
template <bool U8>
__device__ __forceinline__ float in_px(const void* x, const InNorm& nm, long long b, int c, int y, int xx, int S) {
  if (U8) {
    const uint8_t u = reinterpret_cast<const uint8_t*>(x)[((b * S + y) * (long long)S + xx) * 3 + c];
    return (float)u * nm.scale[c] + nm.bias[c];
  }
  return reinterpret_cast<const float*>(x)[((b * 3 + c) * S + y) * (long long)S + xx];
}

template <bool U8, bool F16>
__global__ void prep_image_kernel(const void* __restrict__ x, InNorm nm, __nv_bfloat16* __restrict__ img, int B, int S) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  const long long npix = (long long)B * S * S;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const long long plane = (long long)S * S;
  const long long b = i / plane, r = i - b * plane;
  const int y = (int)(r / S), xx = (int)(r - (long long)y * S);
  uint4 o;
  o.x = pack16x2<F16>(in_px<U8>(x, nm, b, 0, y, xx, S), in_px<U8>(x, nm, b, 1, y, xx, S));
  o.y = pack16x2<F16>(in_px<U8>(x, nm, b, 2, y, xx, S), 0.f);
  o.z = 0u;
  o.w = 0u;
  reinterpret_cast<uint4*>(img)[i] = o;
}

template <bool U8>
__global__ void patch_matrix_kernel(const void* __restrict__ x, InNorm nm, __nv_bfloat16* __restrict__ a, int B, int S,
                                    int g, int ldk) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  // one thread per (patch row, 8-wide k group); ldk = 592
  const int groups = ldk / 8;
  const long long total = (long long)B * g * g * groups;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int kg = (int)(i % groups);
  const long long row = i / groups;
  const int px = (int)(row % g), py = (int)((row / g) % g);
  const long long b = row / ((long long)g * g);
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = kg * 8 + j;
    if (k < 588) {
      const int c = k / 196, rr = k - c * 196, ky = rr / 14, kx = rr - ky * 14;
      f[j] = in_px<U8>(x, nm, b, c, py * 14 + ky, px * 14 + kx, S);
    } else {
      f[j] = 0.f;
    }
  }
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]);
  o.w = pack_bf16x2(f[6], f[7]);
  reinterpret_cast<uint4*>(a + row * ldk)[kg] = o;
}

// PyTorch upsample_bicubic2d coefficients (A = -0.75), align_corners = False, explicit scale factor
__device__ __forceinline__ void cubic_coeffs(float t, float* w) {
  const float A = -0.75f;
  const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  w[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  w[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  w[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  w[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

// tokens bf16 [B, ntok, ldt] (prefix tokens first) -> out NHWC bf16 [B, t, t, D]; one block per output pixel
template <bool F16>
__global__ void tokens_to_map_kernel(const __nv_bfloat16* __restrict__ tok, long long ldt, int ntok, int prefix, int g,
                                     int t, int D, float inv_scale, __nv_bfloat16* __restrict__ out) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  const int ox = blockIdx.x % t, oy = (blockIdx.x / t) % t, b = blockIdx.x / (t * t);
  const float sy = (oy + 0.5f) * inv_scale - 0.5f, sx = (ox + 0.5f) * inv_scale - 0.5f;
  const int iy = (int)floorf(sy), ix = (int)floorf(sx);
  float wy[4], wx[4];
  cubic_coeffs(sy - iy, wy);
  cubic_coeffs(sx - ix, wx);
  const __nv_bfloat16* base = tok + ((long long)b * ntok + prefix) * ldt;
  for (int c = threadIdx.x * 8; c < D; c += blockDim.x * 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = min(max(iy - 1 + a, 0), g - 1);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int xx = min(max(ix - 1 + e, 0), g - 1);
        const uint4 u = *reinterpret_cast<const uint4*>(base + (long long)(yy * g + xx) * ldt + c);
        const float w = wy[a] * wx[e];
        const float2 p0 = unpack16x2<F16>(u.x), p1 = unpack16x2<F16>(u.y), p2 = unpack16x2<F16>(u.z), p3 = unpack16x2<F16>(u.w);
        acc[0] += w * p0.x; acc[1] += w * p0.y; acc[2] += w * p1.x; acc[3] += w * p1.y;
        acc[4] += w * p2.x; acc[5] += w * p2.y; acc[6] += w * p3.x; acc[7] += w * p3.y;
      }
    }
    uint4 o;
    o.x = pack16x2<F16>(acc[0], acc[1]);
    o.y = pack16x2<F16>(acc[2], acc[3]);
    o.z = pack16x2<F16>(acc[4], acc[5]);
    o.w = pack16x2<F16>(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(out + (((long long)b * t + oy) * t + ox) * D + c) = o;
  }
}

// NHWC bf16 [B,h,w,C] -> [B,2h,2w,C], bilinear, align_corners = False (Fusion_Block, mipheivit.py:88-93).
// One thread per (INPUT pixel, 8 channels) writes the 2x2 output block it centres: with clamped neighbours m = max(i-1, 0),
// p = min(i+1, n-1) the exact x2 weights are  out(2i) = 0.25 in(m) + 0.75 in(i),  out(2i+1) = 0.75 in(i) + 0.25 in(p)  in both
// directions (separable) — 9 sixteen-byte loads and one index decode per 4 outputs instead of 16 loads and 4 decodes.
template <bool F16>
__global__ void upsample2x_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int h,
                                  int w, int C) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  const int cg = C / 8;
  const long long total = (long long)B * h * w * cg;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = (int)(i % cg);
  long long r = i / cg;
  const int x = (int)(r % w);
  r /= w;
  const int y = (int)(r % h);
  const long long b = r / h;
  const int ym = max(y - 1, 0), yp = min(y + 1, h - 1), xm = max(x - 1, 0), xp = min(x + 1, w - 1);
  const __nv_bfloat16* ib = in + b * (long long)h * w * C + c8 * 8;
  const int ys[3] = {ym, y, yp}, xs[3] = {xm, x, xp};
  float v[3][3][8];  // [row][col][channel]
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const uint4 u = *reinterpret_cast<const uint4*>(ib + ((long long)ys[a] * w + xs[q]) * C);
      const uint32_t* pu = &u.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack16x2<F16>(pu[j]);
        v[a][q][2 * j] = f.x;
        v[a][q][2 * j + 1] = f.y;
      }
    }
  __nv_bfloat16* ob = out + (b * 4 * h * w + (long long)(2 * y) * (2 * w) + 2 * x) * C + c8 * 8;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy) {
    float t[3][8];  // rows combined: 3 columns x 8 channels
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j)
        t[q][j] = dy == 0 ? 0.25f * v[0][q][j] + 0.75f * v[1][q][j] : 0.75f * v[1][q][j] + 0.25f * v[2][q][j];
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = dx == 0 ? 0.25f * t[0][j] + 0.75f * t[1][j] : 0.75f * t[1][j] + 0.25f * t[2][j];
      uint4 u;
      u.x = pack16x2<F16>(o[0], o[1]);
      u.y = pack16x2<F16>(o[2], o[3]);
      u.z = pack16x2<F16>(o[4], o[5]);
      u.w = pack16x2<F16>(o[6], o[7]);
      *reinterpret_cast<uint4*>(ob + ((long long)dy * (2 * w) + dx) * C) = u;
    }
  }
}

// Adjoint of tokens_to_map_kernel: d_tok[b, prefix + gy*g + gx, :] = sum_{oy,ox} Wy[oy,gy] Wx[ox,gx] d_map[b,oy,ox,:];
// prefix-token rows get zero. One block per token row.
__global__ void tokens_to_map_bwd_kernel(const __nv_bfloat16* __restrict__ dmap, int ntok, int prefix, int g, int t, int D,
                                         float inv_scale, __nv_bfloat16* __restrict__ dtok, long long ldt) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  __shared__ float wy[64], wx[64];
  const int tok = blockIdx.x % ntok, b = blockIdx.x / ntok;
  __nv_bfloat16* dst = dtok + ((long long)b * ntok + tok) * ldt;
  if (tok < prefix) {
    for (int c = threadIdx.x * 8; c < D; c += blockDim.x * 8) *reinterpret_cast<uint4*>(dst + c) = make_uint4(0, 0, 0, 0);
    return;
  }
  const int gy = (tok - prefix) / g, gx = (tok - prefix) % g;
  if (threadIdx.x < 2 * t) {  // weight of output coordinate o on source coordinate gy (threads 0..t-1) / gx (t..2t-1)
    const int o = threadIdx.x % t;
    const int gsel = threadIdx.x < t ? gy : gx;
    const float sc = (o + 0.5f) * inv_scale - 0.5f;
    const int i0 = (int)floorf(sc);
    float w[4];
    cubic_coeffs(sc - i0, w);
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a)
      if (min(max(i0 - 1 + a, 0), g - 1) == gsel) acc += w[a];
    (threadIdx.x < t ? wy : wx)[o] = acc;
  }
  __syncthreads();
  for (int c = threadIdx.x * 8; c < D; c += blockDim.x * 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int oy = 0; oy < t; ++oy) {
      if (wy[oy] == 0.f) continue;
      for (int ox = 0; ox < t; ++ox) {
        const float w = wy[oy] * wx[ox];
        if (w == 0.f) continue;
        const uint4 u = *reinterpret_cast<const uint4*>(dmap + (((long long)b * t + oy) * t + ox) * D + c);
        const float2 p0 = unpack_bf16x2(u.x), p1 = unpack_bf16x2(u.y), p2 = unpack_bf16x2(u.z), p3 = unpack_bf16x2(u.w);
        acc[0] += w * p0.x; acc[1] += w * p0.y; acc[2] += w * p1.x; acc[3] += w * p1.y;
        acc[4] += w * p2.x; acc[5] += w * p2.y; acc[6] += w * p3.x; acc[7] += w * p3.y;
      }
    }
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]);
    o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]);
    o.w = pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(dst + c) = o;
  }
}

// residual stream rows of the prefix tokens: x[b, j, :] = prefix[j, :] (cls + register tokens, no pos-embed:
// timm _pos_embed with no_embed_class=True)
__global__ void fill_prefix_kernel(float* __restrict__ x, long long ldx, const float* __restrict__ prefix, int B, int ntok,
                                   int nprefix, int D) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  const int d4 = D / 4;
  const long long total = (long long)B * nprefix * d4;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % d4);
  const int j = (int)((i / d4) % nprefix);
  const long long b = i / ((long long)d4 * nprefix);
  reinterpret_cast<float4*>(x + (b * ntok + j) * ldx)[c] = reinterpret_cast<const float4*>(prefix + (long long)j * D)[c];
}

}  // namespace mv

extern "C" int mv_fill_prefix(float* x, int64_t ldx, const float* prefix, int batch, int n_tok, int n_prefix, int dim,
                              void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(x && prefix && batch > 0 && dim % 4 == 0 && ldx % 4 == 0, "mv_fill_prefix: shape");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const long long total = (long long)batch * n_prefix * (dim / 4);
  MV_LAUNCH(fill_prefix_kernel, (unsigned)((total + 255) / 256), 256, 0, stream, x, ldx, prefix, batch, n_tok, n_prefix, dim);
  MV_CHECK_LAUNCH("fill_prefix");
  return MV_OK;
}

namespace mv {
template <bool U8>
static int launch_prep(const void* x, const InNorm& nm, void* img_nhwc8, int img_f16, void* patch_matrix, int batch,
                       int size, int ldk, cudaStream_t stream) {
  if (img_nhwc8) {
    const long long npix = (long long)batch * size * size;
    if (img_f16)
      MV_LAUNCH((prep_image_kernel<U8, true>), (unsigned)((npix + 255) / 256), 256, 0, stream, x, nm,
                reinterpret_cast<__nv_bfloat16*>(img_nhwc8), batch, size);
    else
      MV_LAUNCH((prep_image_kernel<U8, false>), (unsigned)((npix + 255) / 256), 256, 0, stream, x, nm,
                reinterpret_cast<__nv_bfloat16*>(img_nhwc8), batch, size);
    MV_CHECK_LAUNCH("prep_image");
  }
  if (patch_matrix) {
    const int g = size / 14;
    const long long total = (long long)batch * g * g * (ldk / 8);
    MV_LAUNCH(patch_matrix_kernel<U8>, (unsigned)((total + 255) / 256), 256, 0, stream, x, nm,
              reinterpret_cast<__nv_bfloat16*>(patch_matrix), batch, size, g, ldk);
    MV_CHECK_LAUNCH("patch_matrix");
  }
  return MV_OK;
}
}  // namespace mv

extern "C" int mv_prep_input(const float* x, void* img_nhwc8, int img_f16, void* patch_matrix, int batch, int size,
                             int ldk, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(x && batch > 0 && size >= 14, "mv_prep_input: null/empty");
  MV_CHECK_ARG(!patch_matrix || ldk == 592, "mv_prep_input: patch matrix pitch must be 592");
  InNorm nm = {{1.f, 1.f, 1.f}, {0.f, 0.f, 0.f}};
  return launch_prep<false>(x, nm, img_nhwc8, img_f16, patch_matrix, batch, size, ldk, reinterpret_cast<cudaStream_t>(stream_));
}

// raw uint8 H&E tiles, NHWC [B, S, S, 3]; v = u8 * scale[c] + bias[c] (scale = 1 / (255 std_c), bias = -mean_c / std_c)
extern "C" int mv_prep_input_u8(const void* tiles_u8, const float* scale3, const float* bias3, void* img_nhwc8,
                                int img_f16, void* patch_matrix, int batch, int size, int ldk, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(tiles_u8 && scale3 && bias3 && batch > 0 && size >= 14, "mv_prep_input_u8: null/empty");
  MV_CHECK_ARG(!patch_matrix || ldk == 592, "mv_prep_input_u8: patch matrix pitch must be 592");
  InNorm nm;
  for (int c = 0; c < 3; ++c) {
    nm.scale[c] = scale3[c];  // HOST pointers: six constants passed by value to the kernels
    nm.bias[c] = bias3[c];
  }
  return launch_prep<true>(tiles_u8, nm, img_nhwc8, img_f16, patch_matrix, batch, size, ldk,
                           reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int mv_tokens_to_map(const void* tokens, int64_t ldt, void* out, int batch, int n_tok, int prefix, int grid,
                                int target, int dim, int f16, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(tokens && out && batch > 0, "mv_tokens_to_map: null/empty");
  MV_CHECK_ARG(n_tok == prefix + grid * grid && dim % 8 == 0 && ldt % 8 == 0, "mv_tokens_to_map: shape");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  // PyTorch uses the reciprocal of the user-supplied scale factor (target/grid) for the source coordinates
  const float inv_scale = (float)(1.0 / ((double)target / (double)grid));
  const int threads = dim / 8 < 256 ? (dim / 8 + 31) / 32 * 32 : 256;
  if (f16)
    MV_LAUNCH(tokens_to_map_kernel<true>, batch * target * target, threads, 0, stream,
              reinterpret_cast<const __nv_bfloat16*>(tokens), ldt, n_tok, prefix, grid, target, dim, inv_scale,
              reinterpret_cast<__nv_bfloat16*>(out));
  else
    MV_LAUNCH(tokens_to_map_kernel<false>, batch * target * target, threads, 0, stream,
              reinterpret_cast<const __nv_bfloat16*>(tokens), ldt, n_tok, prefix, grid, target, dim, inv_scale,
              reinterpret_cast<__nv_bfloat16*>(out));
  MV_CHECK_LAUNCH("tokens_to_map");
  return MV_OK;
}

extern "C" int mv_tokens_to_map_bwd(const void* dmap, void* dtokens, int64_t ldt, int batch, int n_tok, int prefix,
                                    int grid, int target, int dim, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(dmap && dtokens && batch > 0, "mv_tokens_to_map_bwd: null/empty");
  MV_CHECK_ARG(n_tok == prefix + grid * grid && dim % 8 == 0 && ldt % 8 == 0 && target <= 64, "mv_tokens_to_map_bwd: shape");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const float inv_scale = (float)(1.0 / ((double)target / (double)grid));
  int threads = dim / 8 < 256 ? (dim / 8 + 31) / 32 * 32 : 256;
  if (threads < 2 * target) threads = (2 * target + 31) / 32 * 32;
  MV_LAUNCH(tokens_to_map_bwd_kernel, batch * n_tok, threads, 0, stream, reinterpret_cast<const __nv_bfloat16*>(dmap), n_tok, prefix,
                                                                 grid, target, dim, inv_scale,
                                                                 reinterpret_cast<__nv_bfloat16*>(dtokens), ldt);
  MV_CHECK_LAUNCH("tokens_to_map_bwd");
  return MV_OK;
}

extern "C" int mv_upsample2x(const void* in, void* out, int batch, int h, int w, int c, int f16, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(in && out && batch > 0 && h > 0 && w > 0 && c % 8 == 0, "mv_upsample2x: shape");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const long long total = (long long)batch * h * w * (c / 8);
  if (f16)
    MV_LAUNCH(upsample2x_kernel<true>, (unsigned)((total + 255) / 256), 256, 0, stream,
              reinterpret_cast<const __nv_bfloat16*>(in), reinterpret_cast<__nv_bfloat16*>(out), batch, h, w, c);
  else
    MV_LAUNCH(upsample2x_kernel<false>, (unsigned)((total + 255) / 256), 256, 0, stream,
              reinterpret_cast<const __nv_bfloat16*>(in), reinterpret_cast<__nv_bfloat16*>(out), batch, h, w, c);
  MV_CHECK_LAUNCH("upsample2x");
  return MV_OK;
}
