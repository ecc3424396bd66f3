// Memory-bound kernels of the decoder backward pass (NHWC bf16 feature maps).
//
//   mv_transpose_bf16       [M, C] -> [C(+ones row), M]: K-major copy of an activation / gradient for the weight-gradient
//                           GEMMs (MV_GEMM_NN_ATOMIC needs the reduction index contiguous in A)
//   mv_upsample2x_bwd       adjoint of Fusion_Block's bilinear x2 (mipheivit.py:89)
//   mv_zero_insert2x        dgrad of a stride-2 conv = stride-1 conv over the zero-interleaved gradient
//   mv_add_bf16             skip-connection gradient sums
//   mv_heads_ds             ds = dpred * (1 - pred^2) (tanh'), NCHW fp32 -> NHWC bf16, and the head conv bias gradients
//   mv_heads_bwd_stencil    gradients of the gated 3x3 head conv w.r.t. the gates and the per-tap projections
//                           (SegmentationHead / AttentionBlock, src/generators/unet.py:407-438)
#include "mv_host.h"
#include "mv_ptx.cuh"

namespace mv {

// in [M, C] (pitch ldi) -> out [R, ldo] with out[c, m] = in[m, c]; optional extra row C filled with ones.
template <bool IN_F16>
__global__ void transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, long long ldi, __nv_bfloat16* __restrict__ out,
                                      long long ldo, long long M, int C, int ones_row) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  __shared__ __nv_bfloat16 tile[64][66];
  const long long m0 = (long long)blockIdx.x * 64;
  const int c0 = blockIdx.y * 64;
  for (int i = threadIdx.y; i < 64; i += blockDim.y) {
    const long long m = m0 + i;
    const int c = c0 + threadIdx.x * 2;
    __nv_bfloat16 a = __float2bfloat16(0.f), b = a;
    if (m < M && c < C) {
      if (IN_F16) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(in + m * ldi + c));
        a = __float2bfloat16(f.x); b = __float2bfloat16(f.y);
      } else {
        const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(in + m * ldi + c);
        a = v.x; b = v.y;
      }
    }
    tile[i][threadIdx.x * 2] = a;
    tile[i][threadIdx.x * 2 + 1] = b;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 64; i += blockDim.y) {
    const int c = c0 + i;
    const long long m = m0 + threadIdx.x * 2;
    if (c < C && m < M) {
      __nv_bfloat162 v;
      v.x = tile[threadIdx.x * 2][i];
      v.y = tile[threadIdx.x * 2 + 1][i];
      if (m + 1 < M) *reinterpret_cast<__nv_bfloat162*>(out + (long long)c * ldo + m) = v;
      else out[(long long)c * ldo + m] = v.x;
    }
  }
  if (ones_row && blockIdx.y == 0) {
    const __nv_bfloat16 one = __float2bfloat16(1.f);
    for (int i = threadIdx.y * 32 + threadIdx.x; i < 64; i += 32 * blockDim.y)
      if (m0 + i < M) out[(long long)C * ldo + m0 + i] = one;
  }
}

// Vectorised form for 16-byte-aligned operands with C % 8 == 0: a thread loads an 8-row x 8-channel block as eight 16-byte
// vectors, transposes it in registers (byte permutes) and writes eight 16-byte vectors, one per channel, of 8 consecutive
// rows — no shared memory, every access a full sector. (The 2-byte-granular kernel above reached 0.35 of the HBM peak.)
template <bool IN_F16>
__global__ void __launch_bounds__(256) transpose_bf16_vec_kernel(const __nv_bfloat16* __restrict__ in, long long ldi,
                                                                  __nv_bfloat16* __restrict__ out, long long ldo, long long M,
                                                                  int C, int ones_row) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  const int ncg = C / 8;                       // 8-channel groups
  const int per_blk = 256 / ncg > 0 ? 256 / ncg : 1;  // 8-row groups per block (ncg <= 256)
  const int cg = threadIdx.x % ncg, rg = threadIdx.x / ncg;
  if (rg >= per_blk) return;
  const long long r0 = ((long long)blockIdx.x * per_blk + rg) * 8;
  if (r0 >= M) return;
  uint32_t w[8][4];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    uint4 u = make_uint4(0, 0, 0, 0);
    if (r0 + k < M) u = *reinterpret_cast<const uint4*>(in + (r0 + k) * ldi + cg * 8);
    if (IN_F16) {
      const float2 a = unpack16x2<true>(u.x), b = unpack16x2<true>(u.y), c = unpack16x2<true>(u.z), d = unpack16x2<true>(u.w);
      u = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(b.x, b.y), pack_bf16x2(c.x, c.y), pack_bf16x2(d.x, d.y));
    }
    w[k][0] = u.x; w[k][1] = u.y; w[k][2] = u.z; w[k][3] = u.w;
  }
  const bool full = r0 + 8 <= M;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    // channel e of rows (2q, 2q + 1): low or high half of word e / 2 of each row
    const uint32_t sel = (e & 1) ? 0x7632u : 0x5410u;
    uint4 o;
    o.x = __byte_perm(w[0][e >> 1], w[1][e >> 1], sel);
    o.y = __byte_perm(w[2][e >> 1], w[3][e >> 1], sel);
    o.z = __byte_perm(w[4][e >> 1], w[5][e >> 1], sel);
    o.w = __byte_perm(w[6][e >> 1], w[7][e >> 1], sel);
    __nv_bfloat16* dst = out + (long long)(cg * 8 + e) * ldo + r0;
    if (full) {
      *reinterpret_cast<uint4*>(dst) = o;
    } else {
      const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(&o);
      for (int k = 0; k < 8 && r0 + k < M; ++k) dst[k] = h[k];
    }
  }
  if (ones_row && cg == 0) {
    const __nv_bfloat16 one = __float2bfloat16(1.f);
    __nv_bfloat16* dst = out + (long long)C * ldo + r0;
    for (int k = 0; k < 8 && r0 + k < M; ++k) dst[k] = one;
  }
}

// dup [B, 2h, 2w, C] (pitch ldu per pixel) -> dx [B, h, w, C]; one thread per (input pixel, 8 channels)
__global__ void upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ dup, long long ldu, __nv_bfloat16* __restrict__ dx,
                                      int B, int h, int w, int C) {
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
  // adjoint of the x2 bilinear weights (out(2i) = .25 in(i-1) + .75 in(i), out(2i+1) = .75 in(i) + .25 in(i+1), neighbours
  // clamped): input i gathers outputs 2i-1 .. 2i+2 with weights (.25, .75, .75, .25); at the borders the clamped
  // neighbour folds its .25 into the .75 and the out-of-range taps vanish. Separable: 4 x 4 taps, no weight search.
  const float wy[4] = {y >= 1 ? 0.25f : 0.f, y == 0 ? 1.f : 0.75f, y == h - 1 ? 1.f : 0.75f, y <= h - 2 ? 0.25f : 0.f};
  const float wx[4] = {x >= 1 ? 0.25f : 0.f, x == 0 ? 1.f : 0.75f, x == w - 1 ? 1.f : 0.75f, x <= w - 2 ? 0.25f : 0.f};
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int oy = min(max(2 * y - 1 + a, 0), 2 * h - 1);  // clamped rows / columns carry weight 0
    float row[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) row[j] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int ox = min(max(2 * x - 1 + q, 0), 2 * w - 1);
      const uint4 u = *reinterpret_cast<const uint4*>(dup + ((b * 2 * h + oy) * 2 * w + ox) * ldu + c8 * 8);
      const float2 p0 = unpack_bf16x2(u.x), p1 = unpack_bf16x2(u.y), p2 = unpack_bf16x2(u.z), p3 = unpack_bf16x2(u.w);
      row[0] += wx[q] * p0.x; row[1] += wx[q] * p0.y; row[2] += wx[q] * p1.x; row[3] += wx[q] * p1.y;
      row[4] += wx[q] * p2.x; row[5] += wx[q] * p2.y; row[6] += wx[q] * p3.x; row[7] += wx[q] * p3.y;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += wy[a] * row[j];
  }
  uint4 o;
  o.x = pack_bf16x2(acc[0], acc[1]);
  o.y = pack_bf16x2(acc[2], acc[3]);
  o.z = pack_bf16x2(acc[4], acc[5]);
  o.w = pack_bf16x2(acc[6], acc[7]);
  reinterpret_cast<uint4*>(dx)[i] = o;
}

// u [B, 2h, 2w, C]: u[2y, 2x] = dz[y, x], zero elsewhere
__global__ void zero_insert2x_kernel(const __nv_bfloat16* __restrict__ dz, __nv_bfloat16* __restrict__ u, int B, int h,
                                     int w, int C) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  const int cg = C / 8;
  const long long total = (long long)B * 4 * h * w * cg;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = (int)(i % cg);
  long long r = i / cg;
  const int ox = (int)(r % (2 * w));
  r /= (2 * w);
  const int oy = (int)(r % (2 * h));
  const long long b = r / (2 * h);
  uint4 v = make_uint4(0, 0, 0, 0);
  if (!(ox & 1) && !(oy & 1)) v = reinterpret_cast<const uint4*>(dz)[((b * h + oy / 2) * w + ox / 2) * cg + c8];
  reinterpret_cast<uint4*>(u)[i] = v;
}

__global__ void add_bf16_kernel(const __nv_bfloat16* __restrict__ a, long long lda, const __nv_bfloat16* __restrict__ b,
                                long long ldb, __nv_bfloat16* __restrict__ out, long long M, int C) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  const int cg = C / 8;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * cg) return;
  const int c = (int)(i % cg) * 8;
  const long long r = i / cg;
  const uint4 ua = *reinterpret_cast<const uint4*>(a + r * lda + c);
  const uint4 ub = *reinterpret_cast<const uint4*>(b + r * ldb + c);
  const uint32_t* pa = &ua.x; const uint32_t* pb = &ub.x;
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 fa = unpack_bf16x2(pa[j]), fb = unpack_bf16x2(pb[j]);
    o[j] = pack_bf16x2(fa.x + fb.x, fa.y + fb.y);
  }
  reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

// ds[p, h] = dpred[b, h, p] * (1 - pred[b, h, p]^2); dbias[h] += sum_p ds. grid = (pixel blocks, B); block = 256 pixels
__global__ void __launch_bounds__(256) heads_ds_kernel(const float* __restrict__ dpred, const float* __restrict__ pred,
                                                       __nv_bfloat16* __restrict__ ds, float* __restrict__ dbias, int heads,
                                                       int hw) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  __shared__ float sh[8][16];
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = p < hw;
  // all 2 x heads loads of this pixel are in flight together (they used to be serialised by two block barriers per head)
  float vals[16];
#pragma unroll
  for (int h = 0; h < 16; ++h) {
    float v = 0.f;
    if (ok && h < heads) {
      const long long i = ((long long)b * heads + h) * hw + p;
      const float y = pred[i];
      v = dpred[i] * (1.f - y * y);
    }
    vals[h] = v;
  }
  if (ok) {
    __nv_bfloat16* o = ds + ((long long)b * hw + p) * 16;
    uint32_t u[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) u[j] = pack_bf16x2(vals[2 * j], vals[2 * j + 1]);
    reinterpret_cast<uint4*>(o)[0] = make_uint4(u[0], u[1], u[2], u[3]);
    reinterpret_cast<uint4*>(o)[1] = make_uint4(u[4], u[5], u[6], u[7]);
  }
  // bias gradients: warp sums -> one row per warp in shared memory -> 16 atomics per block
#pragma unroll
  for (int h = 0; h < 16; ++h) {
    const float s = warp_sum(vals[h]);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5][h] = s;
  }
  __syncthreads();
  if (threadIdx.x < heads) {
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) t += sh[wv][threadIdx.x];
    atomicAdd(dbias + threadIdx.x, t);
  }
}

// per pixel p' (thread): for tap = (ky, kx): n = ds[p' - (ky-1, kx-1)] (zero outside the image)
//   dt[p', tap*16 + h] = g[p', h] * n[h];   dg[h] += n[h] * t[p', tap*16 + h];   du[h] = dg[h] * g (1 - g)
// t, dt: bf16 [M, 144]; ds, gate: bf16 [M, 16]; du: bf16 [M, 16]; db2[h] += sum_p du
__global__ void __launch_bounds__(128) heads_bwd_stencil_kernel(const __nv_bfloat16* __restrict__ t,
                                                                const __nv_bfloat16* __restrict__ ds,
                                                                const __nv_bfloat16* __restrict__ gate,
                                                                __nv_bfloat16* __restrict__ dt,
                                                                __nv_bfloat16* __restrict__ du, float* __restrict__ db2,
                                                                int H, int W, long long M) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  __shared__ float sh[4][16];
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = m < M;
  float dg[16], g[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) { dg[j] = 0.f; g[j] = 0.f; }
  if (ok) {
    const long long hw = (long long)H * W;
    const long long b = m / hw;
    const int rem = (int)(m - b * hw);
    const int py = rem / W, px = rem - py * W;
    {
      const uint4 a = reinterpret_cast<const uint4*>(gate + m * 16)[0], c = reinterpret_cast<const uint4*>(gate + m * 16)[1];
      const uint32_t u[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float2 f = unpack_bf16x2(u[j]); g[2 * j] = f.x; g[2 * j + 1] = f.y; }
    }
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = py - (tap / 3 - 1), xx = px - (tap % 3 - 1);
      float n[16];
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
        const __nv_bfloat16* q = ds + (b * hw + (long long)yy * W + xx) * 16;
        const uint4 a = reinterpret_cast<const uint4*>(q)[0], c = reinterpret_cast<const uint4*>(q)[1];
        const uint32_t u[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float2 f = unpack_bf16x2(u[j]); n[2 * j] = f.x; n[2 * j + 1] = f.y; }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) n[j] = 0.f;
      }
      const __nv_bfloat16* tp = t + m * 144 + tap * 16;
      const uint4 a = reinterpret_cast<const uint4*>(tp)[0], c = reinterpret_cast<const uint4*>(tp)[1];
      const uint32_t u[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
      uint32_t o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 f = unpack_bf16x2(u[j]);
        dg[2 * j] += n[2 * j] * f.x;
        dg[2 * j + 1] += n[2 * j + 1] * f.y;
        o[j] = pack_bf16x2(g[2 * j] * n[2 * j], g[2 * j + 1] * n[2 * j + 1]);
      }
      __nv_bfloat16* op = dt + m * 144 + tap * 16;
      reinterpret_cast<uint4*>(op)[0] = make_uint4(o[0], o[1], o[2], o[3]);
      reinterpret_cast<uint4*>(op)[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
  }
  float duv[16];
  uint32_t o[8];
#pragma unroll
  for (int j = 0; j < 16; ++j) duv[j] = dg[j] * g[j] * (1.f - g[j]);
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = pack_bf16x2(duv[2 * j], duv[2 * j + 1]);
  if (ok) {
    reinterpret_cast<uint4*>(du + m * 16)[0] = make_uint4(o[0], o[1], o[2], o[3]);
    reinterpret_cast<uint4*>(du + m * 16)[1] = make_uint4(o[4], o[5], o[6], o[7]);
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float s = warp_sum(duv[j]);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5][j] = s;
  }
  __syncthreads();
  if (threadIdx.x < 16) atomicAdd(db2 + threadIdx.x, sh[0][threadIdx.x] + sh[1][threadIdx.x] + sh[2][threadIdx.x] + sh[3][threadIdx.x]);
}

// bf16 twin of an fp16 tensor, 8 elements per thread
__global__ void f16_to_bf16_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, long long n8) {
  griddep_sync();  // PDL: block until the previous kernel of the stream has completed (mv_ptx.cuh)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 u = src[i];
    const float2 a = unpack16x2<true>(u.x), b = unpack16x2<true>(u.y), c = unpack16x2<true>(u.z), d = unpack16x2<true>(u.w);
    dst[i] = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(b.x, b.y), pack_bf16x2(c.x, c.y), pack_bf16x2(d.x, d.y));
  }
}

}  // namespace mv

extern "C" int mv_transpose_bf16(const void* in, int64_t ldi, void* out, int64_t ldo, int64_t m, int c, int ones_row,
                                 int in_f16, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(in && out && m > 0 && c > 0 && c % 2 == 0 && ldi % 2 == 0 && ldo % 2 == 0 && ldo >= m, "mv_transpose_bf16: shape");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (c % 8 == 0 && c <= 2048 && ldi % 8 == 0 && ldo % 8 == 0 &&
      ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    const int ncg = c / 8, per_blk = 256 / ncg > 0 ? 256 / ncg : 1;
    const unsigned vgrid = (unsigned)(((m + 7) / 8 + per_blk - 1) / per_blk);
    const int threads = ncg * per_blk;
    if (in_f16)
      MV_LAUNCH(transpose_bf16_vec_kernel<true>, vgrid, threads, 0, stream, reinterpret_cast<const __nv_bfloat16*>(in), ldi,
                reinterpret_cast<__nv_bfloat16*>(out), ldo, (long long)m, c, ones_row);
    else
      MV_LAUNCH(transpose_bf16_vec_kernel<false>, vgrid, threads, 0, stream, reinterpret_cast<const __nv_bfloat16*>(in), ldi,
                reinterpret_cast<__nv_bfloat16*>(out), ldo, (long long)m, c, ones_row);
    MV_CHECK_LAUNCH("transpose_bf16_vec");
    return MV_OK;
  }
  dim3 grid((unsigned)((m + 63) / 64), (c + 63) / 64), block(32, 8);
  if (in_f16)
    MV_LAUNCH(transpose_bf16_kernel<true>, grid, block, 0, stream, reinterpret_cast<const __nv_bfloat16*>(in), ldi,
              reinterpret_cast<__nv_bfloat16*>(out), ldo, (long long)m, c, ones_row);
  else
    MV_LAUNCH(transpose_bf16_kernel<false>, grid, block, 0, stream, reinterpret_cast<const __nv_bfloat16*>(in), ldi,
              reinterpret_cast<__nv_bfloat16*>(out), ldo, (long long)m, c, ones_row);
  MV_CHECK_LAUNCH("transpose_bf16");
  return MV_OK;
}

extern "C" int mv_upsample2x_bwd(const void* dup, int64_t ldu, void* dx, int batch, int h, int w, int c, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(dup && dx && batch > 0 && c % 8 == 0 && ldu % 8 == 0, "mv_upsample2x_bwd: shape");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const long long total = (long long)batch * h * w * (c / 8);
  MV_LAUNCH(upsample2x_bwd_kernel, (unsigned)((total + 255) / 256), 256, 0, stream, 
      reinterpret_cast<const __nv_bfloat16*>(dup), ldu, reinterpret_cast<__nv_bfloat16*>(dx), batch, h, w, c);
  MV_CHECK_LAUNCH("upsample2x_bwd");
  return MV_OK;
}

extern "C" int mv_zero_insert2x(const void* dz, void* u, int batch, int h, int w, int c, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(dz && u && batch > 0 && c % 8 == 0, "mv_zero_insert2x: shape");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const long long total = (long long)batch * 4 * h * w * (c / 8);
  MV_LAUNCH(zero_insert2x_kernel, (unsigned)((total + 255) / 256), 256, 0, stream, 
      reinterpret_cast<const __nv_bfloat16*>(dz), reinterpret_cast<__nv_bfloat16*>(u), batch, h, w, c);
  MV_CHECK_LAUNCH("zero_insert2x");
  return MV_OK;
}

extern "C" int mv_add_bf16(const void* a, int64_t lda, const void* b, int64_t ldb, void* out, int64_t m, int c,
                           void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(a && b && out && m > 0 && c % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0, "mv_add_bf16: shape");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const long long total = m * (c / 8);
  MV_LAUNCH(add_bf16_kernel, (unsigned)((total + 255) / 256), 256, 0, stream, 
      reinterpret_cast<const __nv_bfloat16*>(a), lda, reinterpret_cast<const __nv_bfloat16*>(b), ldb,
      reinterpret_cast<__nv_bfloat16*>(out), m, c);
  MV_CHECK_LAUNCH("add_bf16");
  return MV_OK;
}

// dpred / pred: NCHW fp32 [B, heads, H*W]; ds: bf16 [B*H*W, 16] (columns >= heads zero); dbias fp32 [heads] (+=)
extern "C" int mv_heads_ds(const float* dpred, const float* pred, void* ds, float* dbias, int batch, int heads, int hw,
                           void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(dpred && pred && ds && dbias && batch > 0 && heads >= 1 && heads <= 16, "mv_heads_ds: shape");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  dim3 grid((hw + 255) / 256, batch);
  MV_LAUNCH(heads_ds_kernel, grid, 256, 0, stream, dpred, pred, reinterpret_cast<__nv_bfloat16*>(ds), dbias, heads, hw);
  MV_CHECK_LAUNCH("heads_ds");
  return MV_OK;
}

extern "C" int mv_heads_bwd_stencil(const void* t, const void* ds, const void* gate, void* dt, void* du, float* db2,
                                    int batch, int h, int w, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(t && ds && gate && dt && du && db2 && batch > 0, "mv_heads_bwd_stencil: null/empty");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const long long M = (long long)batch * h * w;
  MV_LAUNCH(heads_bwd_stencil_kernel, (unsigned)((M + 127) / 128), 128, 0, stream, 
      reinterpret_cast<const __nv_bfloat16*>(t), reinterpret_cast<const __nv_bfloat16*>(ds),
      reinterpret_cast<const __nv_bfloat16*>(gate), reinterpret_cast<__nv_bfloat16*>(dt),
      reinterpret_cast<__nv_bfloat16*>(du), db2, h, w, M);
  MV_CHECK_LAUNCH("heads_bwd_stencil");
  return MV_OK;
}

extern "C" int mv_f16_to_bf16(const void* src, void* dst, int64_t n, void* stream_) {
  using namespace mv;
  MV_CHECK_ARG(src && dst && n > 0 && n % 8 == 0, "mv_f16_to_bf16: n must be a positive multiple of 8");
  MV_CHECK_ARG(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0, "mv_f16_to_bf16: alignment");
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const long long n8 = n / 8;
  long long blocks = (n8 + 255) / 256;
  const long long cap = (long long)(device_sms() > 0 ? device_sms() : 148) * 16;
  if (blocks > cap) blocks = cap;
  MV_LAUNCH(f16_to_bf16_kernel, (unsigned)blocks, 256, 0, stream, reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), n8);
  MV_CHECK_LAUNCH("f16_to_bf16");
  return MV_OK;
}
