// Host-side plumbing shared by all translation units: error reporting, launch accounting, TMA descriptor cache.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/miphei_b200.h"

namespace mv {

void set_error(const char* fmt, ...);
int device_sms();  // SM count of the CURRENT device (0 before mv_init on it)
void count_launch(int n = 1);
// true the first time it is called on the current device for this flag word — per-device one-time setup such as
// cudaFuncSetAttribute (function attributes are per device; a process may drive several GPUs)
bool first_use_on_device(std::atomic<uint64_t>& mask);

// 2-D bf16 row-major tensor map: inner extent `cols` (elements), outer extent `rows`, row pitch `ld` elements.
// Box = box_cols x box_rows, 128-byte swizzle (box_cols * 2 bytes must be 128), OOB reads return zeros.
// Cached by (ptr, rows, cols, ld, box); returns nullptr and sets the error string on failure. The returned pointer is a
// thread-local COPY valid until the calling thread's 16th following lookup (pass it to the kernel by value).
const CUtensorMap* get_tmap_2d_bf16(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                                    uint32_t box_cols = 64);
// 2-D fp32 row-major map with a 32-column (128-byte) x box_rows box, 128-byte swizzle: residual-stream tiles read / written
// by the TMA epilogue of the fp32-residual GEMM. Cached like the bf16 maps.
const CUtensorMap* get_tmap_2d_f32(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
// 4-D NHWC bf16 map [B, H, W, C] with box {64 channels, tw*stride, th*stride, 1} traversed with element stride
// `stride` in W and H (tw x th pixels land in smem as 128-byte rows, 128-byte swizzle). Cached.
const CUtensorMap* get_tmap_nhwc_bf16(const void* ptr, int batch, int h, int w, int c, int tw, int th, int stride);
// Generic encoder (rank <= 5) for kernels that need other layouts; not cached.
int encode_tmap(CUtensorMap* out, CUtensorMapDataType dt, uint32_t rank, void* ptr, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz,
                const uint32_t* elem_strides = nullptr);

// Launch with programmatic stream serialization (PDL) — the kernel MUST call griddep_wait() (mv_ptx.cuh) before touching
// global memory. MV_PDL=0 in the environment turns the attribute off (plain stream order). cluster_x > 1 adds a cluster.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
// MV_LAUNCH(kernel, grid, block, smem, stream, args...): PDL launch, errors surface through MV_CHECK_LAUNCH
#define MV_LAUNCH(kern, grid, block, smem, stream, ...) \
  (void)mv::launch_pdl(kern, dim3(grid), dim3(block), (size_t)(smem), stream, 1, __VA_ARGS__)

#define MV_CHECK_ARG(cond, ...)       \
  do {                                \
    if (!(cond)) {                    \
      mv::set_error(__VA_ARGS__);     \
      return MV_ERR_ARG;              \
    }                                 \
  } while (0)

#define MV_CHECK_LAUNCH(name)                                                   \
  do {                                                                          \
    cudaError_t e__ = cudaGetLastError();                                       \
    if (e__ != cudaSuccess) {                                                   \
      mv::set_error("%s launch failed: %s", name, cudaGetErrorString(e__));     \
      return (int)e__;                                                          \
    }                                                                           \
    mv::count_launch();                                                         \
  } while (0)

}  // namespace mv
