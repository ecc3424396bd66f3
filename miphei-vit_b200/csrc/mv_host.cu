// Host plumbing: init, error strings, launch counter, TMA descriptor creation through the driver entry point
// (resolved at run time with cudaGetDriverEntryPoint so the library does not link libcuda).
#include "mv_host.h"
#include <stdlib.h>

#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <unordered_map>

namespace mv {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};
// one process may drive several GPUs (one host thread each): everything device-specific is indexed by the CURRENT device
constexpr int kMaxDevices = 64;
static int g_sms[kMaxDevices] = {0};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static int current_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) d = 0;
  return d;
}
int device_sms() { return g_sms[current_device()]; }
bool first_use_on_device(std::atomic<uint64_t>& mask) {
  const uint64_t bit = 1ull << current_device();
  return (mask.fetch_or(bit, std::memory_order_acq_rel) & bit) == 0;
}
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("MV_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int encode_tmap(CUtensorMap* out, CUtensorMapDataType dt, uint32_t rank, void* ptr, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz,
                const uint32_t* elem_strides) {
  if (!g_encode) {
    set_error("mv_init() has not been called (no TMA encoder)");
    return MV_ERR_DEVICE;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (uint32_t i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = elem_strides ? elem_strides[i] : 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUresult r = g_encode(out, dt, rank, ptr, gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %u dims %llu,%llu ld %llu box %u,%u ptr %p)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0], rank > 1 ? box[1] : 0, ptr);
    return MV_ERR_ARG;
  }
  return MV_OK;
}

struct TmapKey {
  const void* p;
  uint64_t rows, cols, ld;
  uint32_t br, bc;
  bool operator==(const TmapKey& o) const {
    return p == o.p && rows == o.rows && cols == o.cols && ld == o.ld && br == o.br && bc == o.bc;
  }
};
struct TmapHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = reinterpret_cast<uint64_t>(k.p) * 0x9E3779B97F4A7C15ull;
    h ^= (k.rows + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (k.cols * 1315423911ull + (h << 6) + (h >> 2));
    h ^= (k.ld * 2654435761ull + (h << 6) + (h >> 2));
    h ^= ((uint64_t)k.br << 32 | k.bc) + (h << 6) + (h >> 2);
    return (size_t)h;
  }
};
static std::mutex g_tmap_mu;
static std::unordered_map<TmapKey, CUtensorMap*, TmapHash> g_tmaps;
constexpr size_t kTmapCacheMax = 65536;

// Callers never hold pointers into the cache: every lookup hands out a COPY in a small thread-local ring (a launch takes
// at most four descriptors and passes them to the kernel by value), so evicting cache entries under the lock is always safe.
static const CUtensorMap* hand_out(const CUtensorMap* cached) {
  static thread_local CUtensorMap ring[16];
  static thread_local unsigned next = 0;
  CUtensorMap* slot = &ring[next++ & 15u];
  memcpy(slot, cached, sizeof(CUtensorMap));
  return slot;
}
static void evict_if_full() {  // g_tmap_mu held
  if (g_tmaps.size() <= kTmapCacheMax) return;
  for (auto& kv : g_tmaps) free(kv.second);
  g_tmaps.clear();
}

const CUtensorMap* get_tmap_2d_bf16(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                                    uint32_t box_cols) {
  TmapKey key{ptr, rows, cols, ld, box_rows, box_cols};
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  auto it = g_tmaps.find(key);
  if (it != g_tmaps.end()) return hand_out(it->second);
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * 2) & 15)) {
    set_error("TMA operand must be 16-byte aligned with a 16-byte multiple pitch (ptr %p ld %llu)", ptr,
              (unsigned long long)ld);
    return nullptr;
  }
  void* mem = nullptr;
  if (posix_memalign(&mem, 64, sizeof(CUtensorMap)) != 0) {
    set_error("out of host memory");
    return nullptr;
  }
  CUtensorMap* tm = reinterpret_cast<CUtensorMap*>(mem);
  uint64_t dims[2] = {cols, rows};
  uint64_t strides[1] = {ld * 2};
  uint32_t box[2] = {box_cols, box_rows};
  if (encode_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box,
                  CU_TENSOR_MAP_SWIZZLE_128B) != MV_OK) {
    free(mem);
    return nullptr;
  }
  evict_if_full();  // unbounded growth guard: descriptors are tiny, but operand pointers may churn
  g_tmaps.emplace(key, tm);
  return hand_out(tm);
}

const CUtensorMap* get_tmap_2d_f32(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  TmapKey key{ptr, rows, cols, ld | (1ull << 62), box_rows, 32u};  // tag bit 62: fp32 map (never collides with bf16 keys)
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  auto it = g_tmaps.find(key);
  if (it != g_tmaps.end()) return hand_out(it->second);
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * 4) & 15)) {
    set_error("fp32 TMA operand must be 16-byte aligned with a 16-byte multiple pitch (ptr %p ld %llu)", ptr, (unsigned long long)ld);
    return nullptr;
  }
  void* mem = nullptr;
  if (posix_memalign(&mem, 64, sizeof(CUtensorMap)) != 0) {
    set_error("out of host memory");
    return nullptr;
  }
  CUtensorMap* tm = reinterpret_cast<CUtensorMap*>(mem);
  uint64_t dims[2] = {cols, rows};
  uint64_t strides[1] = {ld * 4};
  uint32_t box[2] = {32, box_rows};
  if (encode_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) !=
      MV_OK) {
    free(mem);
    return nullptr;
  }
  evict_if_full();
  g_tmaps.emplace(key, tm);
  return hand_out(tm);
}

const CUtensorMap* get_tmap_nhwc_bf16(const void* ptr, int batch, int h, int w, int c, int tw, int th, int stride) {
  // key reuse: rows = batch<<32|h, cols = w<<32|c, ld = stride, box = (th, tw) with a tag bit so 2-D maps never collide
  TmapKey key{ptr, ((uint64_t)batch << 32) | (uint32_t)h, ((uint64_t)w << 32) | (uint32_t)c, (uint64_t)stride | (1ull << 63),
              (uint32_t)th, (uint32_t)tw};
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  auto it = g_tmaps.find(key);
  if (it != g_tmaps.end()) return hand_out(it->second);
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (c % 8)) {
    set_error("NHWC TMA source must be 16-byte aligned with C %% 8 == 0 (ptr %p C %d)", ptr, c);
    return nullptr;
  }
  void* mem = nullptr;
  if (posix_memalign(&mem, 64, sizeof(CUtensorMap)) != 0) {
    set_error("out of host memory");
    return nullptr;
  }
  CUtensorMap* tm = reinterpret_cast<CUtensorMap*>(mem);
  uint64_t dims[4] = {(uint64_t)c, (uint64_t)w, (uint64_t)h, (uint64_t)batch};
  uint64_t strides[3] = {(uint64_t)c * 2, (uint64_t)w * c * 2, (uint64_t)h * w * c * 2};
  uint32_t box[4] = {64, (uint32_t)(tw * stride), (uint32_t)(th * stride), 1};
  uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
  if (encode_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box,
                  CU_TENSOR_MAP_SWIZZLE_128B, es) != MV_OK) {
    free(mem);
    return nullptr;
  }
  evict_if_full();
  g_tmaps.emplace(key, tm);
  return hand_out(tm);
}

}  // namespace mv

extern "C" {

int mv_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    mv::set_error("no CUDA device visible (%s); miphei_b200 has no CPU fallback", cudaGetErrorString(e));
    return MV_ERR_DEVICE;
  }
  if (device < 0 || device >= n) {
    mv::set_error("device %d out of range (%d visible)", device, n);
    return MV_ERR_ARG;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    mv::set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    return (int)e;
  }
  if (prop.major != 10) {
    mv::set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    return MV_ERR_DEVICE;
  }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    mv::set_error("cudaSetDevice: %s", cudaGetErrorString(e));
    return (int)e;
  }
  if (device < mv::kMaxDevices) mv::g_sms[device] = prop.multiProcessorCount;
  if (!mv::g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      mv::set_error("cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
      return MV_ERR_DEVICE;
    }
    mv::g_encode = reinterpret_cast<mv::EncodeTiledFn>(fn);
  }
  return MV_OK;
}

const char* mv_last_error(void) { return mv::g_err; }
int mv_version(void) { return 100; }
int mv_num_sms(void) { return mv::device_sms(); }
int64_t mv_launch_count(void) { return mv::g_launches.load(); }
void mv_reset_launch_count(void) { mv::g_launches.store(0); }

}  // extern "C"
