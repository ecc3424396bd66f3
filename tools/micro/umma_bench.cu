// tcgen05.mma issue / completion cost on B200 as a function of N, of how many independent accumulators the k-steps rotate
// over, and of where A comes from (shared memory or TMEM).  One CTA per SM, one issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I miphei-vit_b200/csrc -o /tmp/umma_bench tools/micro/umma_bench.cu && /tmp/umma_bench
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "mv_ptx.cuh"

using namespace mv;

template <int N, int NACC, bool TS, bool WARP>
__global__ void __launch_bounds__(128, 1) bench(int outer, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 96 * 1024, slot = bar + 16;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  constexpr int STRIDE = NACC > 1 ? (448 / NACC) / 32 * 32 : 0;  // accumulators STRIDE columns apart; last 64 columns: A in TMEM
  constexpr uint32_t idesc = umma_idesc_bf16(128, N);
  const uint64_t da = umma_desc_sw128(base), db = umma_desc_sw128(base + 32 * 1024);
  // WARP: the whole warp runs the role code and one ELECTED lane issues (warp-uniform control flow);
  // otherwise the code sits under `if (threadIdx.x == 0)` like the kernels of this repository
  const bool me = WARP ? (threadIdx.x < 32) : (threadIdx.x == 0);
  if (me) {
    const bool issuer = WARP ? elect_one() : true;
    const long long t0 = clock64();
    for (int o = 0; o < outer; ++o) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint32_t d = tmem + (i % NACC) * STRIDE;
        if (issuer) {
          if (TS) umma_bf16_ts(d, tmem + 448 + (i & 3) * 8, db + 2 * (i & 3), idesc, (o | (i >= NACC)) != 0);
          else umma_bf16(d, da + 2 * (i & 3), db + 2 * (i & 3), idesc, (o | (i >= NACC)) != 0);
        }
      }
    }
    const long long t1 = clock64();
    if (issuer) umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, int NACC, bool TS, bool WARP>
void run(long long* d) {
  const int smem = 100 * 1024, outer = 16;
  cudaFuncSetAttribute(bench<N, NACC, TS, WARP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) bench<N, NACC, TS, WARP><<<148, 128, smem>>>(outer, d);
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  printf("%-6s %-8s %-6d %-6d %14.1f %14.1f\n", WARP ? "elect" : "lane0", TS ? "tmem" : "smem", N, NACC, (double)h[0] / (16 * outer),
         (double)h[1] / (16 * outer));
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  printf("%-6s %-8s %-6s %-6s %14s %14s\n", "issue", "A", "N", "accs", "issue clk/mma", "total clk/mma");
  run<32, 1, false, false>(d); run<64, 1, false, false>(d); run<64, 2, false, false>(d); run<64, 4, false, false>(d);
  run<112, 1, false, false>(d); run<112, 2, false, false>(d); run<128, 1, false, false>(d); run<256, 1, false, false>(d);
  run<64, 1, true, false>(d); run<64, 2, true, false>(d); run<64, 4, true, false>(d); run<128, 1, true, false>(d);
  run<64, 1, false, true>(d); run<64, 2, false, true>(d); run<64, 4, false, true>(d); run<112, 1, false, true>(d);
  run<256, 1, false, true>(d); run<64, 1, true, true>(d); run<64, 4, true, true>(d);
  return 0;
}
