// Micro-benchmark: per-SM TMA ingest rate of [128 x 64] bf16 tiles (128B swizzle) through an N-stage mbarrier ring,
// consumer = one thread that releases a stage as soon as it is full.  Modes: tiles streamed once from HBM (distinct
// addresses, row stride 4 KB like the weight matrices) or re-read from a small L2-resident region.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/tma_bench tools/micro/tma_bench.cu -lcudart
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../metavoice-src_b200/csrc/umma.cuh"
#include "../../metavoice-src_b200/csrc/umma_host.cuh"
using namespace mvb;

template <int STAGES, int NP>
__global__ void __launch_bounds__(32 * (NP + 1)) k_ring(const __grid_constant__ CUtensorMap tm, int tiles_per_cta, int row_tiles, int kblocks, int l2_mode,
                                             long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * 16384);
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2 * STAGES; ++s) ptx::mbar_init(ptx::smem_u32(bars + s), 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < NP) {          // producers (one thread in each of NP warps)
    for (int i = threadIdx.x >> 5; i < tiles_per_cta; i += NP) {
      const int s = i % STAGES; const uint32_t ph = (i / STAGES) & 1u;
      ptx::mbar_wait(ptx::smem_u32(bars + STAGES + s), ph ^ 1u);
      ptx::mbar_arrive_expect_tx(ptx::smem_u32(bars + s), 16384);
      int g = l2_mode ? (blockIdx.x * 4 + (i & 3)) : (blockIdx.x * tiles_per_cta + i);   // global tile id
      const int rt = (g / kblocks) % row_tiles, kb = g % kblocks;
      ptx::tma_load_2d(ptx::smem_u32(smem + s * 16384), &tm, ptx::smem_u32(bars + s), kb * 64, rt * 128);
    }
  } else if (threadIdx.x == 32 * NP) {  // consumer
    for (int i = 0; i < tiles_per_cta; ++i) {
      const int s = i % STAGES; const uint32_t ph = (i / STAGES) & 1u;
      ptx::mbar_wait(ptx::smem_u32(bars + s), ph);
      ptx::mbar_arrive(ptx::smem_u32(bars + STAGES + s));
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

template <int STAGES, int NP>
static void run(const CUtensorMap& tm, int ctas, int tiles, int row_tiles, int kblocks, int l2, const char* name) {
  long long* d; cudaMalloc(&d, sizeof(long long) * ctas);
  const size_t smem = 1024 + STAGES * 16384 + 2 * STAGES * 8 + 64;
  cudaFuncSetAttribute(k_ring<STAGES, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_ring<STAGES, NP><<<ctas, 32 * (NP + 1), smem>>>(tm, tiles, row_tiles, kblocks, l2, d);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_ring<STAGES, NP><<<ctas, 32 * (NP + 1), smem>>>(tm, tiles, row_tiles, kblocks, l2, d);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-6s producers=%d stages=%2d ctas=%3d tiles/cta=%4d : %.3f ms  -> %.1f GB/s total, %.1f GB/s per SM, %.3f us/tile  (%s)\n", name, NP, STAGES, ctas, tiles, ms,
         (double)ctas * tiles * 16384 / ms / 1e6, (double)tiles * 16384 / ms / 1e6, ms * 1e3 / tiles, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  const int M = 128 * 1024, K = 2048;                 // 512 MB of bf16, row stride 4 KB
  void* w; cudaMalloc(&w, (size_t)M * K * 2); cudaMemset(w, 1, (size_t)M * K * 2);
  CUtensorMap tm;
  if (!make_tmap_bf16(&tm, w, M, K, 128)) { printf("tensor map failed\n"); return 1; }
  const int row_tiles = M / 128, kblocks = K / 64;
  for (int ctas : {1, 148}) {
    run<8, 1>(tm, ctas, 200, row_tiles, kblocks, 0, "hbm");
    run<8, 2>(tm, ctas, 200, row_tiles, kblocks, 0, "hbm");
    run<8, 4>(tm, ctas, 200, row_tiles, kblocks, 0, "hbm");
    run<8, 1>(tm, ctas, 2000, row_tiles, kblocks, 1, "l2");
    run<8, 2>(tm, ctas, 2000, row_tiles, kblocks, 1, "l2");
    run<8, 4>(tm, ctas, 2000, row_tiles, kblocks, 1, "l2");
    run<12, 4>(tm, ctas, 2000, row_tiles, kblocks, 1, "l2");
  }
  return 0;
}
