// Micro-benchmark: cost of one tcgen05.mma (kind::f16, bf16 operands, K = 16) as a function of its shape and of where the
// A operand lives, issued back-to-back by one thread the way a weight-streaming consumer does (4 K-steps per 16 KB tile,
// one tcgen05.commit per tile).  Modes: SS = both operands in shared memory (128B-swizzled K-major tiles), TS = A copied
// to tensor memory first with tcgen05.cp.128x256b.  Prints cycles per MMA and per 16 KB "tile" (4 instructions).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/umma_bench tools/micro/umma_bench.cu -lcudart
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../../metavoice-src_b200/csrc/umma.cuh"
using namespace mvb;

// mode 0: SS, commit per tile, wait for the commit every `depth` tiles (ring-like)
// mode 1: TS (cp 4 slices, commit, 4 TS MMAs), same waiting
template <bool UNIFORM>
__global__ void __launch_bounds__(128) k_bench(int M, int N, int mode, int tiles, int depth, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* A = smem;                 // 8 x 16 KB tiles [128 rows][64 k]
  uint8_t* B = smem + 8 * 16384;     // [256 rows][64 k] = 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(B + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 16);
  for (int i = threadIdx.x; i < (8 * 16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) ptx::mbar_init(ptx::smem_u32(bars + i), 1);
    ptx::fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    ptx::tmem_alloc(ptx::smem_u32(slot), 512);
    ptx::tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = __shfl_sync(0xffffffffu, *slot, 0);   // provably warp-uniform: no operand waterfall around the tcgen05 instructions
  // UNIFORM: the whole warp runs the loop and one elected lane issues (the CUTLASS pattern); otherwise a lane-0 branch
  if (UNIFORM ? (threadIdx.x < 32) : (threadIdx.x == 0)) {
    const uint32_t idesc = ptx::umma_idesc_bf16(M, N);
    const uint64_t bd = ptx::umma_desc_k_sw128(ptx::smem_u32(B));
    const long long t0 = clock64();
    for (int t = 0; t < tiles; ++t) {
      const uint32_t s = t % depth, ph = (t / depth) & 1u;
      if (t >= depth) ptx::mbar_wait(ptx::smem_u32(bars + s), ph ^ 1u);   // the commit of tile t - depth has fired
      const uint64_t ad = ptx::umma_desc_k_sw128(ptx::smem_u32(A + (t & 7) * 16384));
      if (mode == 0) {
        if (!UNIFORM || ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_bf16(tm, ad + 2 * k, bd + 2 * k, idesc, 1u);
          ptx::umma_commit(ptx::smem_u32(bars + s));
        }
        if (UNIFORM) __syncwarp();
      } else if (!UNIFORM || ptx::elect_one()) {
        const uint32_t ab = tm + 256 + (t & 3) * 32;
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::tmem_cp_128x256b(ab + 8 * k, ad + 2 * k);
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_bf16_ts(tm, ab + 8 * k, bd + 2 * k, idesc, 1u);
        ptx::umma_commit(ptx::smem_u32(bars + s));
      }
    }
    // drain: wait for the last commit
    const int t = tiles - 1;
    ptx::mbar_wait(ptx::smem_u32(bars + (t % depth)), (t / depth) & 1u);
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tm, 512);
  }
}

int main() {
  long long* d; cudaMalloc(&d, 8 * 148);
  const size_t smem = 1024 + 8 * 16384 + 32768 + 16 * 8 + 64;
  cudaFuncSetAttribute(k_bench<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_bench<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int tiles = 2000;
  printf("%-4s %-4s %-4s %-6s %-5s | cycles/MMA  cycles/tile(4 MMA)  us/tile@1.9GHz\n", "mode", "M", "N", "depth", "ctas");
  for (int uni : {0, 1})
  for (int ctas : {1, 148})
    for (int mode : {0, 1})
      for (int M : {128})
        for (int N : {16, 256})
          for (int depth : {1, 8}) {
            if (mode == 1 && M == 64) continue;
            if (ctas == 148 && depth == 1) continue;
            if (uni) k_bench<true><<<ctas, 128, smem>>>(M, N, mode, tiles, depth, d);
            else k_bench<false><<<ctas, 128, smem>>>(M, N, mode, tiles, depth, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[148];
            cudaMemcpy(h, d, 8 * ctas, cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("%s %-4s %-4d %-4d %-6d %-5d | %9.1f  %9.1f  %9.3f  %s\n", uni ? "warp+elect" : "lane0     ", mode ? "TS" : "SS", M, N, depth, ctas, (double)mx / tiles / 4,
                   (double)mx / tiles, (double)mx / tiles / 1900.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
          }
  return 0;
}
