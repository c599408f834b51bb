// Tensor-core lowering of the wide multi-band-diffusion convolutions (csrc/mbd.cu): every Conv1d / ConvTranspose1d whose
// input width is a multiple of 64 channels runs as a sum of shifted GEMMs on tcgen05,
//     out[co][n * ostride + ph] = bias[co] + emb[co] + resid[..] + sum_slot sum_ci  W_slot[co][ci] * F[n + shift_slot][col_slot + ci]
// where F is the TIME-major copy of the layer input after its GroupNorm + ReLU (k_mbd_prep_t), split into two bf16 terms
// F = hi + lo so the activations keep 16 mantissa bits, and W_slot is tap `slot` of the kernel repacked to bf16 [tap][Cout][Cin]
// at create time (k_mbd_pack_w).  Zero padding of the convolution = TMA out-of-bounds fill (negative / past-the-end rows read
// as zero), the stride-s convolution reads the same buffer through a [T/s][s * Cin] view, and the transposed convolution is
// s output phases with two taps each.  One CTA = 128 output channels x 128 time columns: A tiles (weights, 16 KB) and the
// two B tiles (hi, lo: 128 rows each, together the N = 256 UMMA operand) arrive by TMA into a 4-stage ring, one elected
// thread issues 4 x tcgen05.mma (128 x 256 x 16) per 64-channel block, the fp32 accumulator lives in 256 TMEM columns and
// four epilogue warps add hi + lo + bias, transpose 32 x 32 blocks through shared memory and add the residual / store
// channel-major rows with the lanes along time (coalesced).
//
// Numerics: weights are rounded to bf16 (the reference's own convolutions run in TF32 under torch's cuDNN default
// `allow_tf32=True`, 10-bit mantissa on BOTH operands); activations are exact to 2^-17; accumulation is fp32.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace mvb {

struct TcTaps {
  int n_slots, n_ph;
  int a_z[4][8];      // weight tap (z slice of the A map) per (phase, slot)
  int b_shift[4][8];  // row shift into the time-major activations
  int b_col[4][8];    // column base (stride-s convolution: phase * Cin)
};
struct TcConvP {
  int Cin, M, Ncols;          // K per slot; output channels per phase; GEMM columns (time rows producing outputs)
  float* out; int ldo, ostride;
  const float* bias; const float* emb; const float* resid;
  TcTaps taps;
};

constexpr int TC_STAGES = 4;
constexpr int TC_A_BYTES = 128 * 64 * 2, TC_B_BYTES = 256 * 64 * 2, TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr size_t TC_SMEM = 1024 + (size_t)TC_STAGES * TC_STAGE_BYTES + (2 * TC_STAGES + 1) * 8 + 64 + 4 * 32 * 33 * 4;   // + epilogue transpose tiles

__device__ __forceinline__ void tc_tma3(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

static __global__ void __launch_bounds__(256, 1)
k_mbd_tc_conv(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
              const __grid_constant__ CUtensorMap tmBlo, const TcConvP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int S = TC_STAGES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * TC_STAGE_BYTES);   // full[S], empty[S], acc
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_cot = (p.M + 127) >> 7;
  const int ph = blockIdx.x / n_cot, cot = blockIdx.x - ph * n_cot;
  const int t0 = blockIdx.y * 128;
  const int nkb = p.Cin >> 6;
  const int n_it = p.taps.n_slots * nkb;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(ptx::smem_u32(bars + s), 1);
      ptx::mbar_init(ptx::smem_u32(bars + S + s), 1);
    }
    ptx::mbar_init(ptx::smem_u32(bars + 2 * S), 1);
    ptx::fence_barrier_init();
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmBhi);
    ptx::prefetch_tensormap(&tmBlo);
  }
  if (warp == 2) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    int slot = 0, kb = 0;
    for (int it = 0; it < n_it; ++it) {
      const int s = it % S;
      const uint32_t par = (uint32_t)(it / S) & 1u;
      ptx::mbar_wait(ptx::smem_u32(bars + S + s), par ^ 1u);
      if (ptx::elect_one()) {
        const uint32_t full = ptx::smem_u32(bars + s);
        ptx::mbar_arrive_expect_tx(full, (uint32_t)TC_STAGE_BYTES);
        uint8_t* st = smem + (size_t)s * TC_STAGE_BYTES;
        const int row = t0 + p.taps.b_shift[ph][slot], col = p.taps.b_col[ph][slot] + kb * 64;
        tc_tma3(ptx::smem_u32(st), &tmA, full, kb * 64, cot * 128, p.taps.a_z[ph][slot]);
        ptx::tma_load_2d(ptx::smem_u32(st + TC_A_BYTES), &tmBhi, full, col, row);
        ptx::tma_load_2d(ptx::smem_u32(st + TC_A_BYTES + TC_B_BYTES / 2), &tmBlo, full, col, row);
      }
      __syncwarp();
      if (++kb == nkb) { kb = 0; ++slot; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = ptx::umma_idesc_bf16(128, 256);
    for (int it = 0; it < n_it; ++it) {
      const int s = it % S;
      const uint32_t par = (uint32_t)(it / S) & 1u;
      ptx::mbar_wait(ptx::smem_u32(bars + s), par);
      ptx::tc_fence_after();
      uint8_t* st = smem + (size_t)s * TC_STAGE_BYTES;
      const uint64_t ad = ptx::umma_desc_k_sw128(ptx::smem_u32(st));
      const uint64_t bd = ptx::umma_desc_k_sw128(ptx::smem_u32(st + TC_A_BYTES));
      if (ptx::elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem_base, ad + 2 * k, bd + 2 * k, idesc, (uint32_t)((it | k) != 0));
        ptx::umma_commit(ptx::smem_u32(bars + S + s));
      }
      __syncwarp();
    }
    if (ptx::elect_one()) ptx::umma_commit(ptx::smem_u32(bars + 2 * S));
    __syncwarp();
  } else if (warp >= 4) {
    // Epilogue: each warp owns 32 accumulator lanes (output channels).  A 32 x 32 block goes TMEM -> registers (lane = channel)
    // -> a padded smem tile -> global with lane = TIME, so that the residual reads and the stores are contiguous 128-byte rows
    // instead of 32 scattered sectors per instruction (the channel-major activations have stride ldo between channels).
    const int w4 = warp - 4;
    const int co_w = cot * 128 + 32 * w4;               // first channel of this warp
    const uint32_t tbase = tmem_base + ((uint32_t)(32 * w4) << 16);
    float* tp = reinterpret_cast<float*>(tmem_slot + 4) + w4 * (32 * 33);
    ptx::mbar_wait(ptx::smem_u32(bars + 2 * S), 0);
    ptx::tc_fence_after();
    if (co_w < p.M) {
      const int co = co_w + lane;
      float add = 0.f;
      if (co < p.M) add = (p.bias ? p.bias[co] : 0.f) + (p.emb ? p.emb[co] : 0.f);
      const int rows = min(32, p.M - co_w);
      for (int c0 = 0; c0 < 128 && t0 + c0 < p.Ncols; c0 += 32) {
        uint32_t hi[32], lo[32];
        ptx::tmem_ld16(tbase + c0, *reinterpret_cast<uint32_t(*)[16]>(hi));
        ptx::tmem_ld16(tbase + c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(hi + 16));
        ptx::tmem_ld16(tbase + 128 + c0, *reinterpret_cast<uint32_t(*)[16]>(lo));
        ptx::tmem_ld16(tbase + 128 + c0 + 16, *reinterpret_cast<uint32_t(*)[16]>(lo + 16));
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) tp[lane * 33 + i] = __uint_as_float(hi[i]) + __uint_as_float(lo[i]) + add;
        __syncwarp();
        const int n = t0 + c0 + lane;
        if (n < p.Ncols) {
          const size_t col = (size_t)n * p.ostride + ph;
          // residual reads in batches of 8 rows: issued back to back, not one dependent load -> store pair per row
          for (int r0 = 0; r0 < rows; r0 += 8) {
            float rv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int r = min(r0 + j, rows - 1);
              rv[j] = p.resid ? __ldg(p.resid + (size_t)(co_w + r) * p.ldo + col) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (r0 + j < rows) p.out[(size_t)(co_w + r0 + j) * p.ldo + col] = tp[(r0 + j) * 33 + lane] + rv[j];
          }
        }
        __syncwarp();
      }
    }
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 256);
  }
}

// Time-major two-term bf16 copy of a layer input: hi/lo[t][c] = split(f(x[c][t])), f = relu(GroupNorm) when stats != null;
// rows T .. Tpad-1 are written as zero (right padding of the strided convolution).  Tile: 64 channels x 32 samples.
static __global__ void __launch_bounds__(256) k_mbd_prep_t(const float* __restrict__ x, int C, int Cp, int T, int Tpad, const float* __restrict__ stats,
                                                           const float* __restrict__ gw, const float* __restrict__ gb, int cpg,
                                                           __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  __shared__ float tile[64][33];
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + ty * 8 + i, t = t0 + tx;
    float v = 0.f;
    if (c < C && t < T) {
      v = x[(size_t)c * T + t];
      if (stats) {
        const int g = c / cpg;
        v = fmaxf((v - stats[2 * g]) * stats[2 * g + 1] * gw[c] + gb[c], 0.f);
      }
    }
    tile[ty * 8 + i][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int tl = ty * 4 + i, t = t0 + tl, c = c0 + 2 * tx;
    if (t < Tpad && c < Cp) {       // columns C .. Cp-1 (channel padding to a multiple of 64) are written as zero
      const float a = tile[2 * tx][tl], b = tile[2 * tx + 1][tl];
      const __nv_bfloat16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
      __nv_bfloat162 h2, l2;
      h2.x = ah; h2.y = bh;
      l2.x = __float2bfloat16_rn(a - __bfloat162float(ah));
      l2.y = __float2bfloat16_rn(b - __bfloat162float(bh));
      *reinterpret_cast<__nv_bfloat162*>(hi + (size_t)t * Cp + c) = h2;
      *reinterpret_cast<__nv_bfloat162*>(lo + (size_t)t * Cp + c) = l2;
    }
  }
}

// Weight repack to bf16 [tap][Cout][Cp]: transposed = 0: w is Conv1d [Cout][Cin][K]; 1: ConvTranspose1d [Cin][Cout][K].
// (Cp = Cin rounded up to a multiple of 64; the padding columns are zero.)
static __global__ void k_mbd_pack_w(const float* __restrict__ w, int Cout, int Cin, int Cp, int K, int transposed, __nv_bfloat16* __restrict__ out) {
  const size_t n = (size_t)K * Cout * Cp;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cp), co = (int)((i / Cp) % Cout), k = (int)(i / ((size_t)Cp * Cout));
    const size_t src = transposed ? ((size_t)ci * Cout + co) * K + k : ((size_t)co * Cin + ci) * K + k;
    out[i] = __float2bfloat16_rn(ci < Cin ? w[src] : 0.f);
  }
}

}  // namespace mvb
