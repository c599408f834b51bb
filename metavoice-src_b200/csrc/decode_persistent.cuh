// Path C: ONE persistent kernel per decode position (fast_model.py:150-163 for S = 1), all 24 layers.
//
//   * grid = one CTA per SM; every CTA owns a static slice of every weight matrix:
//       (K split s = cta % S, row tiles t = cta / S, + groups, ...)  chosen per matrix on the host.
//   * warp 0 = PRODUCER: one thread walks the whole step's byte schedule -- weight tiles via TMA
//     (cp.async.bulk.tensor, 128B swizzle) into a ring of 16 KB smem stages, KV-cache tiles via cp.async.bulk
//     into two dedicated (K,V) tile slots -- all guarded by full/empty mbarriers.  Weights and old KV do not
//     depend on this step's activations, so the producer never waits for a grid barrier: the HBM stream keeps
//     running across phase boundaries, and a layer's first KV tiles are in flight before its QKV GEMM starts.
//   * warp 1 = MMA ISSUER: one thread issues tcgen05.mma (M=128 weight rows x N activation columns, fp32
//     accumulators double-buffered in TMEM) straight from the ring; activations are the exact hi+lo bf16
//     split of the fp32 vectors (N = 16: 8 rows, N = 32: 16 rows), so products match fp32-activation math to ~1e-5.
//   * warps 2..5 = COMPUTE: wait for the previous phase grid-wide (split arrive/wait counter with
//     release/acquire atomics), stage the B operand into swizzled smem (RMSNorm / attention-merge / SiLU*mul
//     fused here), run the epilogues (tcgen05.ld -> red.global.add.f32 split-K accumulation) and the decode
//     attention on smem KV tiles (online softmax per half-warp).
//   * phases per layer: QKV | attention (+KV-cache append) | wo+residual | w1,w3 | w2+residual, then head.
//     5 grid-wide dependencies per layer, no host involvement, no per-layer launch.
#pragma once
#include "stage1_kernels.cuh"
#include "umma.cuh"

namespace mvb {

constexpr int PC_THREADS = 224;      // producer warp + MMA warp + 4 compute warps + second producer warp
constexpr int PC_STAGE_BYTES = 16384;
constexpr int PC_RPAD = 16;          // rows of the fp32 activation buffers (8 utterances x 2 CFG rows)
constexpr int PC_BKB_MAX = 12;       // k-blocks of B one CTA may own in a phase
constexpr int PC_NKV = 2;            // (K tile, V tile) slots
constexpr int PC_MAX_CHUNKS = 64;    // per (row, head): ceil(2048 / 32) in fp32 mode
constexpr int PC_TRACE_EVENTS = 512;
constexpr int PC_MAX_TILES = 128;    // attention KV tiles one CTA may own per layer

template <int NB> struct PcCfg {
  static constexpr int STAGES = (NB == 16) ? 8 : 6;
  static constexpr int B_BYTES = PC_BKB_MAX * NB * 128;
  static constexpr int RH = NB / 2;                      // activation rows carried (hi rows; lo rows follow)
  static constexpr int TMEM_COLS = 256;   // 2 accumulators (2 NB columns) + 4 weight-tile A buffers of 32 columns (A-in-TMEM mode)
  static constexpr int A_COL0 = 2 * NB;
  static constexpr size_t SMEM = 1024 + (size_t)STAGES * PC_STAGE_BYTES + B_BYTES + (size_t)PC_NKV * 2 * PC_STAGE_BYTES +
                                 (2 * STAGES + 1 + 4 + 2 * PC_NKV + 1) * 8 + 16 + (1024 + 8 + 8 + PC_RPAD + 256 + 64) * 4 +
                                 PC_MAX_TILES * 16 + (PC_RPAD + 4) * 4 + 64;
};

struct PcMat {     // one weight matrix kind, static decomposition
  int T;           // row tiles (of 128)
  int KB;          // k-blocks (of 64)
  int S;           // K splits
  int G;           // tile groups = gridDim / S
};

struct PcParams {
  int n_layer, D, F, V, H, S_max, R, n_utts, kv_fp32;
  int n_prod;      // TMA producer threads (1 or 2): one issuing thread tops out at ~64 GB/s per SM (tools/micro/tma_bench)
  int ts;          // 1: copy each weight tile smem -> TMEM (tcgen05.cp) and run the UMMAs with A in tensor memory
  int pf_ahead;    // weight tiles prefetched into L2 ahead of the smem ring (0 = off)
  float eps;
  PcMat m_qkv, m_o, m_w13, m_w2, m_head;
  const __nv_bfloat16* attn_norm;   // layer 0; layer l at + l * layer_stride
  const __nv_bfloat16* ffn_norm;
  const __nv_bfloat16* out_norm;
  size_t layer_stride;              // elements between consecutive layers' tensors
  const __nv_bfloat16* tok_emb;
  const __nv_bfloat16* pos_emb;
  const float* spk_proj;
  float* x;        // [RPAD, D]   residual stream (red.add target of wo / w2)
  float* qkv;      // [RPAD, 3D]  red.add target, zeroed during the wo phase
  float* gu;       // [RPAD, 2F]  g | u, red.add target, zeroed during the attention phase
  float* logits;   // [2*max_utts, V] sampler buffer (zeroed by the sampler after use)
  float* part_o;   // [RPAD*H*PC_MAX_CHUNKS, 128]
  float* part_ml;  // [RPAD*H*PC_MAX_CHUNKS, 2]
  char* kv;
  size_t kv_half;  // bytes of one layer's K (or V) region
  unsigned* bar;   // grid-wide arrival counter (zero at launch)
  long long* trace;  // optional [gridDim][PC_TRACE_EVENTS] clock64 stamps of compute-thread 0 (debug)
  S1State st;
};

// ---- small device helpers -------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_inc(unsigned* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "l"(pol) : "memory");
}

// Pull a weight tile into L2 only (no smem, no barrier): HBM keeps streaming while the ring is full.
__device__ __forceinline__ void tma_prefetch_3d(const void* tmap, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// byte offset of the 16-byte chunk holding k..k+7 of activation row `row` inside the B operand
// (K-major, 128B swizzle, [k-block][NB rows][128 B])
template <int NB>
__device__ __forceinline__ uint32_t b_chunk_off(int kb_local, int row, int kchunk /* (k % 64) / 8 */) {
  return (uint32_t)(kb_local * (NB * 128) + (row >> 3) * 1024 + (row & 7) * 128 + ((kchunk ^ (row & 7)) << 4));
}

// split 8 fp32 into bf16 hi / lo and store both rows' chunks
template <int NB>
__device__ __forceinline__ void b_store8(uint8_t* B, int kb_local, int n, int kchunk, const float (&v)[8]) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * i]), h1 = __float2bfloat16_rn(v[2 * i + 1]);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(v[2 * i] - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16_rn(v[2 * i + 1] - __bfloat162float(h1));
    hi[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    lo[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  *reinterpret_cast<uint4*>(B + b_chunk_off<NB>(kb_local, n, kchunk)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(B + b_chunk_off<NB>(kb_local, NB / 2 + n, kchunk)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

struct PcSlice {   // what this CTA owns of one matrix
  int kb0, kb1, t0, nt, G;
};
__device__ __forceinline__ PcSlice pc_slice(const PcMat& m, int cta) {
  PcSlice s;
  const int sp = cta % m.S, g = cta / m.S;
  s.kb0 = (int)(((long long)sp * m.KB) / m.S);
  s.kb1 = (int)(((long long)(sp + 1) * m.KB) / m.S);
  s.G = m.G;
  s.t0 = g;
  s.nt = (g < m.G && g < m.T) ? (m.T - g + m.G - 1) / m.G : 0;
  return s;
}

// Attention work decomposition of one layer.  A unit = (activation row r, head h, split z): a contiguous run of
// KV tiles (ppc positions each) of that row's cache; every (r, h) is cut into nz <= zmax splits so that about one
// unit lands on every CTA at batch 1 while large batches get one multi-tile unit per (r, h).  Global unit ids are
// dealt round-robin to CTAs; the producer and the compute warps walk the identical sequence.
struct AttSplit {
  int tt, tps, nz;   // tiles of the row, tiles per split, number of splits
};
__device__ __forceinline__ AttSplit att_split(int L, int ppc, int zmax) {
  AttSplit a;
  a.tt = (L + ppc - 1) / ppc;
  const int zr = min(zmax, a.tt);
  a.tps = (a.tt + zr - 1) / zr;
  a.nz = (a.tt + a.tps - 1) / a.tps;
  return a;
}
// Per-step attention tile table of this CTA (identical for every layer; only the layer base pointer differs):
// built once by one thread, walked by the producer (KV loads) and by the compute warps.
struct AttTile {
  uint32_t off;     // byte offset of the tile inside one layer's K (or V) region
  uint32_t meta;    // npos | has_cur << 8 | unit_first << 9 | unit_last << 10 | owns_cur << 11
  uint32_t where;   // r | h << 8 | z << 16 | cache_row << 24
  int L;            // positions of the row incl. the current token
};
__device__ __forceinline__ int build_att_table(const PcParams& p, int cta, int G, int ppc, int esz, AttTile* tab, int* row_nz) {
  const int zmax = max(1, G / (p.R * p.H));
  int nt = 0, base = 0;
  int id = cta;
  for (int r = 0; r < p.R; ++r) {
    const int u = p.st.slot_map[r >> 1];
    const int L = p.st.pos[u] + 1;
    const AttSplit a = att_split(L, ppc, zmax);
    row_nz[r] = a.nz;
    const int units = p.H * a.nz;
    const int cr = 2 * u + (r & 1);
    while (id < base + units) {            // this CTA's units inside row r: ids cta, cta+G, ...
      const int idx = id - base;
      const int h = idx / a.nz, z = idx - h * a.nz;
      const int t1 = min(a.tt, (z + 1) * a.tps);
      for (int t = z * a.tps; t < t1; ++t) {
        if (nt >= PC_MAX_TILES) __trap();   // fail loudly rather than drop work
        const int p0 = t * ppc;
        const int npos = min(L - 1, p0 + ppc) - p0;
        AttTile e;
        e.off = (uint32_t)((((size_t)cr * p.H + h) * p.S_max + p0) * 128 * esz);
        e.meta = (uint32_t)npos | ((p0 + ppc >= L) ? 0x100u : 0u) | ((t == z * a.tps) ? 0x200u : 0u) |
                 ((t == t1 - 1) ? 0x400u : 0u) | ((t1 == a.tt) ? 0x800u : 0u);
        e.where = (uint32_t)r | ((uint32_t)h << 8) | ((uint32_t)z << 16) | ((uint32_t)cr << 24);
        e.L = L;
        tab[nt++] = e;
      }
      id += G;
    }
    base += units;
  }
  return nt;
}

template <bool KV_FP32>
__device__ __forceinline__ void load8s(const uint8_t* tile, int p, int sub, float (&v)[8]) {
  if (KV_FP32) {
    const float4* q = reinterpret_cast<const float4*>(tile + (size_t)p * 512 + sub * 32);
    const float4 a = q[0], b = q[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 w = *reinterpret_cast<const uint4*>(tile + (size_t)p * 256 + sub * 16);
    v[0] = bf_lo(w.x); v[1] = bf_hi(w.x); v[2] = bf_lo(w.y); v[3] = bf_hi(w.y);
    v[4] = bf_lo(w.z); v[5] = bf_hi(w.z); v[6] = bf_lo(w.w); v[7] = bf_hi(w.w);
  }
}

template <bool KV_FP32, int NB>
__global__ void __launch_bounds__(PC_THREADS, 1)
k_decode_persistent(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_o,
                    const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w3,
                    const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_head,
                    const PcParams p) {
  using Cfg = PcCfg<NB>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int RH = Cfg::RH;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem;
  uint8_t* Bop = ring + STAGES * PC_STAGE_BYTES;
  uint8_t* kvbuf = Bop + Cfg::B_BYTES;                       // [NKV][K tile | V tile]
  uint64_t* bars = reinterpret_cast<uint64_t*>(kvbuf + PC_NKV * 2 * PC_STAGE_BYTES);
  uint64_t* b_full = bars;
  uint64_t* b_empty = b_full + STAGES;
  uint64_t* b_ready = b_empty + STAGES;
  uint64_t* acc_full = b_ready + 1;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* kv_full = acc_empty + 2;
  uint64_t* kv_empty = kv_full + PC_NKV;
  uint64_t* tab_ready = kv_empty + PC_NKV;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tab_ready + 1);
  float* sm_f = reinterpret_cast<float*>(tmem_slot + 4);
  AttTile* att_tab = reinterpret_cast<AttTile*>(sm_f + (1024 + 8 + 8 + PC_RPAD + 256 + 64));
  int* row_nz = reinterpret_cast<int*>(att_tab + PC_MAX_TILES);   // [PC_RPAD] splits per row, then [1] tile count

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, G = gridDim.x;
  const int esz = KV_FP32 ? 4 : 2;
  const int ppc = PC_STAGE_BYTES / (128 * esz);             // positions per KV tile

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(ptx::smem_u32(b_full + s), 1);
      ptx::mbar_init(ptx::smem_u32(b_empty + s), 1);
    }
    ptx::mbar_init(ptx::smem_u32(b_ready), 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(ptx::smem_u32(acc_full + i), 1);
      ptx::mbar_init(ptx::smem_u32(acc_empty + i), 128);
    }
    for (int i = 0; i < PC_NKV; ++i) {
      ptx::mbar_init(ptx::smem_u32(kv_full + i), 1);
      ptx::mbar_init(ptx::smem_u32(kv_empty + i), 1);
    }
    ptx::mbar_init(ptx::smem_u32(tab_ready), 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  // zero the B operand once: rows >= R (hi and lo halves) must read as zero forever
  for (int i = tid; i < Cfg::B_BYTES / 16; i += PC_THREADS) reinterpret_cast<uint4*>(Bop)[i] = make_uint4(0, 0, 0, 0);
  fence_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // provably warp-uniform (no operand waterfall at the tcgen05 sites)
  pdl_launch_dependents();

  const PcSlice s_qkv = pc_slice(p.m_qkv, cta), s_o = pc_slice(p.m_o, cta), s_w13 = pc_slice(p.m_w13, cta),
                s_w2 = pc_slice(p.m_w2, cta), s_head = pc_slice(p.m_head, cta);
  const int T1 = p.m_w13.T >> 1;  // w1 tiles; tiles >= T1 belong to w3

  if (warp == 0 || warp == 6) {
    // =================================== PRODUCER(S) ================================================
    // Weight tile n of the flat schedule is issued by producer (n % n_prod); producer 0 also feeds the KV slots.
    // The whole warp walks the schedule (warp-uniform control flow and operands); one elected lane issues the TMA /
    // bulk-copy instructions.  Issuing them from a `lane == 0` branch makes the compiler wrap each UTMALDG in an
    // ELECT + 4 x R2UR.BROADCAST + vote loop, which capped one producer at a tile per ~0.25 us.
    const int pid = warp == 0 ? 0 : 1;
    const int NPROD = p.n_prod;
    if (pid < NPROD) {
      const uint64_t pol = ptx::policy_evict_first();
      uint32_t slot = 0, kv_ctr = 0;
      // Flat view of this CTA's weight-tile schedule (layer-major: qkv, wo, w1|w3, w2; then the head), used by the
      // L2 prefetch cursor that runs `pf_ahead` tiles in front of the smem ring.
      const int nk_q = s_qkv.kb1 - s_qkv.kb0, nk_o = s_o.kb1 - s_o.kb0, nk_f = s_w13.kb1 - s_w13.kb0,
                nk_2 = s_w2.kb1 - s_w2.kb0, nk_h = s_head.kb1 - s_head.kb0;
      const int c_q = s_qkv.nt * nk_q, c_o = s_o.nt * nk_o, c_f = s_w13.nt * nk_f, c_2 = s_w2.nt * nk_2;
      const int per_layer = c_q + c_o + c_f + c_2;
      const int total_tiles = p.n_layer * per_layer + s_head.nt * nk_h;
      auto prefetch_flat = [&](int n) {
        if (n >= total_tiles) return;
        const CUtensorMap* tm;
        const PcSlice* sl;
        int nk, layer = 0, m;
        if (n >= p.n_layer * per_layer) {
          m = n - p.n_layer * per_layer; tm = &tm_head; sl = &s_head; nk = nk_h;
        } else {
          layer = n / per_layer;
          m = n - layer * per_layer;
          if (m < c_q) { tm = &tm_qkv; sl = &s_qkv; nk = nk_q; }
          else if (m < c_q + c_o) { m -= c_q; tm = &tm_o; sl = &s_o; nk = nk_o; }
          else if (m < c_q + c_o + c_f) { m -= c_q + c_o; tm = &tm_w1; sl = &s_w13; nk = nk_f; }
          else { m -= c_q + c_o + c_f; tm = &tm_w2; sl = &s_w2; nk = nk_2; }
        }
        const int ti = m / nk, kb = sl->kb0 + (m - ti * nk);
        int t = sl->t0 + ti * sl->G;
        if (sl == &s_w13 && t >= T1) { t -= T1; tm = &tm_w3; }
        if (ptx::elect_one()) tma_prefetch_3d(tm, kb * 64, t * 128, layer);
        __syncwarp();
      };
      int issued = 0, pf_next = 0;   // prefetch only while the ring is full: idle producer time -> HBM keeps streaming into L2
      auto gemm_tiles = [&](const CUtensorMap* tmA, const CUtensorMap* tmB2, int split_t, const PcSlice& sl, int layer) {
        for (int i = 0; i < sl.nt; ++i) {
          const int t = sl.t0 + i * sl.G;
          const CUtensorMap* tm = (t < split_t) ? tmA : tmB2;
          const int tt = (t < split_t) ? t : t - split_t;
          for (int kb = sl.kb0; kb < sl.kb1; ++kb) {
            const uint32_t s = slot % STAGES, ph = (slot / STAGES) & 1u;
            ++slot;
            if (NPROD > 1 && (issued % NPROD) != pid) { ++issued; continue; }   // the other producer's tile
            const uint32_t eb = ptx::smem_u32(b_empty + s);
            if (p.pf_ahead > 0 && !ptx::mbar_test_wait(eb, ph ^ 1u)) {
              // the ring is full: spend the idle time pulling upcoming tiles into L2 (non-blocking probe)
              if (pf_next <= issued) pf_next = issued + NPROD;      // never prefetch a tile that is about to be loaded
              while (pf_next < issued + 1 + p.pf_ahead && pf_next < total_tiles && !ptx::mbar_test_wait(eb, ph ^ 1u)) {
                prefetch_flat(pf_next);
                pf_next += NPROD;
              }
            }
            ptx::mbar_wait(eb, ph ^ 1u);
            ++issued;
            if (ptx::elect_one()) {
              const uint32_t full = ptx::smem_u32(b_full + s);
              ptx::mbar_arrive_expect_tx(full, PC_STAGE_BYTES);
              tma_load_3d(ptx::smem_u32(ring + (size_t)s * PC_STAGE_BYTES), tm, full, kb * 64, tt * 128, layer, pol);
            }
            __syncwarp();
          }
        }
      };
      // KV tiles [from, to) (in this CTA's table order) of layer l; positions < pos come from earlier steps
      auto kv_units = [&](int l, int from, int to) {
        const char* kbase = p.kv + (size_t)l * 2 * p.kv_half;
        const char* vbase = kbase + p.kv_half;
        const int n_tiles = row_nz[PC_RPAD];
        int idx = 0;
        for (int i = 0; i < n_tiles; ++i) {
          const AttTile e = att_tab[i];
          const int npos = (int)(e.meta & 0xffu);
          if (npos == 0) continue;                 // only the current position: it comes from registers
          const int my = idx++;
          if (my < from) continue;
          if (my >= to) break;
          const uint32_t ks = kv_ctr % PC_NKV, ph = (kv_ctr / PC_NKV) & 1u;
          ++kv_ctr;
          ptx::mbar_wait(ptx::smem_u32(kv_empty + ks), ph ^ 1u);
          if (ptx::elect_one()) {
            const uint32_t bytes = (uint32_t)npos * 128 * esz;
            const uint32_t full = ptx::smem_u32(kv_full + ks);
            ptx::mbar_arrive_expect_tx(full, 2 * bytes);
            uint8_t* dst = kvbuf + (size_t)ks * 2 * PC_STAGE_BYTES;
            bulk_load(ptx::smem_u32(dst), kbase + e.off, bytes, full);
            bulk_load(ptx::smem_u32(dst + PC_STAGE_BYTES), vbase + e.off, bytes, full);
          }
          __syncwarp();
        }
      };
      for (int l = 0; l < p.n_layer; ++l) {
        if (l == 0) {
          gemm_tiles(&tm_qkv, &tm_qkv, 1 << 30, s_qkv, 0);   // weights first: they need nothing from the previous kernel
          if (pid == 0) {
            ptx::mbar_wait(ptx::smem_u32(tab_ready), 0);      // tile table built (after the PDL wait) by the compute warps
            kv_units(0, 0, 1 << 30);
          }
        } else if (pid == 0) {
          kv_units(l, PC_NKV, 1 << 30);                       // (units beyond the prefetched ones)
        }
        gemm_tiles(&tm_o, &tm_o, 1 << 30, s_o, l);
        gemm_tiles(&tm_w1, &tm_w3, T1, s_w13, l);
        gemm_tiles(&tm_w2, &tm_w2, 1 << 30, s_w2, l);
        if (l + 1 < p.n_layer) {
          if (pid == 0) kv_units(l + 1, 0, PC_NKV);           // next layer's first KV tiles ride ahead of its QKV GEMM
          gemm_tiles(&tm_qkv, &tm_qkv, 1 << 30, s_qkv, l + 1);
        }
      }
      gemm_tiles(&tm_head, &tm_head, 1 << 30, s_head, 0);
    }
  } else if (warp == 1) {
    // =================================== MMA ISSUER =================================================
    // The WHOLE warp walks the schedule (waits, counters, descriptor arithmetic stay warp-uniform -> uniform
    // registers) and one elected lane issues the tcgen05 instructions.  Issuing from a `lane == 0` branch makes the
    // compiler wrap every UTCHMMA in an R2UR / ELECT / vote loop: 175 instead of 134 cycles per instruction
    // (tools/micro/umma_bench.cu).
    {
      const uint32_t idesc = ptx::umma_idesc_bf16(128, NB);
      uint32_t slot = 0, tile_ctr = 0, bphase = 0, a_ctr = 0;
      auto gemm_phase = [&](const PcSlice& sl) {
        if (sl.nt == 0) return;
        ptx::mbar_wait(ptx::smem_u32(b_ready), bphase & 1u);   // B operand of this phase staged
        ++bphase;
        ptx::tc_fence_after();
        for (int i = 0; i < sl.nt; ++i) {
          const uint32_t ab = tile_ctr & 1u, aph = (tile_ctr >> 1) & 1u;
          ptx::mbar_wait(ptx::smem_u32(acc_empty + ab), aph ^ 1u);   // epilogue drained this accumulator
          ptx::tc_fence_after();
          const uint32_t dcol = tmem_base + ab * NB;
          for (int kb = sl.kb0; kb < sl.kb1; ++kb) {
            const uint32_t s = slot % STAGES, ph = (slot / STAGES) & 1u;
            ++slot;
            ptx::mbar_wait(ptx::smem_u32(b_full + s), ph);
            ptx::tc_fence_after();
            const uint32_t a_addr = ptx::smem_u32(ring + (size_t)s * PC_STAGE_BYTES);
            const uint64_t ad = ptx::umma_desc_k_sw128(a_addr);
            const uint64_t bd = ptx::umma_desc_k_sw128(ptx::smem_u32(Bop + (size_t)(kb - sl.kb0) * (NB * 128)));
            const uint32_t first = (uint32_t)(kb != sl.kb0);
            if (p.ts) {
              // Experimental (MVB_PC_TS=1, parity-tested, not faster): copy the tile to tensor memory with tcgen05.cp,
              // release the smem stage once the copy is done, run the UMMAs with A in TMEM.
              const uint32_t abuf = tmem_base + Cfg::A_COL0 + (a_ctr & 3u) * 32;
              ++a_ctr;
              if (ptx::elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) ptx::tmem_cp_128x256b(abuf + 8 * k, ad + 2 * k);
                ptx::umma_commit(ptx::smem_u32(b_empty + s));
#pragma unroll
                for (int k = 0; k < 4; ++k) ptx::umma_bf16_ts(dcol, abuf + 8 * k, bd + 2 * k, idesc, first | (uint32_t)(k != 0));
              }
            } else if (ptx::elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) ptx::umma_bf16(dcol, ad + 2 * k, bd + 2 * k, idesc, first | (uint32_t)(k != 0));
              ptx::umma_commit(ptx::smem_u32(b_empty + s));
            }
            __syncwarp();
          }
          if (ptx::elect_one()) ptx::umma_commit(ptx::smem_u32(acc_full + ab));
          __syncwarp();
          ++tile_ctr;
        }
      };
      for (int l = 0; l < p.n_layer; ++l) {
        gemm_phase(s_qkv);
        gemm_phase(s_o);
        gemm_phase(s_w13);
        gemm_phase(s_w2);
      }
      gemm_phase(s_head);
    }
  } else {
    // =================================== COMPUTE WARPS ==============================================
    const int ct = tid - 64;            // 0..127
    const int cw = ct >> 5;             // 0..3
    const int quad = warp & 3;          // TMEM lane quadrant this warp may read
    float* sm_o = sm_f;                 // [8][128]
    float* sm_m = sm_f + 1024;          // [8]
    float* sm_l = sm_m + 8;             // [8]
    float* sm_rs = sm_l + 8;            // [PC_RPAD]
    float* sm_cur = sm_rs + PC_RPAD;    // [2][128] k, v of the current token
    float* sm_red = sm_cur + 256;       // [PC_RPAD][4] partial sums of squares
    uint32_t tile_ctr = 0, bar_idx = 0, kv_ctr = 0;
    pdl_wait();                         // state / x inputs of the previous kernels are visible
    if (ct == 0) {
      row_nz[PC_RPAD] = build_att_table(p, cta, G, ppc, esz, att_tab, row_nz);
      ptx::mbar_arrive(ptx::smem_u32(tab_ready));   // release: the producer may walk the table
    }
    compute_sync();
    int ev = 0;
    auto stamp = [&]() {
      if (p.trace != nullptr && ct == 0 && ev < PC_TRACE_EVENTS) p.trace[(size_t)cta * PC_TRACE_EVENTS + ev] = clock64();
      ++ev;
    };
    stamp();

    auto grid_arrive = [&]() {          // bar.sync orders every compute thread's writes before the release
      compute_sync();
      if (ct == 0) red_release_inc(p.bar);
      ++bar_idx;
    };
    auto grid_wait = [&]() {            // wait for arrival #bar_idx of every CTA
      if (ct == 0) {
        const unsigned target = bar_idx * (unsigned)G;
        const long long t0 = clock64();
        while (ld_acquire_u32(p.bar) < target) {
          if (clock64() - t0 > 4000000000ll) __trap();
        }
      }
      compute_sync();
    };
    auto b_publish = [&]() {            // generic-proxy smem writes -> visible to the tensor core
      fence_async_smem();
      compute_sync();
      if (ct == 0) ptx::mbar_arrive(ptx::smem_u32(b_ready));
    };
    // B <- hi/lo split of (x * rstd) * gain over this CTA's K range; RMSNorm statistics (fast_model.py:254-255)
    // and the chunk loads are issued together so the phase costs ONE L2 round trip.
    auto stage_norm = [&](const PcSlice& sl, const __nv_bfloat16* gain) {
      if (sl.nt == 0) return;
      const int nchunk = (sl.kb1 - sl.kb0) * 8;
      const int total = p.R * nchunk;
      constexpr int MAXC = 2;             // chunks pre-loaded per thread (covers R = 2 completely: 176 chunks)
      float4 ca[MAXC], cb[MAXC];
      uint4 cg[MAXC];
#pragma unroll
      for (int j = 0; j < MAXC; ++j) {
        const int i = ct + j * 128;
        if (i < total) {
          const int n = i / nchunk, c = i - n * nchunk;
          const float4* xr = reinterpret_cast<const float4*>(p.x + (size_t)n * p.D + sl.kb0 * 64 + c * 8);
          ca[j] = __ldcg(xr);
          cb[j] = __ldcg(xr + 1);
          cg[j] = *reinterpret_cast<const uint4*>(gain + sl.kb0 * 64 + c * 8);
        }
      }
      const int f4_per_row = p.D >> 2;
      for (int n0 = 0; n0 < p.R; n0 += 4) {
        float ss[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int n = n0 + q;
          if (n < p.R) {
            const float4* xr = reinterpret_cast<const float4*>(p.x + (size_t)n * p.D);
            for (int i = ct; i < f4_per_row; i += 128) {
              const float4 v = __ldcg(xr + i);
              ss[q] = fmaf(v.x, v.x, ss[q]); ss[q] = fmaf(v.y, v.y, ss[q]);
              ss[q] = fmaf(v.z, v.z, ss[q]); ss[q] = fmaf(v.w, v.w, ss[q]);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float s = warp_sum(ss[q]);
          if (lane == 0 && n0 + q < p.R) sm_red[(n0 + q) * 4 + cw] = s;
        }
      }
      compute_sync();
      if (ct < p.R) {
        const float t = sm_red[ct * 4] + sm_red[ct * 4 + 1] + sm_red[ct * 4 + 2] + sm_red[ct * 4 + 3];
        sm_rs[ct] = rsqrtf(t / (float)p.D + p.eps);
      }
      compute_sync();
      auto emit = [&](int i, const float4& a, const float4& b, const uint4& gw) {
        const int n = i / nchunk, c = i - n * nchunk;
        const float rs = sm_rs[n];
        float v[8] = {(a.x * rs) * bf_lo(gw.x), (a.y * rs) * bf_hi(gw.x), (a.z * rs) * bf_lo(gw.y), (a.w * rs) * bf_hi(gw.y),
                      (b.x * rs) * bf_lo(gw.z), (b.y * rs) * bf_hi(gw.z), (b.z * rs) * bf_lo(gw.w), (b.w * rs) * bf_hi(gw.w)};
        b_store8<NB>(Bop, c >> 3, n, c & 7, v);
      };
#pragma unroll
      for (int j = 0; j < MAXC; ++j)
        if (ct + j * 128 < total) emit(ct + j * 128, ca[j], cb[j], cg[j]);
      // larger batches: 2 chunks (6 loads) in flight per thread per trip
      for (int i0 = ct + MAXC * 128; i0 < total; i0 += 2 * 128) {
        float4 la[2], lb[2];
        uint4 lg[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int i = i0 + j * 128;
          if (i < total) {
            const int n = i / nchunk, c = i - n * nchunk;
            const float4* xr = reinterpret_cast<const float4*>(p.x + (size_t)n * p.D + sl.kb0 * 64 + c * 8);
            la[j] = __ldcg(xr);
            lb[j] = __ldcg(xr + 1);
            lg[j] = *reinterpret_cast<const uint4*>(gain + sl.kb0 * 64 + c * 8);
          }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
          if (i0 + j * 128 < total) emit(i0 + j * 128, la[j], lb[j], lg[j]);
      }
      b_publish();
    };
    // epilogue of this CTA's tiles: TMEM -> red.add into out[n][col0 + row]
    auto epilogue = [&](const PcSlice& sl, float* out, int ldo, int M, int split_t, float* out2) {
      for (int i = 0; i < sl.nt; ++i) {
        const int t = sl.t0 + i * sl.G;
        const uint32_t ab = tile_ctr & 1u, aph = (tile_ctr >> 1) & 1u;
        ++tile_ctr;
        ptx::mbar_wait(ptx::smem_u32(acc_full + ab), aph);
        ptx::tc_fence_after();
        uint32_t acc[NB];
        const uint32_t ta = tmem_base + ((uint32_t)(32 * quad) << 16) + ab * NB;
        if (NB == 16) {
          ptx::tmem_ld16(ta, *reinterpret_cast<uint32_t(*)[16]>(&acc[0]));
        } else {
          ptx::tmem_ld16(ta, *reinterpret_cast<uint32_t(*)[16]>(&acc[0]));
          ptx::tmem_ld16(ta + 16, *reinterpret_cast<uint32_t(*)[16]>(&acc[NB - 16]));
        }
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(ptx::smem_u32(acc_empty + ab));
        float* o = (t < split_t) ? out : out2;
        const int j = ((t < split_t) ? t : t - split_t) * 128 + 32 * quad + lane;
        if (j < M) {
#pragma unroll
          for (int n = 0; n < RH; ++n)
            if (n < p.R) atomicAdd(o + (size_t)n * ldo + j, __uint_as_float(acc[n]) + __uint_as_float(acc[RH + n]));
        }
      }
    };
    auto zero_slice = [&](float* buf, size_t n_floats) {   // this CTA's share of a buffer
      const size_t per = (n_floats / 4 + G - 1) / G;
      const size_t b = (size_t)cta * per, e = min(n_floats / 4, b + per);
      for (size_t i = b + ct; i < e; i += 128) reinterpret_cast<float4*>(buf)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    };

    // ---- phase E: x = tok_emb + pos_emb + speaker (fast_model.py:152-157); CTA r builds row r
    if (cta < p.R) {
      const int u = p.st.slot_map[cta >> 1], c = cta & 1;
      const __nv_bfloat16* te = p.tok_emb + (size_t)p.st.row_tok[2 * u + c] * p.D;
      const __nv_bfloat16* pe = p.pos_emb + (size_t)p.st.pos[u] * p.D;
      for (int d = ct; d < p.D; d += 128) {
        float v = bf16_to_f32(te[d]) + bf16_to_f32(pe[d]);
        if (c == 0) v += p.spk_proj[(size_t)u * p.D + d];
        p.x[(size_t)cta * p.D + d] = v;
      }
    }
    stamp();
    grid_arrive();

    for (int l = 0; l < p.n_layer; ++l) {
      const size_t lo_ = (size_t)l * p.layer_stride;
      // ---- QKV: B = RMSNorm(x) * attn_norm; out: qkv (zero on entry)
      grid_wait();
      stamp();
      stage_norm(s_qkv, p.attn_norm + lo_);
      stamp();
      epilogue(s_qkv, p.qkv, 3 * p.D, 3 * p.D, 1 << 30, nullptr);
      stamp();
      grid_arrive();

      // ---- attention over [0, pos] + KV-cache append (fast_model.py:104-113, 220-224)
      zero_slice(p.gu, (size_t)PC_RPAD * 2 * p.F);   // free since the previous layer's w2 phase; ordered by the next arrive
      grid_wait();
      stamp();
      {
        char* kbase = p.kv + (size_t)l * 2 * p.kv_half;
        char* vbase = kbase + p.kv_half;
        const int half = lane >> 4, sub = lane & 15;
        const int hw = cw * 2 + half;               // half-warp id 0..7
        const unsigned hmask = half ? 0xffff0000u : 0x0000ffffu;
        float q[8], o[8], m = -INFINITY, lsum = 0.f, kcur = 0.f, vcur = 0.f;
        const int n_tiles = row_nz[PC_RPAD];
        for (int ti = 0; ti < n_tiles; ++ti) {
          const AttTile e = att_tab[ti];
          const int npos = (int)(e.meta & 0xffu);
          const bool has_cur = e.meta & 0x100u, unit_first = e.meta & 0x200u, unit_last = e.meta & 0x400u,
                     owns_cur = e.meta & 0x800u;
          const int r = e.where & 0xff, h = (e.where >> 8) & 0xff, z = (e.where >> 16) & 0xff, cr = e.where >> 24;
          const int L = e.L;
          if (unit_first) {
            // ---- unit begin: one round trip for everything the unit needs from global memory
            const float* qrow = p.qkv + (size_t)r * 3 * p.D + h * 128;
            const float4 qa = __ldcg(reinterpret_cast<const float4*>(qrow + sub * 8));
            const float4 qb = __ldcg(reinterpret_cast<const float4*>(qrow + sub * 8) + 1);
            if (owns_cur) {
              kcur = __ldcg(qrow + p.D + ct);
              vcur = __ldcg(qrow + 2 * p.D + ct);
            }
            const float sc = 0.08838834764831845f;   // 1/sqrt(128)
            q[0] = qa.x * sc; q[1] = qa.y * sc; q[2] = qa.z * sc; q[3] = qa.w * sc;
            q[4] = qb.x * sc; q[5] = qb.y * sc; q[6] = qb.z * sc; q[7] = qb.w * sc;
            m = -INFINITY; lsum = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = 0.f;
          }
          // ---- one KV tile (and, on the row's last tile, the current token)
          if (has_cur) {
            // append the new token's k, v to the cache, rounded as the cache stores them; share them via smem
            const size_t ce = (((size_t)cr * p.H + h) * p.S_max + (L - 1)) * 128 + ct;
            if (KV_FP32) {
              reinterpret_cast<float*>(kbase)[ce] = kcur;
              reinterpret_cast<float*>(vbase)[ce] = vcur;
            } else {
              const __nv_bfloat16 kb16 = __float2bfloat16_rn(kcur), vb16 = __float2bfloat16_rn(vcur);
              reinterpret_cast<__nv_bfloat16*>(kbase)[ce] = kb16;
              reinterpret_cast<__nv_bfloat16*>(vbase)[ce] = vb16;
              kcur = __bfloat162float(kb16);
              vcur = __bfloat162float(vb16);
            }
            sm_cur[ct] = kcur;
            sm_cur[128 + ct] = vcur;
          }
          if (npos > 0) {
            const uint32_t ks = kv_ctr % PC_NKV, ph = (kv_ctr / PC_NKV) & 1u;
            ++kv_ctr;
            ptx::mbar_wait(ptx::smem_u32(kv_full + ks), ph);
            const uint8_t* kt = kvbuf + (size_t)ks * 2 * PC_STAGE_BYTES;
            const uint8_t* vt = kt + PC_STAGE_BYTES;
            // each half-warp owns positions hw, hw+8, ...; two positions per trip for ILP
            for (int pb = hw; pb < npos; pb += 16) {
              const int pA = pb, pB = pb + 8;
              const bool vB = pB < npos;
              float ka[8], kb2[8], sA = 0.f, sB = 0.f;
              load8s<KV_FP32>(kt, pA, sub, ka);
              if (vB) load8s<KV_FP32>(kt, pB, sub, kb2);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                sA = fmaf(q[i], ka[i], sA);
                if (vB) sB = fmaf(q[i], kb2[i], sB);
              }
#pragma unroll
              for (int off = 8; off > 0; off >>= 1) {   // reduce inside the 16-lane group only: trip counts differ per half-warp
                sA += __shfl_xor_sync(hmask, sA, off);
                sB += __shfl_xor_sync(hmask, sB, off);
              }
              load8s<KV_FP32>(vt, pA, sub, ka);
              if (vB) load8s<KV_FP32>(vt, pB, sub, kb2);
              const float mn = fmaxf(m, vB ? fmaxf(sA, sB) : sA);
              const float corr = __expf(m - mn), wA = __expf(sA - mn), wB = vB ? __expf(sB - mn) : 0.f;
              lsum = lsum * corr + wA + wB;
#pragma unroll
              for (int i = 0; i < 8; ++i) o[i] = o[i] * corr + wA * ka[i] + (vB ? wB * kb2[i] : 0.f);
              m = mn;
            }
          }
          if (npos > 0 || has_cur) compute_sync();   // KV tile fully consumed; sm_cur visible
          if (npos > 0 && ct == 0) ptx::mbar_arrive(ptx::smem_u32(kv_empty + ((kv_ctr - 1) % PC_NKV)));
          if (has_cur && hw == 0) {
            float sdot = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) sdot = fmaf(q[i], sm_cur[sub * 8 + i], sdot);
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) sdot += __shfl_xor_sync(0x0000ffffu, sdot, off);
            const float mn = fmaxf(m, sdot), corr = __expf(m - mn), pw = __expf(sdot - mn);
            lsum = lsum * corr + pw;
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = o[i] * corr + pw * sm_cur[128 + sub * 8 + i];
            m = mn;
          }
          if (unit_last) {
            // ---- unit end: merge the 8 half-warp states -> one partial (m, l, o[128]) for split z
            if (sub == 0) { sm_m[hw] = m; sm_l[hw] = lsum; }
#pragma unroll
            for (int i = 0; i < 8; ++i) sm_o[hw * 128 + sub * 8 + i] = o[i];
            compute_sync();
            {
              float M = -INFINITY;
#pragma unroll
              for (int i = 0; i < 8; ++i) M = fmaxf(M, sm_m[i]);
              float Ls = 0.f, O = 0.f;
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (sm_m[i] > -INFINITY) {
                  const float w = __expf(sm_m[i] - M);
                  Ls += sm_l[i] * w;
                  O += sm_o[i * 128 + ct] * w;
                }
              const size_t pidx = ((size_t)r * p.H + h) * PC_MAX_CHUNKS + z;
              p.part_o[pidx * 128 + ct] = O;
              if (ct == 0) { p.part_ml[pidx * 2] = M; p.part_ml[pidx * 2 + 1] = Ls; }
            }
            compute_sync();   // sm_o / sm_m / sm_cur are reused by the next unit
          }
        }
      }
      stamp();
      grid_arrive();

      // ---- wo + residual: B = merged attention output (columns = this CTA's K range of heads)
      grid_wait();
      stamp();
      if (s_o.nt > 0) {
        const int nchunk = (s_o.kb1 - s_o.kb0) * 8;
        for (int i = ct; i < p.R * nchunk; i += 128) {
          const int n = i / nchunk, c = i - n * nchunk;
          const int k = s_o.kb0 * 64 + c * 8;
          const int h = k >> 7, d0 = k & 127;
          const int nch = row_nz[n];
          const size_t pb = ((size_t)n * p.H + h) * PC_MAX_CHUNKS;
          float M = -INFINITY;
          float den = 0.f, v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = 0.f;
          if (nch <= 4) {
            // one round trip: (m, l) and o of every split are requested together
            float2 ml[4];
            float4 oa[4], ob[4];
#pragma unroll
            for (int z = 0; z < 4; ++z)
              if (z < nch) {
                ml[z] = __ldcg(reinterpret_cast<const float2*>(p.part_ml + (pb + z) * 2));
                const float4* po = reinterpret_cast<const float4*>(p.part_o + (pb + z) * 128 + d0);
                oa[z] = __ldcg(po);
                ob[z] = __ldcg(po + 1);
              }
#pragma unroll
            for (int z = 0; z < 4; ++z)
              if (z < nch) M = fmaxf(M, ml[z].x);
#pragma unroll
            for (int z = 0; z < 4; ++z)
              if (z < nch) {
                const float w = __expf(ml[z].x - M);
                den += ml[z].y * w;
                v[0] += oa[z].x * w; v[1] += oa[z].y * w; v[2] += oa[z].z * w; v[3] += oa[z].w * w;
                v[4] += ob[z].x * w; v[5] += ob[z].y * w; v[6] += ob[z].z * w; v[7] += ob[z].w * w;
              }
          } else {
            for (int z = 0; z < nch; ++z) M = fmaxf(M, __ldcg(p.part_ml + (pb + z) * 2));
            for (int z = 0; z < nch; ++z) {
              const float w = __expf(__ldcg(p.part_ml + (pb + z) * 2) - M);
              den += __ldcg(p.part_ml + (pb + z) * 2 + 1) * w;
              const float4* po = reinterpret_cast<const float4*>(p.part_o + (pb + z) * 128 + d0);
              const float4 a = __ldcg(po), b = __ldcg(po + 1);
              v[0] += a.x * w; v[1] += a.y * w; v[2] += a.z * w; v[3] += a.w * w;
              v[4] += b.x * w; v[5] += b.y * w; v[6] += b.z * w; v[7] += b.w * w;
            }
          }
          const float inv = 1.f / den;
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] *= inv;
          b_store8<NB>(Bop, c >> 3, n, c & 7, v);
        }
        b_publish();
      }
      stamp();
      epilogue(s_o, p.x, p.D, p.D, 1 << 30, nullptr);
      stamp();
      grid_arrive();
      zero_slice(p.qkv, (size_t)PC_RPAD * 3 * p.D);  // q, k, v were consumed by the attention phase; ordered by the next arrive

      // ---- w1 | w3: B = RMSNorm(x) * ffn_norm; out: g | u (zero on entry)
      grid_wait();
      stamp();
      stage_norm(s_w13, p.ffn_norm + lo_);
      stamp();
      epilogue(s_w13, p.gu, 2 * p.F, p.F, T1, p.gu + p.F);
      stamp();
      grid_arrive();

      // ---- w2 + residual: B = silu(g) * u (fast_model.py:237)
      grid_wait();
      stamp();
      if (s_w2.nt > 0) {
        const int nchunk = (s_w2.kb1 - s_w2.kb0) * 8;
        const int total = p.R * nchunk;
        for (int i0 = ct; i0 < total; i0 += 2 * 128) {
          float4 g0[2], g1[2], u0[2], u1[2];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int i = i0 + j * 128;
            if (i < total) {
              const int n = i / nchunk, c = i - n * nchunk;
              const int k = s_w2.kb0 * 64 + c * 8;
              const float4* gp = reinterpret_cast<const float4*>(p.gu + (size_t)n * 2 * p.F + k);
              const float4* up = reinterpret_cast<const float4*>(p.gu + (size_t)n * 2 * p.F + p.F + k);
              g0[j] = __ldcg(gp); g1[j] = __ldcg(gp + 1); u0[j] = __ldcg(up); u1[j] = __ldcg(up + 1);
            }
          }
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int i = i0 + j * 128;
            if (i < total) {
              const int n = i / nchunk, c = i - n * nchunk;
              const float g[8] = {g0[j].x, g0[j].y, g0[j].z, g0[j].w, g1[j].x, g1[j].y, g1[j].z, g1[j].w};
              const float uu[8] = {u0[j].x, u0[j].y, u0[j].z, u0[j].w, u1[j].x, u1[j].y, u1[j].z, u1[j].w};
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = (g[e] / (1.f + __expf(-g[e]))) * uu[e];
              b_store8<NB>(Bop, c >> 3, n, c & 7, v);
            }
          }
        }
        b_publish();
      }
      stamp();
      epilogue(s_w2, p.x, p.D, p.D, 1 << 30, nullptr);
      stamp();
      grid_arrive();
    }
    // ---- head: logits += RMSNorm(x) * out_norm . W_out^T   (rows n = batch order = sampler rows 2u, 2u+1)
    grid_wait();
    zero_slice(p.gu, (size_t)PC_RPAD * 2 * p.F);
    stamp();
    stage_norm(s_head, p.out_norm);
    stamp();
    epilogue(s_head, p.logits, p.V, p.V, 1 << 30, nullptr);
    stamp();
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace mvb
