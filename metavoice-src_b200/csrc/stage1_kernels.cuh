// Stage-1 decode-path kernels, CUDA-core generation ("path A"): one kernel per operator, fp32
// activations, bf16 weights streamed once with 128-bit loads.  Each kernel cites the reference
// operator it replaces (metavoiceio/metavoice-src @ de3fa211).
//
// Row convention: utterance slot u owns activation rows 2u (speaker-conditioned) and 2u+1
// (unconditioned) -- the CFG pair of fam/llm/fast_model.py:132-134.
#pragma once
#include "common.cuh"

namespace mvb {

// ---------------------------------------------------------------------------------------------
// Device-resident decode state (lives in the caller's workspace).  Everything the step needs is
// read from here so one captured CUDA graph can be replayed for every token.
struct SamplingDev {
  float guidance, temperature, top_p;
  int top_k, end_of_audio;
  unsigned long long seed;
};

struct S1State {
  int* slot_map;        // [max_utts] logical batch index -> utterance slot
  int* pos;             // [max_utts] cache position the current token is written to
  int* row_tok;         // [2*max_utts] token fed to each row this step
  int* done;            // [max_utts] end-of-audio latch (fast_inference_utils.py:161)
  int* n_gen;           // [max_utts] tokens produced so far
  int* gen_tokens;      // [max_utts, max_new] tokens fed back (== generate()'s appended tokens)
  int* sampled_tokens;  // [max_utts, max_new] tokens the sampler drew (differs only under teacher forcing)
  SamplingDev* samp;    // [max_utts]
  const float** noise;  // [max_utts] optional Exp(1) draws [max_new, vocab]
  const int** forced;   // [max_utts] optional teacher-forced tokens [max_new]
  unsigned* attn_ticket;  // [2*max_utts*n_head] split-KV arrival counters
  int* budget;          // [max_utts] tokens this utterance may produce (generate(): min(T + max_new, block) - T, utils:196-204)
  int* noise_base;      // [max_utts] generation step that row 0 of noise[u] belongs to (host-staged noise arrives in chunks)
  int max_new, block_size;
};

// ---------------------------------------------------------------------------------------------
// spk_proj[u] = W_spk . spk_emb   (fast_model.py:155 speaker_cond_pos, hoisted out of the step)
__global__ void __launch_bounds__(256) k_spk_proj(const __nv_bfloat16* __restrict__ W, const float* __restrict__ spk,
                                                  float* __restrict__ out, int D, int SD) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= D) return;
  float acc = 0.f;
  for (int k = lane; k < SD; k += 32) acc = fmaf(bf16_to_f32(W[(size_t)warp * SD + k]), spk[k], acc);
  acc = warp_sum(acc);
  if (lane == 0) out[warp] = acc;
}

// Install the inputs of one forward position: == the (idx[:, s], input_pos[s]) arguments of
// Transformer.forward (fast_model.py:150).
__global__ void k_set_input(S1State st, int utt, const int* __restrict__ idx, int S, int s, int pos) {
  if (threadIdx.x == 0) {
    st.slot_map[0] = utt;
    st.pos[utt] = pos;
    st.row_tok[2 * utt] = idx[s];
    st.row_tok[2 * utt + 1] = idx[S + s];
  }
}

__global__ void k_begin(S1State st, int utt, int first_token, int pos, SamplingDev sp, const float* noise,
                        const int* forced, int has_first, int budget) {
  if (threadIdx.x == 0) {
    st.samp[utt] = sp;
    st.noise[utt] = noise;
    st.noise_base[utt] = 0;
    st.forced[utt] = forced;
    st.done[utt] = 0;
    st.n_gen[utt] = 0;
    st.pos[utt] = pos;
    st.budget[utt] = budget;
    if (has_first) {
      st.row_tok[2 * utt] = first_token;
      st.row_tok[2 * utt + 1] = first_token;
    }
  }
}

__global__ void k_identity_slots(S1State st, int n) {
  if ((int)threadIdx.x < n) st.slot_map[threadIdx.x] = threadIdx.x;
}

// ---------------------------------------------------------------------------------------------
// x = tok_emb[idx] + pos_emb[input_pos] + speaker_cond * mask      (fast_model.py:152-157)
// The reference has no RoPE: positions are a learned table.
__global__ void __launch_bounds__(256) k_embed(S1State st, const __nv_bfloat16* __restrict__ tok_emb,
                                               const __nv_bfloat16* __restrict__ pos_emb,
                                               const float* __restrict__ spk_proj, float* __restrict__ x, int D) {
  pdl_launch_dependents();
  pdl_wait();
  const int u = st.slot_map[blockIdx.y];
  const int c = blockIdx.x;  // 0 = conditioned row, 1 = unconditioned row
  const int r = 2 * u + c;
  const int tok = st.row_tok[r];
  const int pos = st.pos[u];
  const __nv_bfloat16* te = tok_emb + (size_t)tok * D;
  const __nv_bfloat16* pe = pos_emb + (size_t)pos * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float v = bf16_to_f32(te[d]) + bf16_to_f32(pe[d]);
    if (c == 0) v += spk_proj[(size_t)u * D + d];
    x[(size_t)r * D + d] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Weight-streaming GEMV for the 2 CFG rows of one utterance, with the RMSNorm of
// fast_model.py:250-261 fused into the prologue and the operator-specific epilogue fused in:
//   EPI_QKV    : Attention.wqkv + KVCache.update scatter      (fast_model.py:206-215, 104-113)
//   EPI_RESID  : wo / w2 projection + residual add             (fast_model.py:226,179-180, 247)
//   EPI_SWIGLU : silu(w1 x) * (w3 x)                           (fast_model.py:230-237)
//   EPI_STORE  : final norm + LM head                          (fast_model.py:161-163)
enum { EPI_STORE = 0, EPI_RESID = 1, EPI_SWIGLU = 2, EPI_QKV = 3 };

struct GemvP {
  const __nv_bfloat16* W;
  const __nv_bfloat16* W3;
  const float* x;
  int ldx;
  const __nv_bfloat16* gain;
  float eps;
  float* out;
  int ldo;
  int M, K;
  // EPI_QKV only
  void* kcache;
  void* vcache;
  int H, S_max, D, kv_fp32;
};

// Position of activation k inside the conflict-free shared-memory layout: a lane's eight
// consecutive k values are split into two float4 that sit 128 floats apart.
__device__ __forceinline__ int xs_perm(int k) {
  const int w = k & 255;
  return (k & ~255) + (((w & 7) >> 2) << 7) + ((w >> 3) << 2) + (w & 3);
}

template <int EPI>
__device__ __forceinline__ void gemv_rows(const GemvP& p, int item, int lane, const uint4*& wa, const uint4*& wb) {
  if (EPI == EPI_SWIGLU) {
    wa = reinterpret_cast<const uint4*>(p.W + (size_t)item * p.K) + lane;
    wb = reinterpret_cast<const uint4*>(p.W3 + (size_t)item * p.K) + lane;
  } else {
    wa = reinterpret_cast<const uint4*>(p.W + (size_t)(2 * item) * p.K) + lane;
    wb = wa + (p.K >> 3);
  }
}

template <int EPI>
__global__ void __launch_bounds__(256, 2) k_gemv(GemvP p, S1State st) {
  extern __shared__ float xs[];  // [2][K] permuted
  __shared__ float red[16];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.K;
  const int n_items = (EPI == EPI_SWIGLU) ? p.M : (p.M >> 1);
  const int nwarps = gridDim.x * 8;
  const int KIT = K >> 8;
  pdl_launch_dependents();
  pdl_wait();  // activations written by the previous kernel are now visible

  const int u = st.slot_map[blockIdx.y];
  const int r0 = 2 * u;
  // ---- prologue: stage (and normalise) the two activation rows
  {
    const float* x0 = p.x + (size_t)r0 * p.ldx;
    const float* x1 = x0 + p.ldx;
    float s0 = 0.f, s1 = 0.f;
    for (int k = tid; k < K; k += 256) {
      const float a = x0[k], b = x1[k];
      s0 = fmaf(a, a, s0);
      s1 = fmaf(b, b, s1);
      const int pk = xs_perm(k);
      xs[pk] = a;
      xs[K + pk] = b;
    }
    if (p.gain != nullptr) {
      s0 = warp_sum(s0);
      s1 = warp_sum(s1);
      if (lane == 0) {
        red[warp] = s0;
        red[8 + warp] = s1;
      }
      __syncthreads();
      float t0 = 0.f, t1 = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        t0 += red[i];
        t1 += red[8 + i];
      }
      const float rs0 = rsqrtf(t0 / (float)K + p.eps);
      const float rs1 = rsqrtf(t1 / (float)K + p.eps);
      for (int k = tid; k < K; k += 256) {
        const float g = bf16_to_f32(p.gain[k]);
        const int pk = xs_perm(k);
        xs[pk] = (xs[pk] * rs0) * g;
        xs[K + pk] = (xs[K + pk] * rs1) * g;
      }
    }
    __syncthreads();
  }

  // ---- main: each warp streams two weight rows at a time
  for (int item = blockIdx.x * 8 + warp; item < n_items; item += nwarps) {
    const uint4 *wa, *wb;
    gemv_rows<EPI>(p, item, lane, wa, wb);
    float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;  // [weight row][activation row]
#pragma unroll 4
    for (int it = 0; it < KIT; ++it) {
      const uint4 va = ldg_stream(wa + it * 32);
      const uint4 vb = ldg_stream(wb + it * 32);
      const float4* xp = reinterpret_cast<const float4*>(xs + it * 256);
      const float4* xq = reinterpret_cast<const float4*>(xs + K + it * 256);
      const float4 x0a = xp[lane], x0b = xp[32 + lane];
      const float4 x1a = xq[lane], x1b = xq[32 + lane];
      fma8(a00, va, x0a, x0b);
      fma8(a01, va, x1a, x1b);
      fma8(a10, vb, x0a, x0b);
      fma8(a11, vb, x1a, x1b);
    }
    a00 = warp_sum(a00);
    a01 = warp_sum(a01);
    a10 = warp_sum(a10);
    a11 = warp_sum(a11);
    if (lane == 0) {
      float* o0 = p.out + (size_t)r0 * p.ldo;
      float* o1 = o0 + p.ldo;
      if (EPI == EPI_SWIGLU) {
        o0[item] = (a00 / (1.f + expf(-a00))) * a10;
        o1[item] = (a01 / (1.f + expf(-a01))) * a11;
      } else {
        const int j = 2 * item;
        if (EPI == EPI_RESID) {
          o0[j] += a00;
          o1[j] += a01;
          o0[j + 1] += a10;
          o1[j + 1] += a11;
        } else {
          o0[j] = a00;
          o1[j] = a01;
          o0[j + 1] = a10;
          o1[j + 1] = a11;
        }
        if (EPI == EPI_QKV && j >= p.D) {
          // scatter k / v of the current token into the cache at input_pos (fast_model.py:111-112)
          const int which = (j - p.D) / p.D;  // 0 = k, 1 = v
          const int jj = (j - p.D) - which * p.D;
          const int head = jj >> 7, d = jj & 127;
          const int pos = st.pos[u];
          const size_t e0 = (((size_t)r0 * p.H + head) * p.S_max + pos) * 128 + d;
          const size_t e1 = e0 + (size_t)p.H * p.S_max * 128;
          void* base = which ? p.vcache : p.kcache;
          if (p.kv_fp32) {
            float* c = reinterpret_cast<float*>(base);
            c[e0] = a00;
            c[e0 + 1] = a10;
            c[e1] = a01;
            c[e1 + 1] = a11;
          } else {
            __nv_bfloat16* c = reinterpret_cast<__nv_bfloat16*>(base);
            *reinterpret_cast<__nv_bfloat162*>(c + e0) = __floats2bfloat162_rn(a00, a10);
            *reinterpret_cast<__nv_bfloat162*>(c + e1) = __floats2bfloat162_rn(a01, a11);
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Decode attention: softmax(q K^T / sqrt(128) + causal mask) V over cache slots [0, pos]
// (fast_model.py:220-224 with the mask row of :148,151).  The reference attends to all 2048 slots
// through a boolean mask; masked slots carry exactly zero weight, so reading only [0, pos] is
// the same arithmetic on 393,216*L bytes instead of 805 MB.
// Split-KV flash decoding: grid (head, row, split); the last-arriving split merges.
constexpr int ATT_SPLITS = 8;

template <bool KV_FP32>
__device__ __forceinline__ void load8(const void* base, size_t elem, float (&v)[8]) {
  if (KV_FP32) {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem);
    const float4 a = __ldcg(p), b = __ldcg(p + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 w = __ldcg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + elem));
    v[0] = bf_lo(w.x); v[1] = bf_hi(w.x); v[2] = bf_lo(w.y); v[3] = bf_hi(w.y);
    v[4] = bf_lo(w.z); v[5] = bf_hi(w.z); v[6] = bf_lo(w.w); v[7] = bf_hi(w.w);
  }
}

// One (query row, head, KV split): q = query vector (fp32, 128), cache_row/L select the K/V slice,
// pidx indexes the per-(row, head) partial slots, `ticket` is that pair's arrival counter.
template <bool KV_FP32>
__device__ __forceinline__ void attn_core(const float* __restrict__ qvec, const void* kcache, const void* vcache,
                                          int cache_row, int L, int h, int split, int H, int S_max,
                                          float* __restrict__ part_o, float* __restrict__ part_ml, size_t pair,
                                          unsigned* ticket, float* __restrict__ out_vec) {
  __shared__ float sm_o[8][128];
  __shared__ float sm_m[8], sm_l[8];
  __shared__ int sm_last;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, half = lane >> 4, sub = lane & 15;
  const int chunk = (((L + ATT_SPLITS - 1) / ATT_SPLITS) + 7) & ~7;
  const int start = split * chunk;
  const int end = min(L, start + chunk);

  float q[8];
  {
    const float scale = 0.08838834764831845f;  // 1/sqrt(128)
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = qvec[sub * 8 + i] * scale;
  }
  const size_t row_base = ((size_t)cache_row * H + h) * S_max * 128 + sub * 8;
  float m = -INFINITY, l = 0.f, o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = 0.f;

  for (int pb = start + warp * 2; pb < end; pb += 8) {
    const int p = pb + half;
    const bool valid = p < end;
    float kv[8];
    float s = 0.f;
    if (valid) {
      load8<KV_FP32>(kcache, row_base + (size_t)p * 128, kv);
#pragma unroll
      for (int i = 0; i < 8; ++i) s = fmaf(q[i], kv[i], s);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 8);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if (valid) {
      load8<KV_FP32>(vcache, row_base + (size_t)p * 128, kv);
      const float mn = fmaxf(m, s);
      const float corr = __expf(m - mn);  // m = -inf on the first visit -> 0
      const float pw = __expf(s - mn);
      l = l * corr + pw;
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = o[i] * corr + pw * kv[i];
      m = mn;
    }
  }
  const int g = warp * 2 + half;
  if (sub == 0) {
    sm_m[g] = m;
    sm_l[g] = l;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sm_o[g][sub * 8 + i] = o[i];
  __syncthreads();

  const size_t pidx = pair * ATT_SPLITS + split;
  {
    float M = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) M = fmaxf(M, sm_m[i]);
    float Ls = 0.f, O = 0.f;
    if (M > -INFINITY) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float w = __expf(sm_m[i] - M);
        Ls += sm_l[i] * w;
        O += sm_o[i][tid] * w;
      }
    }
    part_o[pidx * 128 + tid] = O;
    if (tid == 0) {
      part_ml[pidx * 2] = M;
      part_ml[pidx * 2 + 1] = Ls;
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned t = atomicAdd(ticket, 1u);
    sm_last = (t == ATT_SPLITS - 1);
  }
  __syncthreads();
  if (sm_last) {
    __threadfence();
    const size_t b = pair * ATT_SPLITS;
    float M = -INFINITY;
#pragma unroll
    for (int i = 0; i < ATT_SPLITS; ++i) M = fmaxf(M, __ldcg(part_ml + (b + i) * 2));
    float Ls = 0.f, O = 0.f;
#pragma unroll
    for (int i = 0; i < ATT_SPLITS; ++i) {
      const float mi = __ldcg(part_ml + (b + i) * 2);
      if (mi > -INFINITY) {
        const float w = __expf(mi - M);
        Ls += __ldcg(part_ml + (b + i) * 2 + 1) * w;
        O += __ldcg(part_o + (b + i) * 128 + tid) * w;
      }
    }
    out_vec[tid] = O / Ls;
    if (tid == 0) *ticket = 0u;
  }
}

// Path A: the two CFG rows of each utterance, positions from the decode state.
template <bool KV_FP32>
__global__ void __launch_bounds__(128) k_attn_decode(S1State st, const float* __restrict__ qkv, const void* kcache,
                                                     const void* vcache, float* __restrict__ part_o,
                                                     float* __restrict__ part_ml, float* __restrict__ out, int H,
                                                     int S_max, int D) {
  pdl_launch_dependents();
  pdl_wait();
  const int h = blockIdx.x, split = blockIdx.z;
  const int u = st.slot_map[blockIdx.y >> 1];
  const int r = 2 * u + (blockIdx.y & 1);
  attn_core<KV_FP32>(qkv + (size_t)r * 3 * D + h * 128, kcache, vcache, r, st.pos[u] + 1, h, split, H, S_max, part_o,
                     part_ml, (size_t)r * H + h, &st.attn_ticket[r * H + h], out + (size_t)r * D + h * 128);
}

// Path B: arbitrary activation rows (batched decode rows or the 2*S rows of a prefill chunk); each row
// names its cache row and its own position, i.e. causal attention over [0, pos[row]].
struct RowsDev {
  int* cache_row;
  int* pos;
  int* tok;
  int* utt;
  int* cond;
};

template <bool KV_FP32>
__global__ void __launch_bounds__(128) k_attn_rows(RowsDev rw, const float* __restrict__ qkv, const void* kcache,
                                                   const void* vcache, float* __restrict__ part_o,
                                                   float* __restrict__ part_ml, unsigned* tickets,
                                                   float* __restrict__ out, int H, int S_max, int D) {
  const int h = blockIdx.x, n = blockIdx.y, split = blockIdx.z;
  attn_core<KV_FP32>(qkv + (size_t)n * 3 * D + h * 128, kcache, vcache, rw.cache_row[n], rw.pos[n] + 1, h, split, H, S_max,
                     part_o, part_ml, (size_t)n * H + h, &tickets[n * H + h], out + (size_t)n * D + h * 128);
}

__global__ void k_rows_decode(S1State st, RowsDev rw, int n_utts) {
  const int r = threadIdx.x;
  if (r < 2 * n_utts) {
    const int u = st.slot_map[r >> 1], c = r & 1;
    rw.cache_row[r] = 2 * u + c;
    rw.pos[r] = st.pos[u];
    rw.tok[r] = st.row_tok[2 * u + c];
    rw.utt[r] = u;
    rw.cond[r] = (c == 0);
  }
}

// rows of one prefill chunk: n = c * Sc + s  <->  (CFG row c, token s0 + s of idx[2, S])
__global__ void k_rows_prefill(RowsDev rw, int utt, const int* __restrict__ idx, int S, int s0, int Sc, int pos0) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < 2 * Sc) {
    const int c = n / Sc, s = n - c * Sc;
    rw.cache_row[n] = 2 * utt + c;
    rw.pos[n] = pos0 + s0 + s;
    rw.tok[n] = idx[c * S + s0 + s];
    rw.utt[n] = utt;
    rw.cond[n] = (c == 0);
  }
}

__global__ void __launch_bounds__(256) k_embed_rows(RowsDev rw, const __nv_bfloat16* __restrict__ tok_emb,
                                                    const __nv_bfloat16* __restrict__ pos_emb,
                                                    const float* __restrict__ spk_proj, float* __restrict__ x, int D) {
  const int n = blockIdx.x;
  const __nv_bfloat16* te = tok_emb + (size_t)rw.tok[n] * D;
  const __nv_bfloat16* pe = pos_emb + (size_t)rw.pos[n] * D;
  const bool cond = rw.cond[n] != 0;
  const float* sp = spk_proj + (size_t)rw.utt[n] * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float v = bf16_to_f32(te[d]) + bf16_to_f32(pe[d]);
    if (cond) v += sp[d];
    x[(size_t)n * D + d] = v;
  }
}

// gather the two last-position rows of a prefill chunk (n = Sc-1 and 2*Sc-1) into rows {0, 1}
__global__ void k_gather_last(const float* __restrict__ x, float* __restrict__ dst, int Sc, int D) {
  const int c = blockIdx.x;
  const float* src = x + (size_t)(c * Sc + Sc - 1) * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) dst[(size_t)c * D + d] = src[d];
}

// ---------------------------------------------------------------------------------------------
// sample(): CFG mix, temperature, top-k, top-p (ascending sort, drop cum <= 1-p, keep the last),
// softmax, argmax(p / Exp(1))     (fast_inference_utils.py:61-120).  One CTA per utterance.
// `decode_mode` additionally performs the loop bookkeeping of decode_n_tokens / generate
// (utils:160-172, 212-226): append, feed back, bump position, latch end-of-audio.
constexpr int SAMP_THREADS = 1024;
constexpr int SAMP_PAD = 4096;

struct SampleP {
  const float* logits;  // [rows, V] (decode mode: engine logits buffer) or [2, V] (API mode)
  int V;
  int decode_mode;
  // API mode
  SamplingDev sp;
  const float* noise;
  unsigned long long step;
  int* token_out;
  float* probs_out;
  unsigned* grid_bar;  // decode mode: arrival counter of the persistent decode kernel, reset here between steps
};

__global__ void __launch_bounds__(SAMP_THREADS) k_sample(SampleP p, S1State st) {
  __shared__ float key[SAMP_PAD];
  __shared__ unsigned short sid[SAMP_PAD];
  __shared__ float wsum[32];
  __shared__ float s_bcast[4];
  __shared__ unsigned long long s_best[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int V = p.V;
  pdl_launch_dependents();   // the next decode kernel may start prefetching weights while we sample
  pdl_wait();                // logits of this step are complete
  if (p.decode_mode && p.grid_bar && blockIdx.x == 0 && tid == 0) *p.grid_bar = 0u;
  int u = 0;
  SamplingDev sp = p.sp;
  const float* noise = p.noise;
  unsigned long long step = p.step;
  const float* lc = p.logits;
  if (p.decode_mode) {
    u = st.slot_map[blockIdx.x];
    lc = p.logits + (size_t)(2 * u) * V;
    if (st.done[u]) {
      // a finished utterance still rides along in the batch (its rows are recomputed at a frozen position): hand its
      // logits rows back zeroed so the persistent kernel's red.add accumulation starts from zero every step
      float* z = const_cast<float*>(lc);
      for (int v = tid; v < 2 * V; v += SAMP_THREADS) z[v] = 0.f;
      return;
    }
    sp = st.samp[u];
    step = (unsigned long long)st.n_gen[u];
    noise = st.noise[u] ? st.noise[u] + (size_t)(st.n_gen[u] - st.noise_base[u]) * V : nullptr;
  }
  const float* lu = lc + V;

  // -- CFG mix and temperature, with torch's rounding order (utils:116, :92)
  const float g = sp.guidance, omg = __fsub_rn(1.0f, sp.guidance);
  const float tdiv = fmaxf(sp.temperature, 1e-5f);
  for (int v = tid; v < SAMP_PAD; v += SAMP_THREADS) {
    float k = INFINITY;
    if (v < V) k = __fdiv_rn(__fadd_rn(__fmul_rn(g, lc[v]), __fmul_rn(omg, lu[v])), tdiv);
    key[v] = k;
    sid[v] = (unsigned short)v;
  }
  __syncthreads();
  if (p.decode_mode) {
    // the persistent decode kernel accumulates split-K partial logits with red.add: hand the rows back zeroed
    // (after the barrier: every thread has consumed its logits by now)
    float* z = const_cast<float*>(lc);
    for (int v = tid; v < 2 * V; v += SAMP_THREADS) z[v] = 0.f;
  }

  // -- ascending bitonic sort of (logit, index); padding (+inf) sinks to the end
  for (int k = 2; k <= SAMP_PAD; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < SAMP_PAD / 2; t += SAMP_THREADS) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int ixj = i | j;
        const bool up = ((i & k) == 0);
        const float a = key[i], b = key[ixj];
        const unsigned short ia = sid[i], ib = sid[ixj];
        const bool gt = (a > b) || (a == b && ia > ib);
        if (gt == up) {
          key[i] = b; key[ixj] = a;
          sid[i] = ib; sid[ixj] = ia;
        }
      }
      // strides <= 32: a warp's 32 pair slots (and the second batch at +SAMP_THREADS) stay inside its own two
      // 64-element blocks for all of j = 32 .. 1, so consecutive warp-local stages only need a warp barrier
      // (27 block barriers instead of 78 for 4096 keys)
      const int nj = (j > 1) ? (j >> 1) : k;   // stride of the next stage
      if (j > 32 || nj > 32) __syncthreads();
      else __syncwarp();
    }
  }

  // -- top-k: logits < (k-th largest) -> -inf (utils:94-97)
  if (sp.top_k > 0) {
    const int kk = min(sp.top_k, V);
    const float pivot = key[V - kk];
    __syncthreads();
    for (int i = tid; i < V; i += SAMP_THREADS)
      if (key[i] < pivot) key[i] = -INFINITY;
    __syncthreads();
  }

  // -- softmax over the sorted logits, inclusive cumulative sum (utils:72)
  const float mx = key[V - 1];
  float e[4];
  float loc = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid * 4 + i;
    e[i] = (idx < V) ? expf(key[idx] - mx) : 0.f;
    loc += e[i];
  }
  // block inclusive scan of per-thread totals
  float incl = loc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    float w = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float n = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += n;
    }
    wsum[lane] = w;
  }
  __syncthreads();
  const float total = wsum[31];
  const float before = (warp ? wsum[warp - 1] : 0.f) + (incl - loc);

  // -- top-p: drop sorted entries whose cumulative probability <= 1 - top_p, never the last (utils:75-77)
  float kept_sum_loc = 0.f;
  {
    const bool use_p = sp.top_p > 0.f;
    const float thr = __fsub_rn(1.0f, sp.top_p);
    float run = before;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid * 4 + i;
      run += e[i];
      if (idx < V) {
        const float cum = run / total;
        bool drop = use_p && (cum <= thr) && (idx != V - 1);
        if (key[idx] == -INFINITY) drop = true;
        if (drop) e[i] = 0.f;
        kept_sum_loc += e[i];
      }
    }
  }
  __syncthreads();  // everyone has read key[]/wsum[] -> reuse key[] for the un-sorted probabilities
  float ks = warp_sum(kept_sum_loc);
  if (lane == 0) wsum[warp] = ks;
  // scatter kept exp-values back to vocabulary order (utils:80-81)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid * 4 + i;
    if (idx < V) key[SAMP_PAD - 1 - sid[idx]] = e[i];  // reversed addressing: never collides with unread sorted slots
  }
  __syncthreads();
  if (warp == 0) {
    float w = warp_sum(wsum[lane]);
    if (lane == 0) s_bcast[0] = w;
  }
  __syncthreads();
  const float kept_total = s_bcast[0];

  // -- probs = softmax(masked logits); token = argmax(probs / q)   (utils:101, :61-65)
  unsigned long long best = 0ull;
  for (int v = tid; v < V; v += SAMP_THREADS) {
    const float pr = key[SAMP_PAD - 1 - v] / kept_total;
    if (p.probs_out && !p.decode_mode) p.probs_out[v] = pr;
    float qv;
    if (noise) {
      qv = noise[v];
    } else {
      const uint4 rnd = philox4x32_10(make_uint4((unsigned)v, (unsigned)step, (unsigned)(step >> 32), (unsigned)u),
                                      make_uint2((unsigned)sp.seed, (unsigned)(sp.seed >> 32)));
      const float uni = ((float)(rnd.x >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
      qv = -logf(uni);
    }
    const float score = __fdiv_rn(pr, qv);
    // order-preserving key for non-negative floats; ties resolve to the lowest index like torch.argmax
    const unsigned long long cand =
        ((unsigned long long)__float_as_uint(score) << 32) | (unsigned long long)(0xffffffffu - (unsigned)v);
    best = cand > best ? cand : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long n = __shfl_xor_sync(0xffffffffu, best, o);
    best = n > best ? n : best;
  }
  if (lane == 0) s_best[warp] = best;
  __syncthreads();
  if (warp == 0) {
    best = s_best[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long n = __shfl_xor_sync(0xffffffffu, best, o);
      best = n > best ? n : best;
    }
    if (lane == 0) {
      const int tok = (int)(0xffffffffu - (unsigned)(best & 0xffffffffull));
      if (!p.decode_mode) {
        *p.token_out = tok;
      } else {
        const int n = st.n_gen[u];
        int fed = tok;
        if (st.forced[u]) fed = st.forced[u][n];
        st.sampled_tokens[(size_t)u * st.max_new + n] = tok;
        st.gen_tokens[(size_t)u * st.max_new + n] = fed;
        st.row_tok[2 * u] = fed;
        st.row_tok[2 * u + 1] = fed;
        // Latch termination (utils:161 EOA; utils:196-204 token budget; context end) and keep the position inside the
        // cache: a finished utterance is re-run at its last valid slot while the rest of the batch continues, so no
        // body ever reads pos_emb[block_size] or appends K/V past slot block_size - 1.
        const int np = st.pos[u] + 1;
        const bool stop = fed == sp.end_of_audio || np >= st.block_size || n + 1 >= st.max_new || n + 1 >= st.budget[u];
        if (!stop) st.pos[u] = np;
        st.n_gen[u] = n + 1;
        if (stop) st.done[u] = 1;
      }
    }
  }
}

}  // namespace mvb
