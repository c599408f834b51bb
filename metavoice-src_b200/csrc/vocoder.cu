// EnCodec-24 kHz decode path of the vocoder (first half): RVQ decode -> SEANet decoder.
// This is what `mbd.tokens_to_wav` (fam/llm/decoders.py:85 -> audiocraft 1.2.0) computes first, as the loudness /
// band-energy reference and as the diffusion condition (`decode_latent`).  audiocraft delegates the 24 kHz codec to
// transformers.EncodecModel; operator semantics follow modeling_encodec.py (cited per kernel).  fp32 throughout
// (the reference forces fp32 autocast for the vocoder, decoders.py:84).
//
// Activations are [C][T] fp32 per utterance (time contiguous): every conv streams its input window through a
// shared-memory line buffer (128-bit coalesced loads of the tile body, scalar loads for the K-1 halo samples), keeps an
// 8-channel x 4-sample register tile per thread and writes it back with 128-bit stores.
#include <cuda_runtime.h>

#include <vector>

#include "../../include/mvb200.h"
#include "common.cuh"

using namespace mvb;
namespace mvb { int set_error(int code, const char* fmt, ...); }

#define VCK(expr)                                                                                    \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      return mvb::set_error(MVB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

namespace {

__device__ __forceinline__ float elu1(float v) { return v > 0.f ? v : expm1f(v); }

// latent[c][t] = sum_q codebook_q[codes[q][t]][c]     (EncodecResidualVectorQuantizer.decode)
__global__ void __launch_bounds__(128) k_rvq_decode(const int* __restrict__ codes, int n_q, int T, const float* const* __restrict__ books,
                                                    int dim, float* __restrict__ out) {
  __shared__ float tile[32][129];
  const int t0 = blockIdx.x * 32, b = blockIdx.y;
  const int c = threadIdx.x;
  for (int i = 0; i < 32; ++i) {
    const int t = t0 + i;
    float acc = 0.f;
    if (t < T && c < dim)
      for (int q = 0; q < n_q; ++q) acc += books[q][(size_t)codes[((size_t)b * n_q + q) * T + t] * dim + c];
    tile[i][c] = acc;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 32 * dim; idx += 128) {
    const int cc = idx / 32, i = idx - cc * 32;
    if (t0 + i < T) out[((size_t)b * dim + cc) * T + t0 + i] = tile[i][cc];
  }
}

// Causal Conv1d, stride 1, dilation 1, reflect padding of K-1 samples on the left (EncodecConv1d, causal branch),
// optional ELU on the input (the nn.ELU() that precedes the conv in the stack) and optional accumulation into y
// (residual / shortcut sum of EncodecResnetBlock).  Tile: 32 output channels x 256 samples, 16 input channels per pass.
template <int K>
__global__ void __launch_bounds__(256) k_conv1d(const float* __restrict__ x, int Cin, int T, const float* __restrict__ w,
                                                const float* __restrict__ bias, float* __restrict__ y, int Cout, int elu_in, int accumulate) {
  constexpr int CI = 16, TT = 256, CO = 32;
  __shared__ float xs[CI][TT + K - 1 + 1];
  __shared__ float ws[CO][CI][K];
  const int t0 = blockIdx.x * TT, co0 = blockIdx.y * CO, b = blockIdx.z;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;    // 64 time groups of 4 samples x 4 groups of 8 channels
  const float* xb = x + (size_t)b * Cin * T;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int c0 = 0; c0 < Cin; c0 += CI) {
    __syncthreads();
    // line buffer: samples t0-(K-1) .. t0+TT-1 of 16 channels.  The TT-sample body is streamed with 128-bit loads
    // (rows are 16-byte aligned when T % 4 == 0, which holds for every layer of the decoder: T = frames x 8 x ...);
    // the K-1 halo samples in front are loaded one by one, negative times reflect (x[-j] = x[j]).
    const bool vec = (T & 3) == 0;
    if (vec) {
      for (int i = threadIdx.x; i < CI * (TT / 4); i += 256) {
        const int ci = i / (TT / 4), q4 = i - ci * (TT / 4);
        const int t = t0 + 4 * q4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c0 + ci < Cin && t < T) {          // T % 4 == 0: a float4 is either fully inside or fully outside
          v = *reinterpret_cast<const float4*>(xb + (size_t)(c0 + ci) * T + t);
          if (elu_in) { v.x = elu1(v.x); v.y = elu1(v.y); v.z = elu1(v.z); v.w = elu1(v.w); }
        }
        float* d = &xs[ci][K - 1 + 4 * q4];
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      }
    }
    for (int i = threadIdx.x; i < CI * (vec ? K - 1 : TT + K - 1); i += 256) {
      const int span = vec ? K - 1 : TT + K - 1;
      const int ci = i / span, tt = i - ci * span;
      int t = t0 + tt - (K - 1);
      if (t < 0) t = -t;
      float v = 0.f;
      if (c0 + ci < Cin && t < T) {
        v = xb[(size_t)(c0 + ci) * T + t];
        if (elu_in) v = elu1(v);
      }
      xs[ci][tt] = v;
    }
    for (int i = threadIdx.x; i < CO * CI * K; i += 256) {
      const int co = i / (CI * K), r = i - co * (CI * K), ci = r / K, k = r - ci * K;
      ws[co][ci][k] = (co0 + co < Cout && c0 + ci < Cin) ? w[((size_t)(co0 + co) * Cin + c0 + ci) * K + k] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int ci = 0; ci < CI; ++ci) {
      float xv[4 + K - 1];
#pragma unroll
      for (int j = 0; j < 4 + K - 1; ++j) xv[j] = xs[ci][tx * 4 + j];
#pragma unroll
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float wv = ws[ty * 8 + i][ci][k];
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wv, xv[j + k], acc[i][j]);
        }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int co = co0 + ty * 8 + i;
    if (co >= Cout) continue;
    const float bv = bias ? bias[co] : 0.f;
    const int tb = t0 + tx * 4;
    if (((T & 3) == 0) && tb < T) {            // 128-bit coalesced store of the thread's 4 consecutive samples
      float4* o = reinterpret_cast<float4*>(y + ((size_t)b * Cout + co) * T + tb);
      float4 v = make_float4(acc[i][0] + bv, acc[i][1] + bv, acc[i][2] + bv, acc[i][3] + bv);
      if (accumulate) { const float4 p = *o; v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w; }
      *o = v;
      continue;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 + tx * 4 + j;
      if (t < T) {
        float* o = y + ((size_t)b * Cout + co) * T + t;
        const float v = acc[i][j] + bv;
        *o = accumulate ? *o + v : v;
      }
    }
  }
}

// ConvTranspose1d(k = 2r, stride r) trimmed by r samples on the right (EncodecConvTranspose1d, causal) with the
// preceding ELU fused:  y[co][t*r + p] = b[co] + sum_ci w[ci][co][p] * e[ci][t] + w[ci][co][p + r] * e[ci][t-1].
__global__ void __launch_bounds__(256) k_convtr1d(const float* __restrict__ x, int Cin, int T, const float* __restrict__ w,
                                                  const float* __restrict__ bias, float* __restrict__ y, int Cout, int r) {
  constexpr int CI = 16, TI = 32, CO = 32;
  extern __shared__ float sm[];
  float* xs = sm;                       // [CI][TI + 1]   (index 0 = t0 - 1)
  float* ws = sm + CI * (TI + 1);       // [CI][CO][2r]
  const int t0 = blockIdx.x * TI, co0 = blockIdx.y * CO, b = blockIdx.z;
  const int ox = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int n_out = TI * r;             // outputs of this tile
  const float* xb = x + (size_t)b * Cin * T;
  float acc[8][4];
  int tl[4], pp[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int o = ox + 64 * j;
    tl[j] = o / r;
    pp[j] = o - tl[j] * r;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int c0 = 0; c0 < Cin; c0 += CI) {
    __syncthreads();
    for (int i = threadIdx.x; i < CI * (TI + 1); i += 256) {
      const int ci = i / (TI + 1), tt = i - ci * (TI + 1);
      const int t = t0 + tt - 1;
      float v = 0.f;
      if (c0 + ci < Cin && t >= 0 && t < T) v = elu1(xb[(size_t)(c0 + ci) * T + t]);
      xs[i] = v;
    }
    for (int i = threadIdx.x; i < CI * CO * 2 * r; i += 256) {
      const int ci = i / (CO * 2 * r), rem = i - ci * (CO * 2 * r), co = rem / (2 * r), k = rem - co * (2 * r);
      ws[i] = (c0 + ci < Cin && co0 + co < Cout) ? w[((size_t)(c0 + ci) * Cout + co0 + co) * (2 * r) + k] : 0.f;
    }
    __syncthreads();
    for (int ci = 0; ci < CI; ++ci) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (ox + 64 * j < n_out) {
          const float e1 = xs[ci * (TI + 1) + tl[j] + 1], e0 = xs[ci * (TI + 1) + tl[j]];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float* wp = ws + ((size_t)ci * CO + ty * 8 + i) * 2 * r;
            acc[i][j] = fmaf(wp[pp[j]], e1, fmaf(wp[pp[j] + r], e0, acc[i][j]));
          }
        }
      }
    }
  }
  const int To = T * r;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int co = co0 + ty * 8 + i;
    if (co >= Cout) continue;
    const float bv = bias[co];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int o = ox + 64 * j;
      const int n = t0 * r + o;
      if (o < n_out && n < To) y[((size_t)b * Cout + co) * To + n] = acc[i][j] + bv;
    }
  }
}

// One LSTM layer (EncodecLSTM: nn.LSTM, gate order i, f, g, o, zero initial state), persistent over time:
// CTA c owns hidden units [4c, 4c+4) = 16 gate rows of W_hh held in shared memory; h_{t-1} is exchanged through
// global memory and one grid-wide arrive/wait per step.  pre[row][t] = W_ih x_t + b_ih + b_hh is computed beforehand
// by k_conv1d<1>.  The skip connection of EncodecLSTM (output + input) is fused into the last layer's store.
__global__ void __launch_bounds__(128) k_lstm_layer(const float* __restrict__ pre, const float* __restrict__ whh, int H, int T,
                                                    float* __restrict__ hbuf /*[2][H]*/, unsigned* bar, float* __restrict__ out,
                                                    const float* __restrict__ skip) {
  extern __shared__ float sm[];
  float* wsm = sm;               // [16][H]
  float* hs = sm + 16 * H;       // [H]
  __shared__ float gates[16];
  const int cta = blockIdx.x, tid = threadIdx.x;
  const int u0 = cta * 4;
  for (int i = tid; i < 16 * H; i += 128) {
    const int r = i / H, k = i - r * H;
    const int grow = (r >> 2) * H + u0 + (r & 3);
    wsm[i] = whh[(size_t)grow * H + k];
  }
  float c_state = 0.f;
  const int row = tid >> 3, part = tid & 7;      // 16 rows x 8 lanes
  const int seg = H / 8;
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    if (t > 0) {
      if (tid == 0) {
        const unsigned target = (unsigned)t * gridDim.x;
        const long long t0 = clock64();
        unsigned v;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
          if (clock64() - t0 > 4000000000ll) __trap();
        } while (v < target);
      }
      __syncthreads();
      const float* hp = hbuf + (size_t)((t - 1) & 1) * H;
      for (int k = tid; k < H; k += 128) hs[k] = __ldcg(hp + k);
    } else {
      for (int k = tid; k < H; k += 128) hs[k] = 0.f;
    }
    __syncthreads();
    float a = 0.f;
    const float* wr = wsm + row * H + part * seg;
    const float* hh = hs + part * seg;
#pragma unroll 8
    for (int k = 0; k < seg; ++k) a = fmaf(wr[k], hh[k], a);
    a += __shfl_xor_sync(0xffffffffu, a, 4);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    if (part == 0) {
      const int grow = (row >> 2) * H + u0 + (row & 3);
      gates[row] = a + pre[(size_t)grow * T + t];
    }
    __syncthreads();
    if (tid < 4) {
      const float ig = 1.f / (1.f + expf(-gates[tid]));
      const float fg = 1.f / (1.f + expf(-gates[4 + tid]));
      const float gg = tanhf(gates[8 + tid]);
      const float og = 1.f / (1.f + expf(-gates[12 + tid]));
      c_state = fg * c_state + ig * gg;
      const float h = og * tanhf(c_state);
      hbuf[(size_t)(t & 1) * H + u0 + tid] = h;
      out[(size_t)(u0 + tid) * T + t] = skip ? h + skip[(size_t)(u0 + tid) * T + t] : h;
    }
    __syncthreads();
    if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
struct mvb_voc {
  mvb_voc_config cfg;
  const char* arena;
  std::vector<uint64_t> off;
  char* ws;
  float *bufA, *bufB, *bufC, *pre;
  float* hbuf;
  unsigned* bar;
  const float** books_dev;
  const float* w(int i) const { return reinterpret_cast<const float*>(arena + off[i]); }
};

// tensor order in the arena (fp32, weight norm folded by the loader):
//   codebooks[n_q] | conv_in {w,b} | lstm l0 {w_ih, w_hh, b_ih+b_hh} | lstm l1 {...} |
//   per ratio: up {w,b}, res.c1 {w,b}, res.c2 {w,b}, res.shortcut {w,b} | conv_out {w,b}
static int voc_n_tensors(const mvb_voc_config* c) { return c->n_q + 2 + 6 + c->n_ratios * 8 + 2; }

static size_t voc_layout(const mvb_voc_config* c, size_t* oA, size_t* oB, size_t* oC, size_t* oPre, size_t* oH, size_t* oBar, size_t* oBooks) {
  size_t T = c->max_frames, maxelems = (size_t)c->hidden * c->n_filters / c->n_filters * T;  // placeholder, refined below
  int ch = c->n_filters << c->n_ratios;   // 512
  maxelems = (size_t)ch * T;
  size_t Tl = T;
  for (int i = 0; i < c->n_ratios; ++i) {
    Tl *= c->ratios[i];
    ch >>= 1;
    if ((size_t)ch * Tl > maxelems) maxelems = (size_t)ch * Tl;
  }
  const size_t top = (size_t)(c->n_filters << c->n_ratios);
  size_t o = 0;
  auto take = [&](size_t b) { size_t r = o; o = (o + b + 255) / 256 * 256; return r; };
  *oA = take(maxelems * 4); *oB = take(maxelems * 4); *oC = take(maxelems * 4);
  *oPre = take(4 * top * T * 4);
  *oH = take(2 * top * 4);
  *oBar = take(256);
  *oBooks = take(64 * sizeof(void*));
  return o;
}

static int voc_validate(const mvb_voc_config* c) {
  if (!c) return mvb::set_error(MVB_ERR_ARG, "null config");
  if (c->n_ratios < 1 || c->n_ratios > 8 || c->n_q < 1 || c->n_q > 32) return mvb::set_error(MVB_ERR_ARG, "bad codec shape");
  if (c->hidden > 128) return mvb::set_error(MVB_ERR_UNSUPPORTED, "codebook dim > 128");
  const int top = c->n_filters << c->n_ratios;
  if (top % 32 || top / 4 > 148) return mvb::set_error(MVB_ERR_UNSUPPORTED, "LSTM width %d unsupported", top);
  for (int i = 0; i < c->n_ratios; ++i)
    if (c->ratios[i] < 1 || c->ratios[i] > 8) return mvb::set_error(MVB_ERR_UNSUPPORTED, "upsampling ratio out of range");
  if (c->kernel != 7 || c->res_kernel != 3 || c->last_kernel != 7) return mvb::set_error(MVB_ERR_UNSUPPORTED, "kernel sizes must be 7/3/7");
  if (c->max_frames < 8) return mvb::set_error(MVB_ERR_ARG, "max_frames too small");
  return MVB_OK;
}

extern "C" size_t mvb_voc_workspace_bytes(const mvb_voc_config* c) {
  if (voc_validate(c)) return 0;
  size_t a, b, cc, d, e, f, g;
  return voc_layout(c, &a, &b, &cc, &d, &e, &f, &g);
}

extern "C" int mvb_voc_create(const mvb_voc_config* cfg, const void* d_arena, size_t arena_bytes, const uint64_t* offsets, void* d_ws,
                              mvb_voc** out) {
  if (int e = voc_validate(cfg)) return e;
  if (!d_arena || !offsets || !d_ws || !out) return mvb::set_error(MVB_ERR_ARG, "null pointer argument");
  mvb_voc* h = new mvb_voc();
  h->cfg = *cfg;
  h->arena = reinterpret_cast<const char*>(d_arena);
  h->off.assign(offsets, offsets + voc_n_tensors(cfg));
  for (uint64_t o : h->off)
    if (o % 16 || o >= arena_bytes) { delete h; return mvb::set_error(MVB_ERR_ARG, "bad weight offset"); }
  h->ws = reinterpret_cast<char*>(d_ws);
  size_t oA, oB, oC, oP, oH, oBar, oBk;
  voc_layout(cfg, &oA, &oB, &oC, &oP, &oH, &oBar, &oBk);
  h->bufA = (float*)(h->ws + oA); h->bufB = (float*)(h->ws + oB); h->bufC = (float*)(h->ws + oC); h->pre = (float*)(h->ws + oP);
  h->hbuf = (float*)(h->ws + oH); h->bar = (unsigned*)(h->ws + oBar); h->books_dev = (const float**)(h->ws + oBk);
  std::vector<const float*> books(cfg->n_q);
  for (int q = 0; q < cfg->n_q; ++q) books[q] = h->w(q);
  VCK(cudaMemcpy(h->books_dev, books.data(), sizeof(void*) * cfg->n_q, cudaMemcpyHostToDevice));
  *out = h;
  return MVB_OK;
}

extern "C" int mvb_voc_destroy(mvb_voc* h) {
  delete h;
  return MVB_OK;
}

template <int K>
static cudaError_t conv(cudaStream_t s, const float* x, int Cin, int T, const float* w, const float* b, float* y, int Cout, int elu, int acc) {
  dim3 grid((T + 255) / 256, (Cout + 31) / 32, 1);
  k_conv1d<K><<<grid, 256, 0, s>>>(x, Cin, T, w, b, y, Cout, elu, acc);
  return cudaGetLastError();
}

// RVQ decode only: the diffusion condition `decode_latent` (and the decoder's input).  d_codes int32 [n_q, T].
extern "C" int mvb_voc_decode_latent(mvb_voc* h, const int32_t* d_codes, int32_t T, float* d_latent, void* stream) {
  if (!h || !d_codes || !d_latent) return mvb::set_error(MVB_ERR_ARG, "null argument");
  if (T < 1 || T > h->cfg.max_frames) return mvb::set_error(MVB_ERR_ARG, "frame count %d out of range", T);
  k_rvq_decode<<<dim3((T + 31) / 32, 1), 128, 0, (cudaStream_t)stream>>>(d_codes, h->cfg.n_q, T, h->books_dev, h->cfg.hidden, d_latent);
  VCK(cudaGetLastError());
  return MVB_OK;
}

// codes int32 [n_q, T] -> waveform fp32 [T * prod(ratios)]   (== EncodecModel.decode for one chunk)
extern "C" int mvb_voc_decode(mvb_voc* h, const int32_t* d_codes, int32_t T, float* d_wav, void* stream) {
  if (!h || !d_codes || !d_wav) return mvb::set_error(MVB_ERR_ARG, "null argument");
  const mvb_voc_config& c = h->cfg;
  if (T < 8 || T > c.max_frames) return mvb::set_error(MVB_ERR_ARG, "frame count %d out of range [8, %d]", T, c.max_frames);
  cudaStream_t s = (cudaStream_t)stream;
  const int top = c.n_filters << c.n_ratios;
  int ti = c.n_q;
  float *A = h->bufA, *B = h->bufB, *Cc = h->bufC;
  if (int e = mvb_voc_decode_latent(h, d_codes, T, A, stream)) return e;
  VCK(conv<7>(s, A, c.hidden, T, h->w(ti), h->w(ti + 1), B, top, 0, 0));      // conv_in
  ti += 2;
  // LSTM: x = B;  layer 0 -> A ; layer 1 (+ skip x) -> Cc
  const size_t lsmem = (size_t)(16 * top + top) * 4;
  static PerDeviceOnce attr;
  if (attr.pending()) { VCK(cudaFuncSetAttribute(k_lstm_layer, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); attr.mark(); }
  const float* lin = B;
  float* louts[2] = {A, Cc};
  for (int l = 0; l < 2; ++l) {
    VCK(conv<1>(s, lin, top, T, h->w(ti), h->w(ti + 2), h->pre, 4 * top, 0, 0));   // W_ih x + (b_ih + b_hh)
    VCK(cudaMemsetAsync(h->bar, 0, 4, s));
    k_lstm_layer<<<top / 4, 128, lsmem, s>>>(h->pre, h->w(ti + 1), top, T, h->hbuf, h->bar, louts[l], l == 1 ? B : nullptr);
    VCK(cudaGetLastError());
    lin = louts[l];
    ti += 3;
  }
  float* x = Cc;           // current activation
  float* f1 = A;
  float* f2 = B;
  int ch = top, Tl = T;
  for (int i = 0; i < c.n_ratios; ++i) {
    const int r = c.ratios[i], co = ch / 2;
    const size_t smem = (size_t)(16 * 33 + 16 * 32 * 2 * r) * 4;
    k_convtr1d<<<dim3((Tl + 31) / 32, (co + 31) / 32, 1), 256, smem, s>>>(x, ch, Tl, h->w(ti), h->w(ti + 1), f1, co, r);
    VCK(cudaGetLastError());
    Tl *= r; ch = co;
    // resnet block: y = shortcut(u) + conv_k1(elu(conv_k3(elu(u))))
    VCK(conv<3>(s, f1, ch, Tl, h->w(ti + 2), h->w(ti + 3), f2, ch / c.compress, 1, 0));
    VCK(conv<1>(s, f1, ch, Tl, h->w(ti + 6), h->w(ti + 7), x, ch, 0, 0));
    VCK(conv<1>(s, f2, ch / c.compress, Tl, h->w(ti + 4), h->w(ti + 5), x, ch, 1, 1));
    ti += 8;
  }
  VCK(conv<7>(s, x, ch, Tl, h->w(ti), h->w(ti + 1), d_wav, 1, 1, 0));
  return MVB_OK;
}
