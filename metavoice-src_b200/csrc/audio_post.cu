// On-device audio post-processing: what `_save_audio` does on the host in the reference
// (fam/llm/decoders.py:40-47 -> audiocraft audio_write(strategy="loudness", loudness_compressor=True)):
//   integrated loudness (ITU-R BS.1770-4 as implemented by torchaudio.functional.loudness, the function audiocraft
//   calls: K-weighting = treble shelf + 38 Hz high-pass biquads, 400 ms blocks with 75 % overlap, absolute -70 LKFS and
//   relative -10 LU gates) -> gain to -14 LUFS -> tanh compressor -> clip -> PCM16.
// The waveform never leaves the device as fp32; the caller gets 16-bit samples (and the measured loudness).
//
// The two K-weighting biquads are recursive; they are evaluated in independent chunks, each warmed up over the
// preceding 4096 samples from a zero state (the slowest pole, the 38 Hz high-pass at 24 kHz, has decayed by e^-40 by
// then), so the filter runs in ~6 K sequential steps instead of T.
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/mvb200.h"
#include "common.cuh"

using namespace mvb;
namespace mvb { int set_error(int code, const char* fmt, ...); }

#define ACK(expr)                                                                                    \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess)                                                                           \
      return mvb::set_error(MVB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

namespace {

struct Biquad { float b0, b1, b2, a1, a2; };   // normalised by a0

constexpr int KW_CHUNK = 2048, KW_WARM = 4096;

// y = clamp(hp(clamp(shelf(x))))   (torchaudio lfilter clamps each biquad's OUTPUT to [-1, 1], not its state)
__global__ void __launch_bounds__(128) k_kweight(const float* __restrict__ x, int T, Biquad s, Biquad h, float* __restrict__ y) {
  const int chunk = blockIdx.x * blockDim.x + threadIdx.x;
  const int t0 = chunk * KW_CHUNK;
  if (t0 >= T) return;
  const int start = max(0, t0 - KW_WARM), end = min(T, t0 + KW_CHUNK);
  float x1 = 0.f, x2 = 0.f, u1 = 0.f, u2 = 0.f;      // shelf: input / output history
  float c1 = 0.f, c2 = 0.f, v1 = 0.f, v2 = 0.f;      // high-pass: (clamped) input / output history
  for (int t = start; t < end; ++t) {
    const float xv = x[t];
    const float u = s.b0 * xv + s.b1 * x1 + s.b2 * x2 - s.a1 * u1 - s.a2 * u2;
    x2 = x1; x1 = xv; u2 = u1; u1 = u;
    const float c = fminf(fmaxf(u, -1.f), 1.f);
    const float v = h.b0 * c + h.b1 * c1 + h.b2 * c2 - h.a1 * v1 - h.a2 * v2;
    c2 = c1; c1 = c; v2 = v1; v1 = v;
    if (t >= t0) y[t] = fminf(fmaxf(v, -1.f), 1.f);
  }
}

// block b: mean of y^2 over [b*step, b*step + gate); block n_blocks: sum of x^2 over the whole signal (energy floor test)
__global__ void __launch_bounds__(256) k_block_energy(const float* __restrict__ y, const float* __restrict__ x, int T, int gate, int step,
                                                      int n_blocks, float* __restrict__ energy) {
  __shared__ float red[8];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* src = b < n_blocks ? y + (size_t)b * step : x;
  const int n = b < n_blocks ? gate : T;
  float acc = 0.f;
  for (int i = tid; i < n; i += 256) acc = fmaf(src[i], src[i], acc);
  acc = warp_sum(acc);
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    energy[b] = b < n_blocks ? t / (float)gate : t;
  }
}

// gating + gain (one thread: <= a few hundred blocks).  out[0] = LKFS, out[1] = linear gain (1 when the signal is
// below the energy floor, as audiocraft's normalize_loudness leaves it untouched; the compressor is skipped too: out[2] = 0)
__global__ void k_loudness_gain(const float* __restrict__ energy, int n_blocks, int T, float headroom_db, float energy_floor,
                                float* __restrict__ out) {
  if (threadIdx.x != 0) return;
  const float rms = sqrtf(energy[n_blocks] / (float)T);
  if (rms < energy_floor || n_blocks < 1) { out[0] = -INFINITY; out[1] = 1.f; out[2] = 0.f; return; }
  const float bias = -0.691f, gamma_abs = -70.f;
  float sum = 0.f; int cnt = 0;
  for (int b = 0; b < n_blocks; ++b) {
    const float l = bias + 10.f * log10f(energy[b]);
    if (l > gamma_abs) { sum += energy[b]; ++cnt; }
  }
  const float gamma_rel = bias + 10.f * log10f(sum / (float)cnt) - 10.f;
  sum = 0.f; cnt = 0;
  for (int b = 0; b < n_blocks; ++b) {
    const float l = bias + 10.f * log10f(energy[b]);
    if (l > gamma_abs && l > gamma_rel) { sum += energy[b]; ++cnt; }
  }
  const float lkfs = bias + 10.f * log10f(sum / (float)cnt);
  out[0] = lkfs;
  out[1] = powf(10.f, (-headroom_db - lkfs) / 20.f);
  out[2] = 1.f;
}

// out = clip(tanh(gain * x)) -> PCM16 (round to nearest even, like numpy's round in audio_out.py / torchaudio.save)
__global__ void __launch_bounds__(256) k_apply_pcm16(const float* __restrict__ x, int T, const float* __restrict__ g, int compressor,
                                                     short* __restrict__ pcm, float* __restrict__ wav_out) {
  const float gain = g[1];
  const bool comp = compressor && g[2] != 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T; i += gridDim.x * blockDim.x) {
    float v = gain * x[i];
    if (comp) v = tanhf(v);
    v = fminf(fmaxf(v, -1.f), 1.f);
    if (wav_out) wav_out[i] = v;
    pcm[i] = (short)__float2int_rn(v * 32767.f);
  }
}

static Biquad treble_shelf(double sr, double gain_db, double f0, double Q) {   // torchaudio.functional.treble_biquad
  const double w0 = 2 * M_PI * f0 / sr, alpha = sin(w0) / 2 / Q, A = exp(gain_db / 40 * log(10.0));
  const double t1 = 2 * sqrt(A) * alpha, t2 = (A - 1) * cos(w0), t3 = (A + 1) * cos(w0);
  const double b0 = A * ((A + 1) + t2 + t1), b1 = -2 * A * ((A - 1) + t3), b2 = A * ((A + 1) + t2 - t1);
  const double a0 = (A + 1) - t2 + t1, a1 = 2 * ((A - 1) - t3), a2 = (A + 1) - t2 - t1;
  return Biquad{(float)(b0 / a0), (float)(b1 / a0), (float)(b2 / a0), (float)(a1 / a0), (float)(a2 / a0)};
}
static Biquad highpass(double sr, double f0, double Q) {                      // torchaudio.functional.highpass_biquad
  const double w0 = 2 * M_PI * f0 / sr, alpha = sin(w0) / 2 / Q;
  const double b0 = (1 + cos(w0)) / 2, b1 = -1 - cos(w0), b2 = b0, a0 = 1 + alpha, a1 = -2 * cos(w0), a2 = 1 - alpha;
  return Biquad{(float)(b0 / a0), (float)(b1 / a0), (float)(b2 / a0), (float)(a1 / a0), (float)(a2 / a0)};
}

}  // namespace

extern "C" size_t mvb_audio_post_workspace_bytes(int32_t max_samples) {
  if (max_samples < 1) return 0;
  return ((size_t)max_samples * 4 + 255) / 256 * 256 + 65536;     // K-weighted copy | block energies + result scalars
}

extern "C" int mvb_audio_post(const float* d_wav, int32_t n_samples, int32_t sample_rate, float loudness_headroom_db,
                              int32_t loudness_compressor, void* d_workspace, int16_t* d_pcm16, float* d_wav_out,
                              float* d_lkfs_gain, void* stream) {
  if (!d_wav || !d_workspace || !d_pcm16) return mvb::set_error(MVB_ERR_ARG, "null argument");
  if (n_samples < 1 || sample_rate < 8000) return mvb::set_error(MVB_ERR_ARG, "audio_post: bad length / sample rate");
  cudaStream_t s = (cudaStream_t)stream;
  float* y = reinterpret_cast<float*>(d_workspace);
  float* energy = reinterpret_cast<float*>(reinterpret_cast<char*>(d_workspace) + ((size_t)n_samples * 4 + 255) / 256 * 256);
  const int gate = (int)lround(0.4 * sample_rate), step = (int)lround(gate * 0.25);
  const int n_blocks = n_samples >= gate ? (n_samples - gate) / step + 1 : 0;
  if (n_blocks + 4 > 16000) return mvb::set_error(MVB_ERR_UNSUPPORTED, "audio_post: signal too long for the workspace");
  float* res = energy + n_blocks + 1;
  const int chunks = (n_samples + KW_CHUNK - 1) / KW_CHUNK;
  k_kweight<<<(chunks + 127) / 128, 128, 0, s>>>(d_wav, n_samples, treble_shelf(sample_rate, 4.0, 1500.0, 1.0 / sqrt(2.0)),
                                                 highpass(sample_rate, 38.0, 0.5), y);
  ACK(cudaGetLastError());
  k_block_energy<<<n_blocks + 1, 256, 0, s>>>(y, d_wav, n_samples, gate, step, n_blocks, energy);
  ACK(cudaGetLastError());
  k_loudness_gain<<<1, 32, 0, s>>>(energy, n_blocks, n_samples, loudness_headroom_db, 2e-3f, res);
  ACK(cudaGetLastError());
  k_apply_pcm16<<<148 * 4, 256, 0, s>>>(d_wav, n_samples, res, loudness_compressor, d_pcm16, d_wav_out);
  ACK(cudaGetLastError());
  if (d_lkfs_gain) ACK(cudaMemcpyAsync(d_lkfs_gain, res, 2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return MVB_OK;
}
