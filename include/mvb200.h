/*
 * mvb200.h -- C ABI of libmvb200.so: the B200 (sm_100a) engine behind MetaVoice-1B's
 * TTS.synthesise() hot path.  Plain pointers and sizes only; no torch types.
 *
 * Every entry point names the reference interface it replaces (paths relative to the
 * reference tree metavoiceio/metavoice-src @ de3fa211).  Device buffers are owned by the
 * caller (the Python shim allocates them as torch tensors and passes data_ptr()); the
 * library never frees caller memory.  All functions return 0 on success, non-zero on
 * failure with the message available from mvb_last_error().  A handle is re-entrant per
 * handle, not thread-safe per handle (the reference is single-threaded per model,
 * SURVEY.md 8b "Threading / state").
 */
#ifndef MVB200_H
#define MVB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVB_ABI_VERSION 2

/* KV-cache element type.  bf16 is what the reference stores (fast_model.py:97-102);
 * fp32 is the validation mode used for the <=1e-3 parity gate against the fp32 oracle. */
enum { MVB_KV_BF16 = 0, MVB_KV_FP32 = 1 };

/* Shape of the stage-1 causal LM: fam/llm/fast_model.py:52-94 (ModelArgs / "metavoice-1B"). */
typedef struct mvb_s1_config {
  int32_t n_layer;       /* 24 */
  int32_t n_head;        /* 16 (MHA: n_local_heads == n_head) */
  int32_t head_dim;      /* 128 (only 128 is supported) */
  int32_t dim;           /* 2048 */
  int32_t intermediate;  /* 5632 */
  int32_t vocab;         /* 2562 */
  int32_t block_size;    /* 2048: learned-position table length and KV slots per row */
  int32_t spk_dim;       /* 256 */
  float   norm_eps;      /* 1e-5 */
  int32_t max_utts;      /* utterance slots; each owns 2 CFG rows {cond, uncond} (fast_model.py:132-134) */
  int32_t kv_dtype;      /* MVB_KV_BF16 | MVB_KV_FP32 */
  int32_t max_new;       /* capacity of the per-utterance generated-token ring (<= block_size) */
} mvb_s1_config;

/* Byte offsets (into one bf16 weight arena) of every stage-1 tensor, row-major [out, in]
 * exactly as stored in first_stage.pt.  Built by the checkpoint loader that
 * replaces fast_inference_utils.py:236-281 (_load_model).  Per layer, in this order:
 * attn_norm[dim], wqkv[3*dim, dim], wo[dim, dim], ffn_norm[dim], w1[inter, dim], w3[inter, dim],
 * w2[dim, inter]. */
#define MVB_S1_GLOBAL_TENSORS 5 /* tok_emb[vocab,dim], pos_emb[block,dim], spk_proj[dim,spk_dim], out_norm[dim], lm_head[vocab,dim] */
#define MVB_S1_LAYER_TENSORS 7

/* Sampling parameters of fast_inference_utils.py:107-120 (sample) for one utterance. */
typedef struct mvb_sampling {
  float    guidance_scale; /* g in g*cond + (1-g)*uncond */
  float    temperature;    /* clamped below at 1e-5 (utils:92) */
  float    top_p;          /* <= 0 disables (top_p=None) */
  int32_t  top_k;          /* <= 0 disables (top_k=None) */
  int32_t  end_of_audio;   /* token id that latches termination (2048; 9999 disables), utils:161 */
  uint64_t seed;           /* Philox key for on-device Exp(1) noise when no noise buffer is supplied */
} mvb_sampling;

typedef struct mvb_s1 mvb_s1; /* opaque engine handle */

int         mvb_abi_version(void);
const char* mvb_last_error(void);

/* Sizes the caller must allocate (device memory) before mvb_s1_create.
 * Replaces Transformer.setup_caches (fast_model.py:136-148): KV = n_layer x {K,V} x rows x heads x
 * block_size x head_dim elements. */
size_t mvb_s1_kv_bytes(const mvb_s1_config* cfg);
size_t mvb_s1_workspace_bytes(const mvb_s1_config* cfg);

/* Build an engine over caller-owned device memory.  `offsets` is a HOST array of
 * MVB_S1_GLOBAL_TENSORS + n_layer*MVB_S1_LAYER_TENSORS byte offsets into d_arena.  d_workspace must
 * be zero-filled; with a bf16 cache the call zero-fills d_kv itself (the attention kernel reads whole 64-position tiles,
 * so positions that were never written must hold finite values).  Replaces build_model()/Transformer.__init__/setup_caches/setup_spk_cond_mask
 * (fast_inference_utils.py:324-352, fast_model.py:116-148); there is no JIT/compile step. */
int mvb_s1_create(const mvb_s1_config* cfg, const void* d_arena, size_t arena_bytes,
                  const uint64_t* offsets, void* d_kv, void* d_workspace, mvb_s1** out);
int mvb_s1_destroy(mvb_s1* h);

/* Hoisted speaker conditioning: spk_proj[utt] = W_spk * spk_emb (fast_model.py:152-157 recomputes it
 * every step; it only changes per utterance).  d_spk_emb: fp32 [spk_dim] on device. */
int mvb_s1_set_speaker(mvb_s1* h, int32_t utt, const float* d_spk_emb, void* stream);

/* Parity hook == Transformer.forward(idx[2,S], spk_emb, input_pos=arange(pos0, pos0+S)) for slot
 * `utt` (fast_model.py:150-163).  d_idx: int32 [2, S] on device (row 0 = cond, row 1 = uncond).
 * Writes the KV cache at positions pos0..pos0+S-1.  d_logits: fp32 [2, S, vocab] when
 * all_positions != 0, else fp32 [2, vocab] for the last position only. */
int mvb_s1_forward(mvb_s1* h, int32_t utt, const int32_t* d_idx, int32_t S, int32_t pos0,
                   float* d_logits, int32_t all_positions, void* stream);

/* == sample() (fast_inference_utils.py:107-120) on device.  d_logits fp32 [2, vocab];
 * d_noise: fp32 [vocab] Exp(1) draws (the reference's `q`, utils:64) or NULL for on-device Philox;
 * d_token_out int32 [1]; d_probs_out fp32 [vocab] or NULL. */
int mvb_s1_sample(mvb_s1* h, const float* d_logits, const mvb_sampling* p, const float* d_noise,
                  uint64_t step, int32_t* d_token_out, float* d_probs_out, void* stream);

/* Plugin call with HOST buffers == generate() (fast_inference_utils.py:181-228) for n_utts independent
 * utterances decoded together (per-row positions; semantics of mixins/causal.py:179-287, which equal
 * running each utterance alone).  Host->device copies of prompts/speaker vectors and the device->host
 * copy of the tokens happen inside this call.
 *   prompts      host int32, concatenated; prompt_lens[n_utts]
 *   spk_embs     host fp32 [n_utts, spk_dim]
 *   params       [n_utts]
 *   noise        NULL, or fp32 [n_utts, max_new_tokens, vocab] Exp(1) draws in the reference's call order
 *                (utils:61-65: one `q` per generated token).  noise_on_device == 0: HOST buffer, staged to the device
 *                one burst of decode steps at a time; != 0: DEVICE buffer used in place (how the Python shim hands
 *                over the draws it makes with torch's generator, so that torch.manual_seed reproduces the reference)
 *   forced       host or NULL: int32 [n_utts, max_new_tokens] teacher-forced feedback tokens (test hook)
 *   out_tokens   host int32 [n_utts, max_new_tokens]; out_lens[n_utts] = tokens produced (EOA included, utils:226)
 * Raises (returns MVB_ERR_PROMPT_TOO_LONG) when a prompt leaves no room: utils:203-204. */
int mvb_s1_generate(mvb_s1* h, int32_t n_utts, const int32_t* prompts, const int32_t* prompt_lens,
                    const float* spk_embs, const mvb_sampling* params, int32_t max_new_tokens,
                    const float* noise, int32_t noise_on_device, const int32_t* forced, int32_t* out_tokens,
                    int32_t* out_lens, void* stream);

/* Same loop with inputs already resident: prompts/speakers must have been installed with
 * mvb_s1_set_speaker + mvb_s1_forward (prefill).  Runs `n_steps` decode steps for utterances
 * [0, n_utts) entirely on device (sampler, EOA latch and position bump included) and returns without
 * a host sync.  Used by bench.py for the HBM-resident `value` and by the shim's generate(). */
int mvb_s1_decode(mvb_s1* h, int32_t n_utts, int32_t n_steps, void* stream);

/* Continuous batching (the reference serialises requests, serving.py:59-109): utterances enter and
 * leave KV slots between decode bursts of the persistent kernel.
 *   mvb_s1_admit   == the per-utterance prologue of generate() (utils:196-212) for slot `utt`: HOST prompt int32 [T] and
 *                  speaker vector fp32 [spk_dim] are uploaded, the prompt is prefilled and the first token is sampled;
 *                  d_noise: optional DEVICE Exp(1) draws [max_new_tokens, vocab] (NULL = on-device Philox, params->seed).
 *                  Returns MVB_ERR_PROMPT_TOO_LONG like generate().
 *   mvb_s1_decode  then advances every slot [0, n_slots) by a burst; finished / parked slots are skipped by the sampler.
 *   mvb_s1_poll    done flags + generated-token counts of slots [0, n_slots) (one device->host round trip).
 *   mvb_s1_release parks a slot (done latch set) until the next admit. */
int mvb_s1_admit(mvb_s1* h, int32_t utt, const int32_t* prompt, int32_t T, const float* spk_emb,
                 const mvb_sampling* params, int32_t max_new_tokens, const float* d_noise, void* stream);
int mvb_s1_release(mvb_s1* h, int32_t utt, void* stream);
int mvb_s1_poll(mvb_s1* h, int32_t n_slots, int32_t* done_out, int32_t* n_gen_out, void* stream);

/* Install per-utterance decode state after prefill: first token, sampling params, optional device
 * noise [max_new, vocab] / forced-token [max_new] buffers (NULL = none). */
int mvb_s1_begin(mvb_s1* h, int32_t utt, int32_t first_token, int32_t pos, const mvb_sampling* p,
                 const float* d_noise, const int32_t* d_forced, void* stream);

/* Read back decode state: tokens generated so far for `utt` (host buffer, capacity cap). */
int mvb_s1_fetch(mvb_s1* h, int32_t utt, int32_t* out_tokens, int32_t cap, int32_t* n_out,
                 int32_t* done, void* stream);

/* Parity hook for the persistent fused decode kernel: one decode position (== Transformer.forward with S = 1,
 * fast_model.py:150-163) for utterances [0, n_utts) from the state installed by mvb_s1_begin, through the single
 * persistent kernel; logits fp32 [2*n_utts, vocab] are copied to d_logits, nothing is sampled, positions do not
 * advance (the KV cache is appended at the current positions). */
int mvb_s1_step_logits(mvb_s1* h, int32_t n_utts, float* d_logits, void* stream);

/* Kernel launches issued by this handle since creation (bench.py "gpu_launches"). */
uint64_t mvb_s1_launch_count(const mvb_s1* h);

/* Test / debug hooks (not part of the drop-in surface; used by tests/ and tools/ only):
 *   mvb_s1_fetch_sampled : the sampler's own draws for `utt` (they differ from the fed-back tokens of mvb_s1_fetch only
 *                          under teacher forcing), host int32 [cap];
 *   mvb_s1_trace_fetch   : per-CTA clock64 phase stamps of the last persistent-kernel step (MVB_PC_TRACE=1), host
 *                          int64 [max_ctas, 512]; returns the number of CTAs copied. */
int mvb_s1_fetch_sampled(mvb_s1* h, int32_t utt, int32_t* out_tokens, int32_t cap, void* stream);
int mvb_s1_trace_fetch(mvb_s1* h, long long* out, int32_t max_ctas, void* stream);

/* Tensor-core linear operator (tcgen05 + TMA weight streaming), the building block of the batched /
 * prefill path and of the stage-2 model: y[R, M] (+)= f(x)[R, K] . W[M, K]^T, W bf16 row-major [out, in]
 * as every nn.Linear weight in the reference checkpoints (e.g. fast_model.py:194-195, layers/attn.py:61-64).
 * f = optional RMSNorm with bf16 gain (fast_model.py:250-261) when d_gain != NULL.  split_lo != 0 carries the
 * activations as two bf16 terms (hi + lo) so the result matches an fp32-activation product to ~1e-5.
 * K % 64 == 0, 1 <= R <= 128.  ksplit_override <= 0 selects the split automatically.  Synchronises `stream`. */
int mvb_linear(const void* d_W, int32_t M, int32_t K, const float* d_x, int32_t ldx, int32_t R,
               const void* d_gain, float eps, int32_t split_lo, int32_t ksplit_override, float* d_out,
               int32_t ldo, int32_t accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stage 2: the non-causal codebook-expansion model (fam/llm/model.py:195-314 with causal=False) and its sampler
 * (fam/llm/mixins/non_causal.py:15-67).  Shapes come from second_stage.pt["model_args"] (inference.py:124-128).
 * Supported: norm_type "rmsnorm", nonlinearity "swiglu", bias False, head size 64 or 128 (anything else fails loudly). */
typedef struct mvb_s2_config {
  int32_t n_layer, n_head, n_embd;
  int32_t hidden;          /* SwiGLU width (layers.py:51-53) */
  int32_t block_size;      /* sequence length t of the single non-causal pass (non_causal.py:31) */
  int32_t n_in;            /* input hierarchies (2) */
  int32_t vocab_in[8];
  int32_t n_out;           /* predicted hierarchies (6) */
  int32_t vocab_out[8];
  int32_t spk_dim;
  float   norm_eps;
  int32_t max_batch;
} mvb_s2_config;
typedef struct mvb_s2 mvb_s2;

size_t mvb_s2_workspace_bytes(const mvb_s2_config* cfg);
/* offsets (host): wte[n_in], wpe, speaker_cond_pos, ln_f, lm_heads[n_out], then per layer
 * {ln_1, attn.c_attn, attn.c_proj, ln_2, mlp.swiglu.w1, mlp.swiglu.w3, mlp.c_proj}; bf16 arena as for stage 1.
 * d_workspace must be zero-filled.  Replaces Model._init_model (inference.py:102-141). */
int mvb_s2_create(const mvb_s2_config* cfg, const void* d_arena, size_t arena_bytes, const uint64_t* offsets,
                  void* d_workspace, mvb_s2** out);
int mvb_s2_destroy(mvb_s2* h);
/* == GPT.generate -> _non_causal_sample for `batch` rows: d_idx int32 [batch, n_in, block_size] (built as
 * inference.py:283-306), d_spk fp32 [batch, spk_dim] or NULL, top_k <= 0 disables, d_noise fp32
 * [n_out, batch*block_size, vocab_out] = the Exp(1) draws torch.multinomial consumes (NULL = on-device Philox),
 * d_tokens int32 [batch, n_out, block_size]; d_logits_out optional fp32 [n_out, batch*block_size, vocab_out]. */
int mvb_s2_forward(mvb_s2* h, int32_t batch, const int32_t* d_idx, const float* d_spk, float temperature,
                   int32_t top_k, const float* d_noise, uint64_t seed, int32_t* d_tokens, float* d_logits_out,
                   void* stream);

/* ------------------------------------------------------------------------------------------------
 * Vocoder, EnCodec-24 kHz decode path: RVQ decode (the diffusion condition `decode_latent`) and the SEANet decoder
 * (causal conv stack + 2-layer LSTM), i.e. what mbd.tokens_to_wav (fam/llm/decoders.py:85, audiocraft 1.2.0 ->
 * transformers EncodecModel.decode) evaluates first.  The multi-band diffusion UNets themselves are NOT implemented
 * (their source/weights are unavailable: parity unpinned).  fp32 arena, weight norm folded. */
typedef struct mvb_voc_config {
  int32_t n_q;          /* 8 codebooks at 6 kbps */
  int32_t hidden;       /* codebook / latent dim 128 */
  int32_t n_filters;    /* 32 */
  int32_t n_ratios;     /* 4 */
  int32_t ratios[8];    /* 8, 5, 4, 2 */
  int32_t kernel, res_kernel, last_kernel; /* 7, 3, 7 */
  int32_t compress;     /* 2 */
  int32_t max_frames;   /* workspace capacity in 75 Hz frames */
} mvb_voc_config;
typedef struct mvb_voc mvb_voc;

size_t mvb_voc_workspace_bytes(const mvb_voc_config* cfg);
/* offsets (host), fp32 tensors: codebooks[n_q] [1024,hidden] | conv_in {w[512,hidden,7], b} | lstm layer 0
 * {w_ih, w_hh, b_ih+b_hh} | lstm layer 1 {..} | per ratio: up {w[Cin,Cout,2r], b}, res.block1 {w,b}, res.block3 {w,b},
 * res.shortcut {w,b} | conv_out {w[1,32,7], b}. */
int mvb_voc_create(const mvb_voc_config* cfg, const void* d_arena, size_t arena_bytes, const uint64_t* offsets,
                   void* d_workspace, mvb_voc** out);
int mvb_voc_destroy(mvb_voc* h);
/* d_codes int32 [n_q, T] -> d_latent fp32 [hidden, T]  (EncodecResidualVectorQuantizer.decode) */
int mvb_voc_decode_latent(mvb_voc* h, const int32_t* d_codes, int32_t T, float* d_latent, void* stream);
/* d_codes int32 [n_q, T] -> d_wav fp32 [T * prod(ratios)]  (EncodecModel.decode, one chunk, 24 kHz) */
int mvb_voc_decode(mvb_voc* h, const int32_t* d_codes, int32_t T, float* d_wav, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-band-diffusion vocoder (second half) == `mbd.tokens_to_wav` (fam/llm/decoders.py:84-85 ->
 * audiocraft 1.2.0 MultiBandDiffusion.tokens_to_wav) after the codec decode: n_models band UNets x n_calls sampler steps
 * conditioned on the codec latent, per-band output processors, sum, 32-band re-EQ against the EnCodec waveform.
 * PARITY UNPINNED (audiocraft / julius / mbd_comp_8.pt are not available): the configuration is a parameter. */
typedef struct mvb_mbd_config {
  int32_t n_models;        /* band models summed by MultiBandDiffusion.generate (4) */
  int32_t chin, hidden, depth, res_blocks, norm_groups, kernel, stride;   /* DiffusionUnet kwargs; (kernel, stride) in {(8,4), (4,2)} */
  float   growth;
  int32_t emb_all_layers;  /* step embedding added after every encoder level (else only the first) */
  int32_t codec_dim;       /* 128: channels of the condition (codec latent) */
  int32_t num_steps;       /* rows of the step-embedding tables (1000) */
  int32_t n_calls;         /* UNet evaluations per band model (20 for step_list = range(1000)[::-50] + [0]) */
  float   noise_scale, clip;
  int32_t proc_bands, proc_taps;   /* MultiBandProcessor: SplitBands(n_bands) and its filter length */
  int32_t eq_bands, eq_taps;       /* re_eq: 32 bands and the filter length */
  int32_t max_samples;     /* workspace capacity (samples at 24 kHz) */
} mvb_mbd_config;
typedef struct mvb_mbd mvb_mbd;
/* fp32 arena tensors (offsets, host): for every band model, in this order
 *   per encoder level: conv.weight [C, Cin, k], norm.weight, norm.bias, res_blocks x {norm1.w, norm1.b, conv1.w [C,C,3], conv1.b,
 *                      norm2.w, norm2.b, conv2.w, conv2.b}, step-embedding table [num_steps, C]
 *   conv_codec.weight [C_bottleneck, codec_dim, 1], conv_codec.bias
 *   per decoder level (deepest first): res_blocks x {8 tensors}, norm.weight, norm.bias, convtr.weight [C, C_out, k]
 *   processor scale [proc_bands] = (std / target_std) ** power_std, processor mean [proc_bands]
 * then three globals: processor low-pass bank [proc_bands - 1, proc_taps], re-EQ bank [eq_bands - 1, eq_taps],
 * schedule table [n_calls, 4] = {a, b, sigma, step}:  previous = clamp((current - a * estimate) * b + sigma * noise). */
size_t mvb_mbd_workspace_bytes(const mvb_mbd_config* cfg);
int mvb_mbd_create(const mvb_mbd_config* cfg, const void* d_arena, size_t arena_bytes, const uint64_t* offsets, void* d_workspace,
                   mvb_mbd** out);
int mvb_mbd_destroy(mvb_mbd* h);
/* d_cond fp32 [codec_dim, n_frames] (mvb_voc_decode_latent), d_wav_encodec fp32 [n_samples] (mvb_voc_decode),
 * d_noise fp32 [n_models, n_calls, n_samples] standard-normal draws (row 0 = initial sample, row i = the draw added after
 * call i - 1 ... see oracle/mbd_port.py) or NULL for on-device Philox (seed); d_wav_out fp32 [n_samples]. */
int mvb_mbd_tokens_to_wav(mvb_mbd* h, const float* d_cond, int32_t n_frames, const float* d_wav_encodec, int32_t n_samples,
                          const float* d_noise, uint64_t seed, float* d_wav_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Speaker encoder: fam/quantiser/audio/speaker_encoder/{audio.py:10-22, model.py:50-103}.
 * 16 kHz mono fp32 waveform -> power mel spectrogram (n_fft 400, hop 160, 40 Slaney bands, centered, zero padded) ->
 * 3-layer LSTM(40 -> 256) over partial windows of 160 frames -> linear -> ReLU -> L2 norm -> mean -> L2 norm. */
typedef struct mvb_spk_config {
  int32_t n_mels;          /* 40 */
  int32_t hidden;          /* 256 */
  int32_t n_layers;        /* 3 */
  int32_t emb;             /* 256 */
  int32_t n_fft;           /* 400 = 25 ms at 16 kHz */
  int32_t hop;             /* 160 = 10 ms */
  int32_t partial_frames;  /* 160 */
  int32_t max_samples;     /* workspace capacity (samples at 16 kHz) */
} mvb_spk_config;
typedef struct mvb_spk mvb_spk;
/* fp32 arena tensors, in this order: per LSTM layer {weight_ih [4H, in], weight_hh [4H, H], bias_ih + bias_hh [4H]},
 * linear.weight [256, 256], linear.bias [256], mel filterbank [n_mels, n_fft/2 + 1], analysis window [n_fft],
 * cos(2 pi i / n_fft) [n_fft], -sin(2 pi i / n_fft) [n_fft]. */
#define MVB_SPK_TENSORS 15
size_t mvb_spk_workspace_bytes(const mvb_spk_config* cfg);
int mvb_spk_create(const mvb_spk_config* cfg, const void* d_arena, size_t arena_bytes, const uint64_t* offsets,
                   void* d_workspace, mvb_spk** out);
int mvb_spk_destroy(mvb_spk* h);
/* == audio.wav_to_mel_spectrogram: d_wav fp32 [n_samples] -> d_mel fp32 [1 + n_samples / hop, n_mels] (parity hook). */
int mvb_spk_mel(mvb_spk* h, const float* d_wav, int32_t n_samples, float* d_mel, void* stream);
/* == SpeakerEncoder.embed_utterance on an already padded waveform: slice_starts (HOST int32 [n_partials]) are the first mel
 * frames of the partial windows (compute_partial_slices, model.py:55-79); d_embed fp32 [256]; d_partials_out optional
 * fp32 [n_partials, 256]. */
int mvb_spk_embed(mvb_spk* h, const float* d_wav, int32_t n_samples, const int32_t* slice_starts, int32_t n_partials,
                  float* d_embed, float* d_partials_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Audio post-processing on the device == `_save_audio` (fam/llm/decoders.py:40-47 -> audiocraft
 * audio_write(strategy="loudness", loudness_compressor=True)): BS.1770-4 integrated loudness exactly as
 * torchaudio.functional.loudness computes it (the function audiocraft calls), gain to -loudness_headroom_db LUFS, tanh
 * compressor, clip, PCM16.  d_wav fp32 mono [n_samples]; d_pcm16 int16 [n_samples]; d_wav_out optional fp32 [n_samples];
 * d_lkfs_gain optional fp32 [2] = {measured LKFS, applied linear gain}.  Signals below the 2e-3 RMS floor pass through. */
size_t mvb_audio_post_workspace_bytes(int32_t max_samples);
int mvb_audio_post(const float* d_wav, int32_t n_samples, int32_t sample_rate, float loudness_headroom_db,
                   int32_t loudness_compressor, void* d_workspace, int16_t* d_pcm16, float* d_wav_out,
                   float* d_lkfs_gain, void* stream);

#define MVB_OK 0
#define MVB_ERR_CUDA 1
#define MVB_ERR_ARG 2
#define MVB_ERR_PROMPT_TOO_LONG 3
#define MVB_ERR_UNSUPPORTED 4

#ifdef __cplusplus
}
#endif
#endif /* MVB200_H */
