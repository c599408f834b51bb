// libmvb200: host side of the stage-1 engine and its C ABI (include/mvb200.h).
// No torch, no CPU fallback: every entry point either runs the CUDA path or fails loudly.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/mvb200.h"
#include "stage1_kernels.cuh"
#include "umma_host.cuh"
#include "decode_persistent.cuh"

using namespace mvb;

static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
namespace mvb {
int set_error(int code, const char* fmt, ...) {  // shared with the other translation units
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
}  // namespace mvb

#define CK(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return fail(MVB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

static inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

struct WsLayout {
  size_t x, qkv, att, ffn, logits, spk, part_o, part_ml;
  size_t slot_map, pos, row_tok, done, n_gen, gen_tokens, sampled, samp, noise, forced, ticket, budget, noise_base;
  size_t stage_idx, stage_spk, stage_forced, stage_noise;
  // path B (tensor-core rows path): activations for up to RB_MAX rows
  size_t b_x, b_qkv, b_att, b_ffn, b_logits, b_last, b_B, b_rows, b_scratch, b_tickets, b_part_o, b_part_ml, b_att_tickets;
  // path C (persistent decode kernel)
  size_t c_x, c_qkv, c_gu, c_part_o, c_part_ml, c_bar, c_trace, c_stat, c_samp_part, c_samp_best;
  size_t total;
};

constexpr int RB_MAX = 128;        // rows per tensor-core pass: 64-token prefill chunk x 2 CFG rows, or 64 utterances
constexpr int PREFILL_CHUNK = 64;
constexpr int GEN_BURST = 32;       // decode steps between two host polls of the done flags in mvb_s1_generate

static WsLayout make_layout(const mvb_s1_config& c) {
  WsLayout L;
  const size_t U = c.max_utts, R = 2 * U, D = c.dim, F = c.intermediate, V = c.vocab, H = c.n_head;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes); return r; };
  L.x = take(R * D * 4);
  L.qkv = take(R * 3 * D * 4);
  L.att = take(R * D * 4);
  L.ffn = take(R * F * 4);
  L.logits = take(R * V * 4);
  L.spk = take(U * D * 4);
  L.part_o = take(R * H * ATT_SPLITS * 128 * 4);
  L.part_ml = take(R * H * ATT_SPLITS * 2 * 4);
  L.slot_map = take(U * 4);
  L.pos = take(U * 4);
  L.row_tok = take(R * 4);
  L.done = take(U * 4);
  L.n_gen = take(U * 4);
  L.gen_tokens = take(U * (size_t)c.max_new * 4);
  L.sampled = take(U * (size_t)c.max_new * 4);
  L.samp = take(U * sizeof(SamplingDev));
  L.noise = take(U * sizeof(void*));
  L.forced = take(U * sizeof(void*));
  L.ticket = take(R * H * 4);
  L.budget = take(U * 4);
  L.noise_base = take(U * 4);
  L.stage_idx = take(U * 2 * (size_t)c.block_size * 4);
  L.stage_spk = take(U * (size_t)c.spk_dim * 4);
  L.stage_forced = take(U * (size_t)c.max_new * 4);
  L.stage_noise = take(U * (size_t)GEN_BURST * V * 4);   // host-supplied Exp(1) draws are staged one burst at a time
  {
    const size_t RB = RB_MAX, NBM = 2 * RB_MAX;
    L.b_x = take(RB * D * 4);
    L.b_qkv = take(RB * 3 * D * 4);
    L.b_att = take(RB * D * 4);
    L.b_ffn = take(RB * F * 4);
    L.b_logits = take(RB * V * 4);
    L.b_last = take(2 * D * 4);
    L.b_B = take(NBM * (F > D ? F : D) * 2);
    L.b_rows = take(5 * RB * 4);
    size_t sc = 0;
    const size_t mats[5][3] = {{3 * D, D, 1}, {D, D, 1}, {F, D, 2}, {D, F, 1}, {V, D, 1}};
    for (auto& m : mats) {
      GemmPlan g = plan_gemm((int)m[0], (int)m[1], (int)NBM, m[2] == 2, 148);
      if (g.scratch_floats > sc) sc = g.scratch_floats;
    }
    L.b_scratch = take(sc * 4);
    L.b_tickets = take(256 * 4);
    L.b_part_o = take(RB * H * ATT_SPLITS * 128 * 4);
    L.b_part_ml = take(RB * H * ATT_SPLITS * 2 * 4);
    L.b_att_tickets = take(RB * H * 4);
  }
  L.c_x = take((size_t)PC_RPAD * D * 4);
  L.c_qkv = take((size_t)PC_RPAD * 3 * D * 4);
  L.c_gu = take((size_t)PC_RPAD * 2 * F * 4);
  L.c_part_o = take((size_t)PC_RPAD * H * PC_MAX_CHUNKS * 128 * 4);
  L.c_part_ml = take((size_t)PC_RPAD * H * PC_MAX_CHUNKS * 2 * 4);
  L.c_bar = take(256);
  L.c_trace = take((size_t)160 * PC_TRACE_EVENTS * 8);
  L.c_stat = take((size_t)(2 * c.n_layer + 1) * PC_RPAD * 4);
  L.c_samp_part = take((size_t)(PC_RPAD / 2) * 256 * 4);
  L.c_samp_best = take((size_t)2 * (PC_RPAD / 2) * 8);
  L.total = o;
  return L;
}

struct mvb_s1 {
  mvb_s1_config cfg;
  int n_sm = 148;
  const char* arena = nullptr;
  std::vector<uint64_t> off;
  char* kv = nullptr;
  char* ws = nullptr;
  WsLayout L;
  S1State st;
  std::map<int, cudaGraphExec_t> graphs;
  cudaStream_t cap_stream = nullptr;
  uint64_t launches = 0;
  std::map<int, int> graph_nodes;
  bool use_graph = true;
  int* h_flags = nullptr;  // pinned: done flags / counters read back by generate()
  // path B
  std::vector<CUtensorMap> tmW;                       // per layer {wqkv, wo, w1, w3, w2}, then lm_head
  std::map<std::pair<int, int>, CUtensorMap> tmB;     // (NB, K) -> activation buffer map
  bool path_b = true;                                 // tensor-core rows path for batched decode and prefill
  bool pdl = true;                                    // programmatic dependent launch between body kernels
  int decode_b_min = 2;                               // utterances from which decode uses the tensor-core rows path
  int split_lo = 1;                                   // carry activations as hi+lo bf16 terms
  RowsDev rows;
  // path C
  bool path_c = true;
  bool pc_ok = false;
  bool trace = false;
  // persistent-kernel switches, read once at create (MVB_PC_WB, MVB_PC_FUSED, MVB_PF_MODE, MVB_PF_AHEAD)
  bool wb = false;                                    // streamed weights as the N = 256 UMMA B operand (measured slower: off)
  bool fused = true;                                  // sample inside the persistent kernel (multi-token launches)
  int pf_mode = 1, pf_ahead = 8, epi_mode = 2, kv_pf = 0;
  int att_mma = 1;                                    // bf16 cache: attention on mma.sync over TMA-swizzled KV tiles (MVB_PC_ATT_MMA=0: scalar)
  CUtensorMap tm_kv;                                  // 2-D map over the whole bf16 KV arena: [rows = layer, k|v, cache row, head, pos][128]
  CUtensorMap tm3[2][6];                              // [WB] 3-D (k, row, layer) maps: wqkv, wo, w1, w3, w2, head
  PcMat pm[2][5];
  bool pc_ok2[2] = {false, false};
  std::vector<int> h_topk;                            // top_k of every utterance slot as installed by mvb_s1_begin
  size_t layer_stride_elems = 0;

  template <typename T>
  T* wsp(size_t o) const { return reinterpret_cast<T*>(ws + o); }
  const __nv_bfloat16* w(int i) const { return reinterpret_cast<const __nv_bfloat16*>(arena + off[i]); }
  const __nv_bfloat16* lw(int layer, int t) const { return w(MVB_S1_GLOBAL_TENSORS + layer * MVB_S1_LAYER_TENSORS + t); }
  size_t kv_elem() const { return cfg.kv_dtype == MVB_KV_FP32 ? 4 : 2; }
  size_t kv_half_bytes() const {
    return (size_t)2 * cfg.max_utts * cfg.n_head * cfg.block_size * cfg.head_dim * kv_elem();
  }
};


// Static decomposition of one matrix over `ctas` CTAs: the K split that minimises the busiest CTA's modelled phase time.
static PcMat plan_pc(int M, int K, int ctas, int tile_rows, int force_S = 0) {
  PcMat best{};
  long best_cost = -1;
  const int T = (M + tile_rows - 1) / tile_rows, KB = K / 64;
  for (int S = 1; S <= KB && S <= 32 && S <= ctas; ++S) {
    if (force_S > 0 && S != force_S) continue;   // MVB_PC_SPLITS experiment switch
    const int Gp = ctas / S;
    const int tiles_per = (T + Gp - 1) / Gp;
    const int kb_per = (KB + S - 1) / S;
    if (kb_per > PC_BKB_MAX) continue;
    // Cost of a phase in tile-load units (fitted on B200, profiles/r2_step_time_ksplit_sweep.jsonl): every tile a CTA owns
    // costs its k-blocks plus ~2.5 for the accumulator drain (tcgen05.ld + red.add), and staging the activation operand
    // costs ~2 per k-block of the CTA's K slice (one dependent L2 round trip per 256 chunks).  The round-1 cost
    // (tiles x k-blocks only) chose 3 / 8 / 8 / 9 splits for qkv / wo / w1|w3 / w2; this one chooses 8 / 16 / 8 / 22.
    const long cost = (long)tiles_per * (2 * kb_per + 5) + 4 * kb_per;
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best.T = T; best.KB = KB; best.S = S; best.G = Gp;
    }
  }
  if (best_cost < 0) best.S = 0;
  return best;
}

static int validate(const mvb_s1_config* c) {
  if (!c) return fail(MVB_ERR_ARG, "null config");
  if (c->head_dim != 128) return fail(MVB_ERR_UNSUPPORTED, "head_dim must be 128 (got %d)", c->head_dim);
  if (c->dim != c->n_head * c->head_dim) return fail(MVB_ERR_ARG, "dim != n_head*head_dim");
  if (c->dim % 256 || c->intermediate % 256) return fail(MVB_ERR_UNSUPPORTED, "dim and intermediate must be multiples of 256");
  if (c->vocab % 2 || c->vocab > SAMP_PAD) return fail(MVB_ERR_UNSUPPORTED, "vocab must be even and <= %d", SAMP_PAD);
  if (c->max_utts < 1 || c->max_utts > 64) return fail(MVB_ERR_ARG, "max_utts out of range");
  if (c->max_new < 1 || c->max_new > c->block_size) return fail(MVB_ERR_ARG, "max_new out of range");
  if (c->kv_dtype != MVB_KV_BF16 && c->kv_dtype != MVB_KV_FP32) return fail(MVB_ERR_ARG, "bad kv_dtype");
  if ((size_t)c->intermediate * 8 > 48 * 1024 - 256) return fail(MVB_ERR_UNSUPPORTED, "intermediate too large for the staging buffer");
  return MVB_OK;
}

extern "C" int mvb_abi_version(void) { return MVB_ABI_VERSION; }
extern "C" const char* mvb_last_error(void) { return g_err.c_str(); }

extern "C" size_t mvb_s1_kv_bytes(const mvb_s1_config* c) {
  if (validate(c)) return 0;
  const size_t esz = c->kv_dtype == MVB_KV_FP32 ? 4 : 2;
  return (size_t)c->n_layer * 2 * (2 * (size_t)c->max_utts) * c->n_head * c->block_size * c->head_dim * esz;
}

extern "C" size_t mvb_s1_workspace_bytes(const mvb_s1_config* c) {
  if (validate(c)) return 0;
  return make_layout(*c).total;
}

extern "C" int mvb_s1_create(const mvb_s1_config* cfg, const void* d_arena, size_t arena_bytes, const uint64_t* offsets,
                             void* d_kv, void* d_ws, mvb_s1** out) {
  if (int e = validate(cfg)) return e;
  if (!d_arena || !offsets || !d_kv || !d_ws || !out) return fail(MVB_ERR_ARG, "null pointer argument");
  int dev = 0, major = 0, minor = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  CK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) return fail(MVB_ERR_UNSUPPORTED, "libmvb200 is built for sm_100a only; device is sm_%d%d", major, minor);
  struct Guard {   // frees the half-built handle on every early return below
    mvb_s1* h;
    ~Guard() { if (h) mvb_s1_destroy(h); }
  } guard{new mvb_s1()};
  mvb_s1* h = guard.h;
  h->cfg = *cfg;
  CK(cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, dev));
  h->arena = reinterpret_cast<const char*>(d_arena);
  const int n_off = MVB_S1_GLOBAL_TENSORS + cfg->n_layer * MVB_S1_LAYER_TENSORS;
  h->off.assign(offsets, offsets + n_off);
  for (uint64_t o : h->off)
    if (o % 16 || o >= arena_bytes) return fail(MVB_ERR_ARG, "weight offset %llu not 16B-aligned or outside the arena", (unsigned long long)o);
  h->kv = reinterpret_cast<char*>(d_kv);
  h->ws = reinterpret_cast<char*>(d_ws);
  h->L = make_layout(*cfg);
  S1State& s = h->st;
  s.slot_map = h->wsp<int>(h->L.slot_map);
  s.pos = h->wsp<int>(h->L.pos);
  s.row_tok = h->wsp<int>(h->L.row_tok);
  s.done = h->wsp<int>(h->L.done);
  s.n_gen = h->wsp<int>(h->L.n_gen);
  s.gen_tokens = h->wsp<int>(h->L.gen_tokens);
  s.sampled_tokens = h->wsp<int>(h->L.sampled);
  s.samp = h->wsp<SamplingDev>(h->L.samp);
  s.noise = h->wsp<const float*>(h->L.noise);
  s.forced = h->wsp<const int*>(h->L.forced);
  s.attn_ticket = h->wsp<unsigned>(h->L.ticket);
  s.budget = h->wsp<int>(h->L.budget);
  s.noise_base = h->wsp<int>(h->L.noise_base);
  s.max_new = cfg->max_new;
  s.block_size = cfg->block_size;
  h->use_graph = getenv("MVB_NO_GRAPH") == nullptr;
  if (const char* e = getenv("MVB_PATHB")) h->path_b = atoi(e) != 0;
  if (const char* e = getenv("MVB_PDL")) h->pdl = atoi(e) != 0;
  if (const char* e = getenv("MVB_DECODE_B_MIN")) h->decode_b_min = atoi(e);
  if (const char* e = getenv("MVB_SPLIT_LO")) h->split_lo = atoi(e) != 0;
  {
    int* rb = h->wsp<int>(h->L.b_rows);
    h->rows.cache_row = rb; h->rows.pos = rb + RB_MAX; h->rows.tok = rb + 2 * RB_MAX; h->rows.utt = rb + 3 * RB_MAX;
    h->rows.cond = rb + 4 * RB_MAX;
    const int D = cfg->dim, F = cfg->intermediate;
    h->tmW.resize((size_t)cfg->n_layer * 5 + 1);
    bool ok = true;
    for (int l = 0; l < cfg->n_layer && ok; ++l) {
      ok = ok && make_tmap_bf16(&h->tmW[l * 5 + 0], h->lw(l, 1), 3 * D, D, 128);
      ok = ok && make_tmap_bf16(&h->tmW[l * 5 + 1], h->lw(l, 2), D, D, 128);
      ok = ok && make_tmap_bf16(&h->tmW[l * 5 + 2], h->lw(l, 4), F, D, 128);
      ok = ok && make_tmap_bf16(&h->tmW[l * 5 + 3], h->lw(l, 5), F, D, 128);
      ok = ok && make_tmap_bf16(&h->tmW[l * 5 + 4], h->lw(l, 6), D, F, 128);
    }
    ok = ok && make_tmap_bf16(&h->tmW[(size_t)cfg->n_layer * 5], h->w(4), cfg->vocab, D, 128);
    if (!ok) return fail(MVB_ERR_CUDA, "cuTensorMapEncodeTiled failed for a weight matrix");
  }
  if (const char* e = getenv("MVB_PATHC")) h->path_c = atoi(e) != 0;
  if (const char* e = getenv("MVB_PC_TRACE")) h->trace = atoi(e) != 0;
  if (const char* e = getenv("MVB_PF_AHEAD")) h->pf_ahead = atoi(e);
  if (const char* e = getenv("MVB_PF_MODE")) h->pf_mode = atoi(e);
  if (const char* e = getenv("MVB_PC_EPI")) h->epi_mode = atoi(e);
  if (const char* e = getenv("MVB_PC_KVPF")) h->kv_pf = atoi(e);
  if (const char* e = getenv("MVB_PC_ATT_MMA")) h->att_mma = atoi(e);
  memset(&h->tm_kv, 0, sizeof(h->tm_kv));
  if (cfg->kv_dtype != MVB_KV_BF16 || cfg->block_size % 64 || cfg->head_dim != 128) h->att_mma = 0;
  if (h->att_mma) {
    // the tensor-core attention reads whole 64-position boxes: positions past the valid ones must hold finite values
    CK(cudaMemset(h->kv, 0, (size_t)cfg->n_layer * 2 * h->kv_half_bytes()));
    EncodeTiledFn fn = encode_tiled_fn();
    const uint64_t rows = (uint64_t)cfg->n_layer * 2 * (h->kv_half_bytes() / 256);
    cuuint64_t dims[2] = {128, rows};
    cuuint64_t strides[1] = {256};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    if (!fn || rows >= (1ull << 31) ||
        fn(&h->tm_kv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, h->kv, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      h->att_mma = 0;
  }
  if (const char* e = getenv("MVB_PC_WB")) h->wb = atoi(e) != 0;
  if (const char* e = getenv("MVB_PC_FUSED")) h->fused = atoi(e) != 0;
  h->h_topk.assign(cfg->max_utts, 0);
  {
    // persistent decode kernel: needs a uniform layer stride (true for arenas packed in checkpoint order)
    const int D = cfg->dim, F = cfg->intermediate, V = cfg->vocab;
    bool ok = cfg->n_layer >= 1;
    const int o0 = MVB_S1_GLOBAL_TENSORS;
    uint64_t stride = cfg->n_layer > 1 ? h->off[o0 + MVB_S1_LAYER_TENSORS] - h->off[o0] : 2 * (uint64_t)(4 * D * D + 3 * D * F + 2 * D);
    for (int l = 1; l < cfg->n_layer && ok; ++l)
      for (int t = 0; t < MVB_S1_LAYER_TENSORS; ++t)
        ok = ok && (h->off[o0 + l * MVB_S1_LAYER_TENSORS + t] - h->off[o0 + (l - 1) * MVB_S1_LAYER_TENSORS + t] == stride);
    ok = ok && (stride % 16 == 0);
    h->layer_stride_elems = stride / 2;
    const uint64_t NL = cfg->n_layer;
    for (int wbv = 0; wbv < 2; ++wbv) {
      const int rows = wbv ? 256 : 128;
      bool okv = ok;
      okv = okv && make_tmap_bf16_3d(&h->tm3[wbv][0], h->lw(0, 1), D, 3 * D, NL, stride, rows);
      okv = okv && make_tmap_bf16_3d(&h->tm3[wbv][1], h->lw(0, 2), D, D, NL, stride, rows);
      okv = okv && make_tmap_bf16_3d(&h->tm3[wbv][2], h->lw(0, 4), D, F, NL, stride, rows);
      okv = okv && make_tmap_bf16_3d(&h->tm3[wbv][3], h->lw(0, 5), D, F, NL, stride, rows);
      okv = okv && make_tmap_bf16_3d(&h->tm3[wbv][4], h->lw(0, 6), F, D, NL, stride, rows);
      okv = okv && make_tmap_bf16_3d(&h->tm3[wbv][5], h->w(4), D, V, 1, (uint64_t)V * D * 2, rows);
      int fs[5] = {0, 0, 0, 0, 0};               // MVB_PC_SPLITS="q_o_w13_w2_head": force the K split of a matrix (0 = planner)
      if (const char* e = getenv("MVB_PC_SPLITS")) sscanf(e, "%d_%d_%d_%d_%d", &fs[0], &fs[1], &fs[2], &fs[3], &fs[4]);
      h->pm[wbv][0] = plan_pc(3 * D, D, h->n_sm, rows, fs[0]);
      h->pm[wbv][1] = plan_pc(D, D, h->n_sm, rows, fs[1]);
      h->pm[wbv][2] = plan_pc(2 * F, D, h->n_sm, rows, fs[2]);
      h->pm[wbv][3] = plan_pc(D, F, h->n_sm, rows, fs[3]);
      h->pm[wbv][4] = plan_pc(V, D, h->n_sm, rows, fs[4]);
      for (int i = 0; i < 5; ++i) okv = okv && h->pm[wbv][i].S > 0;
      okv = okv && (F % rows == 0) && (D % rows == 0);     // w1 | w3 share one tile index space: F must be whole tiles
      h->pc_ok2[wbv] = okv;
    }
    if (!h->pc_ok2[1]) h->wb = false;
    ok = h->pc_ok2[h->wb ? 1 : 0];
    h->pc_ok = ok;
  }
  // Opt-in shared-memory size is a per-device function attribute: set it for the device this handle lives on
  // (a process may hold engines on several GPUs).
#define MVB_PC_ATTR(FP, NBV, WBV)                                                                                   \
  CK(cudaFuncSetAttribute(k_decode_persistent<FP, NBV, WBV>, cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                          (int)PcCfg<NBV, WBV>::SMEM))
  MVB_PC_ATTR(true, 32, false); MVB_PC_ATTR(true, 16, false); MVB_PC_ATTR(false, 32, false); MVB_PC_ATTR(false, 16, false);
  MVB_PC_ATTR(true, 32, true); MVB_PC_ATTR(true, 16, true); MVB_PC_ATTR(false, 32, true); MVB_PC_ATTR(false, 16, true);
#undef MVB_PC_ATTR
  CK(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
  CK(cudaMallocHost(&h->h_flags, sizeof(int) * 4 * 64));
  guard.h = nullptr;
  *out = h;
  return MVB_OK;
}

extern "C" int mvb_s1_destroy(mvb_s1* h) {
  if (!h) return MVB_OK;
  for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  if (h->h_flags) cudaFreeHost(h->h_flags);
  delete h;
  return MVB_OK;
}

extern "C" uint64_t mvb_s1_launch_count(const mvb_s1* h) { return h ? h->launches : 0; }

// ------------------------------------------------------------------------------------------------
static int gemv_grid(int n_items, int n_sm) {
  const int cap = 2 * n_sm * 8;                       // resident warps at 2 CTAs/SM
  const int ipw = (n_items + cap - 1) / cap;          // items per warp
  return (n_items + 8 * ipw - 1) / (8 * ipw);
}

// Launch with the programmatic-dependent-launch attribute: the kernel may start while its predecessor in the
// stream drains; it prefetches weights, then griddepcontrol.wait orders it after the predecessor's writes.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <int EPI>
static cudaError_t launch_gemv(mvb_s1* h, cudaStream_t s, int n_utts, GemvP p) {
  const int n_items = (EPI == EPI_SWIGLU) ? p.M : p.M / 2;
  dim3 grid(gemv_grid(n_items, h->n_sm), n_utts);
  const size_t smem = (size_t)2 * p.K * sizeof(float);
  h->launches++;
  return launch_pdl(h->pdl, k_gemv<EPI>, grid, dim3(256), smem, s, p, h->st);
}

// One forward position for `n_utts` logical utterances: embed -> 24 x (attention, FFN) -> head.
static int launch_body(mvb_s1* h, cudaStream_t s, int n_utts) {
  const mvb_s1_config& c = h->cfg;
  const int D = c.dim, F = c.intermediate, V = c.vocab, H = c.n_head;
  float* x = h->wsp<float>(h->L.x);
  float* qkv = h->wsp<float>(h->L.qkv);
  float* att = h->wsp<float>(h->L.att);
  float* ffn = h->wsp<float>(h->L.ffn);
  float* logits = h->wsp<float>(h->L.logits);
  CK(launch_pdl(h->pdl, k_embed, dim3(2, n_utts), dim3(256), 0, s, h->st, h->w(0), h->w(1),
                (const float*)h->wsp<float>(h->L.spk), x, D));
  h->launches++;
  const size_t half = h->kv_half_bytes();
  for (int l = 0; l < c.n_layer; ++l) {
    char* kc = h->kv + (size_t)l * 2 * half;
    char* vc = kc + half;
    GemvP p{};
    p.eps = c.norm_eps;
    // attention_norm + wqkv + cache scatter
    p.W = h->lw(l, 1); p.W3 = nullptr; p.x = x; p.ldx = D; p.gain = h->lw(l, 0);
    p.out = qkv; p.ldo = 3 * D; p.M = 3 * D; p.K = D;
    p.kcache = kc; p.vcache = vc; p.H = H; p.S_max = c.block_size; p.D = D; p.kv_fp32 = c.kv_dtype == MVB_KV_FP32;
    CK(launch_gemv<EPI_QKV>(h, s, n_utts, p));
    // attention over [0, pos]
    dim3 ag(H, 2 * n_utts, ATT_SPLITS);
    if (c.kv_dtype == MVB_KV_FP32)
      CK(launch_pdl(h->pdl, k_attn_decode<true>, ag, dim3(128), 0, s, h->st, (const float*)qkv, (const void*)kc, (const void*)vc,
                    h->wsp<float>(h->L.part_o), h->wsp<float>(h->L.part_ml), att, H, c.block_size, D));
    else
      CK(launch_pdl(h->pdl, k_attn_decode<false>, ag, dim3(128), 0, s, h->st, (const float*)qkv, (const void*)kc, (const void*)vc,
                    h->wsp<float>(h->L.part_o), h->wsp<float>(h->L.part_ml), att, H, c.block_size, D));
    h->launches++;
    // wo + residual
    p = GemvP{};
    p.eps = c.norm_eps;
    p.W = h->lw(l, 2); p.x = att; p.ldx = D; p.gain = nullptr; p.out = x; p.ldo = D; p.M = D; p.K = D;
    CK(launch_gemv<EPI_RESID>(h, s, n_utts, p));
    // ffn_norm + silu(w1 x) * w3 x
    p.W = h->lw(l, 4); p.W3 = h->lw(l, 5); p.x = x; p.ldx = D; p.gain = h->lw(l, 3); p.out = ffn; p.ldo = F; p.M = F; p.K = D;
    CK(launch_gemv<EPI_SWIGLU>(h, s, n_utts, p));
    // w2 + residual
    p.W = h->lw(l, 6); p.W3 = nullptr; p.x = ffn; p.ldx = F; p.gain = nullptr; p.out = x; p.ldo = D; p.M = D; p.K = F;
    CK(launch_gemv<EPI_RESID>(h, s, n_utts, p));
  }
  GemvP p{};
  p.eps = c.norm_eps;
  p.W = h->w(4); p.x = x; p.ldx = D; p.gain = h->w(3); p.out = logits; p.ldo = V; p.M = V; p.K = D;
  CK(launch_gemv<EPI_STORE>(h, s, n_utts, p));
  return MVB_OK;
}


// ------------------------------------------------------------------------------------------------
// Path B: the same forward pass for R arbitrary activation rows on the tensor cores (umma_gemm.cuh).
// Used for batched decode (R = 2 * n_utts) and for prefill chunks (R = 2 * tokens).
static const CUtensorMap* tmap_b(mvb_s1* h, int NB, int K) {
  auto key = std::make_pair(NB, K);
  auto it = h->tmB.find(key);
  if (it == h->tmB.end()) {
    CUtensorMap tm;
    if (!make_tmap_bf16(&tm, h->wsp<char>(h->L.b_B), (uint64_t)NB, (uint64_t)K, (uint32_t)NB)) return nullptr;
    it = h->tmB.emplace(key, tm).first;
  }
  return &it->second;
}

struct LinB {
  const float* x; int ldx; const __nv_bfloat16* gain;  // input rows (+ optional RMSNorm gain)
  int widx, widx3;                                     // tensor-map indices of W (and W3 for SwiGLU)
  int M, K;
  float* out; int ldo;
};

template <int EPI>
static int linear_b(mvb_s1* h, cudaStream_t s, int R, const LinB& a, void* kc = nullptr, void* vc = nullptr) {
  const mvb_s1_config& c = h->cfg;
  const int Rpad = (R + 15) / 16 * 16;
  const int NB = Rpad * (h->split_lo ? 2 : 1);
  const CUtensorMap* tB = tmap_b(h, NB, a.K);
  if (!tB) return fail(MVB_ERR_CUDA, "cuTensorMapEncodeTiled failed for the activation buffer");
  __nv_bfloat16* B = h->wsp<__nv_bfloat16>(h->L.b_B);
  k_prep_b<<<Rpad, 256, 0, s>>>(a.x, a.ldx, a.gain, c.norm_eps, a.K, Rpad, R, h->split_lo, B);
  h->launches++;
  CK(cudaGetLastError());
  GemmP p{};
  p.M = a.M; p.K = a.K; p.NB = NB; p.Rpad = Rpad; p.R = R; p.split_lo = h->split_lo;
  p.scratch = h->wsp<float>(h->L.b_scratch);
  p.tickets = h->wsp<unsigned>(h->L.b_tickets);
  p.out = a.out; p.ldo = a.ldo;
  p.kcache = kc; p.vcache = vc; p.row_cache = h->rows.cache_row; p.row_pos = h->rows.pos;
  p.H = c.n_head; p.S_max = c.block_size; p.D = c.dim; p.kv_fp32 = c.kv_dtype == MVB_KV_FP32;
  const GemmPlan g = plan_gemm(a.M, a.K, NB, EPI == G_SWIGLU, h->n_sm);
  CK(launch_umma_gemm<EPI>(s, h->tmW[a.widx], h->tmW[a.widx3], *tB, p, g));
  h->launches++;
  return MVB_OK;
}

// embed + all layers for the R rows described by h->rows (already filled on the device)
static int launch_layers_b(mvb_s1* h, cudaStream_t s, int R) {
  const mvb_s1_config& c = h->cfg;
  const int D = c.dim, F = c.intermediate, H = c.n_head;
  float* x = h->wsp<float>(h->L.b_x);
  float* qkv = h->wsp<float>(h->L.b_qkv);
  float* att = h->wsp<float>(h->L.b_att);
  float* ffn = h->wsp<float>(h->L.b_ffn);
  k_embed_rows<<<R, 256, 0, s>>>(h->rows, h->w(0), h->w(1), h->wsp<float>(h->L.spk), x, D);
  h->launches++;
  CK(cudaGetLastError());
  const size_t half = h->kv_half_bytes();
  for (int l = 0; l < c.n_layer; ++l) {
    char* kc = h->kv + (size_t)l * 2 * half;
    char* vc = kc + half;
    if (int e = linear_b<G_QKV>(h, s, R, LinB{x, D, h->lw(l, 0), l * 5 + 0, l * 5 + 0, 3 * D, D, qkv, 3 * D}, kc, vc)) return e;
    dim3 ag(H, R, ATT_SPLITS);
    if (c.kv_dtype == MVB_KV_FP32)
      k_attn_rows<true><<<ag, 128, 0, s>>>(h->rows, qkv, kc, vc, h->wsp<float>(h->L.b_part_o), h->wsp<float>(h->L.b_part_ml),
                                           h->wsp<unsigned>(h->L.b_att_tickets), att, H, c.block_size, D);
    else
      k_attn_rows<false><<<ag, 128, 0, s>>>(h->rows, qkv, kc, vc, h->wsp<float>(h->L.b_part_o), h->wsp<float>(h->L.b_part_ml),
                                            h->wsp<unsigned>(h->L.b_att_tickets), att, H, c.block_size, D);
    h->launches++;
    CK(cudaGetLastError());
    if (int e = linear_b<G_RESID>(h, s, R, LinB{att, D, nullptr, l * 5 + 1, l * 5 + 1, D, D, x, D})) return e;
    if (int e = linear_b<G_SWIGLU>(h, s, R, LinB{x, D, h->lw(l, 3), l * 5 + 2, l * 5 + 3, F, D, ffn, F})) return e;
    if (int e = linear_b<G_RESID>(h, s, R, LinB{ffn, F, nullptr, l * 5 + 4, l * 5 + 4, D, F, x, D})) return e;
  }
  return MVB_OK;
}

// batched decode position: rows from the decode state, logits land in the sampler's buffer
static int launch_body_b(mvb_s1* h, cudaStream_t s, int n_utts) {
  const mvb_s1_config& c = h->cfg;
  const int R = 2 * n_utts;
  k_rows_decode<<<1, 128, 0, s>>>(h->st, h->rows, n_utts);
  h->launches++;
  CK(cudaGetLastError());
  if (int e = launch_layers_b(h, s, R)) return e;
  const int hw = c.n_layer * 5;
  return linear_b<G_STORE>(h, s, R, LinB{h->wsp<float>(h->L.b_x), c.dim, h->w(3), hw, hw, c.vocab, c.dim,
                                         h->wsp<float>(h->L.logits), c.vocab});
}


// ------------------------------------------------------------------------------------------------
// Path C: one persistent kernel per decode position (decode_persistent.cuh) followed by the sampler.
static int launch_persistent(mvb_s1* h, cudaStream_t s, int n_utts, int n_steps, bool fused) {
  const mvb_s1_config& c = h->cfg;
  const int wbv = h->wb ? 1 : 0;
  PcParams p{};
  p.n_layer = c.n_layer; p.D = c.dim; p.F = c.intermediate; p.V = c.vocab; p.H = c.n_head; p.S_max = c.block_size;
  p.n_steps = n_steps; p.fused = fused ? 1 : 0;
  p.pf_mode = h->pf_mode; p.pf_ahead = h->pf_ahead; p.epi_mode = h->epi_mode; p.kv_pf = h->kv_pf;
  p.R = 2 * n_utts; p.n_utts = n_utts; p.kv_fp32 = c.kv_dtype == MVB_KV_FP32; p.eps = c.norm_eps;
  p.m_qkv = h->pm[wbv][0]; p.m_o = h->pm[wbv][1]; p.m_w13 = h->pm[wbv][2]; p.m_w2 = h->pm[wbv][3]; p.m_head = h->pm[wbv][4];
  p.attn_norm = h->lw(0, 0); p.ffn_norm = h->lw(0, 3); p.out_norm = h->w(3);
  p.layer_stride = h->layer_stride_elems;
  p.tok_emb = h->w(0); p.pos_emb = h->w(1); p.spk_proj = h->wsp<float>(h->L.spk);
  p.x = h->wsp<float>(h->L.c_x); p.qkv = h->wsp<float>(h->L.c_qkv); p.gu = h->wsp<float>(h->L.c_gu);
  p.logits = h->wsp<float>(h->L.logits);
  p.part_o = h->wsp<float>(h->L.c_part_o); p.part_ml = h->wsp<float>(h->L.c_part_ml);
  p.stat = h->wsp<float>(h->L.c_stat);
  p.samp_part = h->wsp<float>(h->L.c_samp_part);
  p.samp_best = h->wsp<unsigned long long>(h->L.c_samp_best);
  p.kv = h->kv; p.kv_half = h->kv_half_bytes();
  p.att_mma = h->att_mma; p.kv_rows_half = (long long)(h->kv_half_bytes() / 256);
  p.bar = h->wsp<unsigned>(h->L.c_bar);
  p.trace = (h->trace && h->n_sm <= 160) ? h->wsp<long long>(h->L.c_trace) : nullptr;
  p.st = h->st;
  // launch-time invariants of the kernel: grid-barrier counter, RMSNorm statistics and arg-max slots start from zero
  CK(cudaMemsetAsync(p.bar, 0, 4, s));
  CK(cudaMemsetAsync(p.stat, 0, sizeof(float) * (size_t)(2 * c.n_layer + 1) * PC_RPAD, s));
  CK(cudaMemsetAsync(p.samp_best, 0, sizeof(unsigned long long) * 2 * (PC_RPAD / 2), s));
  h->launches++;
  const bool fp = p.kv_fp32 != 0, wide = p.R > 8;
  const CUtensorMap* tm = h->tm3[wbv];
#define MVB_PC_LAUNCH(FP, NBV, WBV)                                                                                   \
  CK(launch_pdl(h->pdl, k_decode_persistent<FP, NBV, WBV>, dim3(h->n_sm), dim3(PC_THREADS), PcCfg<NBV, WBV>::SMEM, s, \
                tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], h->tm_kv, p))
  if (wbv) {
    if (fp && wide) MVB_PC_LAUNCH(true, 32, true);
    else if (fp) MVB_PC_LAUNCH(true, 16, true);
    else if (wide) MVB_PC_LAUNCH(false, 32, true);
    else MVB_PC_LAUNCH(false, 16, true);
  } else {
    if (fp && wide) MVB_PC_LAUNCH(true, 32, false);
    else if (fp) MVB_PC_LAUNCH(true, 16, false);
    else if (wide) MVB_PC_LAUNCH(false, 32, false);
    else MVB_PC_LAUNCH(false, 16, false);
  }
#undef MVB_PC_LAUNCH
  return MVB_OK;
}

static int run_body(mvb_s1* h, cudaStream_t s, int n_utts, bool allow_b = true) {
  // Path A (CUDA-core GEMV) streams the weights once per utterance; from 2 utterances up the tensor-core
  // rows path streams them once per step for the whole batch.  slot_map must be the identity for path B
  // (its logits rows are batch-ordered), which mvb_s1_decode guarantees.
  const bool use_b = allow_b && h->path_b && n_utts >= h->decode_b_min;
  auto body = [&](cudaStream_t st) { return use_b ? launch_body_b(h, st, n_utts) : launch_body(h, st, n_utts); };
  if (!h->use_graph) return body(s);
  const int key = n_utts + (use_b ? 1000 : 0);
  auto it = h->graphs.find(key);
  if (it == h->graphs.end()) {
    cudaGraph_t g;
    const uint64_t before = h->launches;
    CK(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
    int e = body(h->cap_stream);
    cudaError_t ce = cudaStreamEndCapture(h->cap_stream, &g);
    h->graph_nodes[key] = (int)(h->launches - before);
    h->launches = before;
    if (e) return e;
    CK(ce);
    cudaGraphExec_t ge;
    CK(cudaGraphInstantiate(&ge, g, 0));
    CK(cudaGraphDestroy(g));
    it = h->graphs.emplace(key, ge).first;
  }
  CK(cudaGraphLaunch(it->second, s));
  h->launches += h->graph_nodes[key];
  return MVB_OK;
}

static SamplingDev to_dev(const mvb_sampling* p) {
  SamplingDev d;
  d.guidance = p->guidance_scale;
  d.temperature = p->temperature;
  d.top_p = p->top_p;
  d.top_k = p->top_k;
  d.end_of_audio = p->end_of_audio;
  d.seed = p->seed;
  return d;
}

extern "C" int mvb_s1_set_speaker(mvb_s1* h, int32_t utt, const float* d_spk, void* stream) {
  if (!h || !d_spk) return fail(MVB_ERR_ARG, "null argument");
  if (utt < 0 || utt >= h->cfg.max_utts) return fail(MVB_ERR_ARG, "utterance slot %d out of range", utt);
  cudaStream_t s = (cudaStream_t)stream;
  const int D = h->cfg.dim;
  k_spk_proj<<<(D * 32 + 255) / 256, 256, 0, s>>>(h->w(2), d_spk, h->wsp<float>(h->L.spk) + (size_t)utt * D, D, h->cfg.spk_dim);
  h->launches++;
  CK(cudaGetLastError());
  return MVB_OK;
}

extern "C" int mvb_s1_forward(mvb_s1* h, int32_t utt, const int32_t* d_idx, int32_t S, int32_t pos0, float* d_logits,
                              int32_t all_positions, void* stream) {
  if (!h || !d_idx) return fail(MVB_ERR_ARG, "null argument");
  if (utt < 0 || utt >= h->cfg.max_utts) return fail(MVB_ERR_ARG, "utterance slot %d out of range", utt);
  if (S < 1 || pos0 < 0 || pos0 + S > h->cfg.block_size)
    return fail(MVB_ERR_ARG, "positions [%d, %d) outside the %d-slot context", pos0, pos0 + S, h->cfg.block_size);
  cudaStream_t s = (cudaStream_t)stream;
  const int V = h->cfg.vocab;
  const float* lg = h->wsp<float>(h->L.logits) + (size_t)(2 * utt) * V;
  if (h->path_b && S >= 2) {
    // Prefill on the tensor cores in chunks of PREFILL_CHUNK tokens: every weight matrix is streamed once per
    // chunk for all 2*Sc rows (the reference prefill, utils:123-132, does the same with a [2,T] batch).
    const mvb_s1_config& c = h->cfg;
    for (int s0 = 0; s0 < S; s0 += PREFILL_CHUNK) {
      const int Sc = (S - s0) < PREFILL_CHUNK ? (S - s0) : PREFILL_CHUNK;
      const int R = 2 * Sc;
      k_rows_prefill<<<1, 128, 0, s>>>(h->rows, utt, d_idx, S, s0, Sc, pos0);
      h->launches++;
      CK(cudaGetLastError());
      if (int e = launch_layers_b(h, s, R)) return e;
      const int hw = c.n_layer * 5;
      float* bx = h->wsp<float>(h->L.b_x);
      if (d_logits && all_positions) {
        float* bl = h->wsp<float>(h->L.b_logits);
        if (int e = linear_b<G_STORE>(h, s, R, LinB{bx, c.dim, h->w(3), hw, hw, V, c.dim, bl, V})) return e;
        for (int cc = 0; cc < 2; ++cc)
          CK(cudaMemcpyAsync(d_logits + ((size_t)cc * S + s0) * V, bl + (size_t)cc * Sc * V, sizeof(float) * (size_t)Sc * V,
                             cudaMemcpyDeviceToDevice, s));
      }
      if (s0 + Sc == S) {
        // last position -> the sampler's logits rows of this utterance (and the decode state's position)
        float* last = h->wsp<float>(h->L.b_last);
        k_gather_last<<<2, 256, 0, s>>>(bx, last, Sc, c.dim);
        h->launches++;
        CK(cudaGetLastError());
        if (int e = linear_b<G_STORE>(h, s, 2, LinB{last, c.dim, h->w(3), hw, hw, V, c.dim,
                                                    h->wsp<float>(h->L.logits) + (size_t)(2 * utt) * V, V})) return e;
        k_set_input<<<1, 32, 0, s>>>(h->st, utt, d_idx, S, S - 1, pos0 + S - 1);
        h->launches++;
        CK(cudaGetLastError());
      }
    }
    if (d_logits && !all_positions) CK(cudaMemcpyAsync(d_logits, lg, sizeof(float) * 2 * V, cudaMemcpyDeviceToDevice, s));
    return MVB_OK;
  }
  for (int i = 0; i < S; ++i) {
    k_set_input<<<1, 32, 0, s>>>(h->st, utt, d_idx, S, i, pos0 + i);
    h->launches++;
    CK(cudaGetLastError());
    if (int e = run_body(h, s, 1, false)) return e;
    if (d_logits && all_positions) {
      CK(cudaMemcpyAsync(d_logits + (size_t)i * V, lg, sizeof(float) * V, cudaMemcpyDeviceToDevice, s));
      CK(cudaMemcpyAsync(d_logits + ((size_t)S + i) * V, lg + V, sizeof(float) * V, cudaMemcpyDeviceToDevice, s));
    }
  }
  if (d_logits && !all_positions) CK(cudaMemcpyAsync(d_logits, lg, sizeof(float) * 2 * V, cudaMemcpyDeviceToDevice, s));
  return MVB_OK;
}

extern "C" int mvb_s1_sample(mvb_s1* h, const float* d_logits, const mvb_sampling* p, const float* d_noise, uint64_t step,
                             int32_t* d_token_out, float* d_probs_out, void* stream) {
  if (!h || !d_logits || !p || !d_token_out) return fail(MVB_ERR_ARG, "null argument");
  SampleP sp{};
  sp.logits = d_logits;
  sp.V = h->cfg.vocab;
  sp.decode_mode = 0;
  sp.sp = to_dev(p);
  sp.noise = d_noise;
  sp.step = step;
  sp.token_out = d_token_out;
  sp.probs_out = d_probs_out;
  k_sample<<<1, SAMP_THREADS, 0, (cudaStream_t)stream>>>(sp, h->st);
  h->launches++;
  CK(cudaGetLastError());
  return MVB_OK;
}

extern "C" int mvb_s1_begin(mvb_s1* h, int32_t utt, int32_t first_token, int32_t pos, const mvb_sampling* p,
                            const float* d_noise, const int32_t* d_forced, void* stream) {
  if (!h || !p) return fail(MVB_ERR_ARG, "null argument");
  if (utt < 0 || utt >= h->cfg.max_utts) return fail(MVB_ERR_ARG, "utterance slot %d out of range", utt);
  if (pos < 0 || pos >= h->cfg.block_size) return fail(MVB_ERR_ARG, "position %d outside the context", pos);
  h->h_topk[utt] = p->top_k;
  k_begin<<<1, 32, 0, (cudaStream_t)stream>>>(h->st, utt, first_token, pos, to_dev(p), d_noise, d_forced, first_token >= 0,
                                              h->cfg.max_new);
  h->launches++;
  CK(cudaGetLastError());
  return MVB_OK;
}

// per-utterance token budget (generate(): utils:196-204) and the step that row 0 of the staged noise belongs to
__global__ void k_set_budget(S1State st, int utt, int budget) {
  if (threadIdx.x == 0) st.budget[utt] = budget;
}
__global__ void k_set_noise_base(S1State st, int n_utts, int base) {
  if ((int)threadIdx.x < n_utts) st.noise_base[threadIdx.x] = base;
}
__global__ void k_set_done(S1State st, int utt, int done) {
  if (threadIdx.x == 0) st.done[utt] = done;
}

// ------------------------------------------------------------------------------------------------
// Continuous batching: utterances enter and leave KV slots between decode bursts.
// mvb_s1_admit == the per-utterance prologue of generate() (utils:196-212): budget check, prompt / speaker upload,
// prefill, first token sampled from the last prompt position.  The slot then rides along in every mvb_s1_decode burst
// until its done flag latches (mvb_s1_poll), after which the host fetches its tokens and may admit another request.
extern "C" int mvb_s1_admit(mvb_s1* h, int32_t utt, const int32_t* prompt, int32_t T, const float* spk_emb,
                            const mvb_sampling* params, int32_t max_new_tokens, const float* d_noise, void* stream) {
  if (!h || !prompt || !spk_emb || !params) return fail(MVB_ERR_ARG, "null argument");
  const mvb_s1_config& c = h->cfg;
  if (utt < 0 || utt >= c.max_utts) return fail(MVB_ERR_ARG, "utterance slot %d out of range", utt);
  if (T < 1) return fail(MVB_ERR_ARG, "empty prompt");
  if (max_new_tokens < 1 || max_new_tokens > c.max_new) return fail(MVB_ERR_ARG, "max_new_tokens %d out of range", max_new_tokens);
  const int room = (T + max_new_tokens < c.block_size ? T + max_new_tokens : c.block_size) - T;
  if (room <= 0) return fail(MVB_ERR_PROMPT_TOO_LONG, "Prompt is too long to generate more tokens");
  cudaStream_t s = (cudaStream_t)stream;
  int* di = h->wsp<int>(h->L.stage_idx) + (size_t)utt * 2 * c.block_size;
  float* d_spk = h->wsp<float>(h->L.stage_spk) + (size_t)utt * c.spk_dim;
  CK(cudaMemcpyAsync(d_spk, spk_emb, sizeof(float) * c.spk_dim, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(di, prompt, sizeof(int) * T, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(di + T, prompt, sizeof(int) * T, cudaMemcpyHostToDevice, s));
  if (int e = mvb_s1_set_speaker(h, utt, d_spk, s)) return e;
  if (int e = mvb_s1_begin(h, utt, -1, 0, params, d_noise, nullptr, s)) return e;
  k_set_budget<<<1, 32, 0, s>>>(h->st, utt, room);
  h->launches++;
  if (int e = mvb_s1_forward(h, utt, di, T, 0, nullptr, 0, s)) return e;
  SampleP spp{};
  spp.logits = h->wsp<float>(h->L.logits);
  spp.V = c.vocab;
  spp.decode_mode = 1;
  k_sample<<<1, SAMP_THREADS, 0, s>>>(spp, h->st);  // slot_map[0] == utt after the prefill
  h->launches++;
  CK(cudaGetLastError());
  return MVB_OK;
}

// Park a slot: its rows still ride along in the batch (recomputed at a frozen position) but nothing is sampled or
// appended for it until the next mvb_s1_admit / mvb_s1_begin.
extern "C" int mvb_s1_release(mvb_s1* h, int32_t utt, void* stream) {
  if (!h) return fail(MVB_ERR_ARG, "null handle");
  if (utt < 0 || utt >= h->cfg.max_utts) return fail(MVB_ERR_ARG, "utterance slot %d out of range", utt);
  k_set_done<<<1, 32, 0, (cudaStream_t)stream>>>(h->st, utt, 1);
  h->launches++;
  CK(cudaGetLastError());
  return MVB_OK;
}

// done flags and generated-token counts of slots [0, n_slots) in one device->host round trip (synchronises `stream`).
extern "C" int mvb_s1_poll(mvb_s1* h, int32_t n_slots, int32_t* done_out, int32_t* n_gen_out, void* stream) {
  if (!h || !done_out || !n_gen_out) return fail(MVB_ERR_ARG, "null argument");
  if (n_slots < 1 || n_slots > h->cfg.max_utts) return fail(MVB_ERR_ARG, "n_slots %d out of range", n_slots);
  cudaStream_t s = (cudaStream_t)stream;
  CK(cudaMemcpyAsync(h->h_flags, h->st.done, sizeof(int) * n_slots, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(h->h_flags + 64, h->st.n_gen, sizeof(int) * n_slots, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  for (int i = 0; i < n_slots; ++i) { done_out[i] = h->h_flags[i]; n_gen_out[i] = h->h_flags[64 + i]; }
  return MVB_OK;
}

static int sample_step(mvb_s1* h, cudaStream_t s, int n_utts) {
  SampleP sp{};
  sp.logits = h->wsp<float>(h->L.logits);
  sp.V = h->cfg.vocab;
  sp.decode_mode = 1;
  sp.grid_bar = h->wsp<unsigned>(h->L.c_bar);
  h->launches++;
  CK(launch_pdl(h->pdl, k_sample, dim3(n_utts), dim3(SAMP_THREADS), 0, s, sp, h->st));
  return MVB_OK;
}

extern "C" int mvb_s1_decode(mvb_s1* h, int32_t n_utts, int32_t n_steps, void* stream) {
  if (!h) return fail(MVB_ERR_ARG, "null handle");
  if (n_utts < 1 || n_utts > h->cfg.max_utts) return fail(MVB_ERR_ARG, "n_utts %d out of range", n_utts);
  if (n_steps < 1) return MVB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  k_identity_slots<<<1, 64, 0, s>>>(h->st, n_utts);
  h->launches++;
  CK(cudaGetLastError());
  const bool use_c = h->path_c && h->pc_ok && 2 * n_utts <= PC_RPAD;
  if (use_c) {
    // The persistent kernel ACCUMULATES split-K partial logits (red.add); a prefill (mvb_s1_forward) leaves the last
    // position's logits in the same rows, so they must be zero before the first fused step.  Its grid-barrier counter
    // starts from zero at every launch.
    CK(cudaMemsetAsync(h->wsp<float>(h->L.logits), 0, sizeof(float) * 2 * (size_t)n_utts * h->cfg.vocab, s));
    // One launch for the whole burst: the kernel samples on the device (top_k = None only, the stage-1 default of
    // TTS.synthesise) and stays resident across positions.
    bool fused = h->fused && (h->cfg.vocab + h->n_sm - 1) / h->n_sm <= PC_SAMP_OWN && h->cfg.vocab <= SAMP_PAD;
    for (int u = 0; u < n_utts; ++u) fused = fused && h->h_topk[u] <= 0;
    if (fused) return launch_persistent(h, s, n_utts, n_steps, true);
  }
  for (int i = 0; i < n_steps; ++i) {
    if (use_c) {
      if (int e = launch_persistent(h, s, n_utts, 1, false)) return e;   // (the sampler below resets the barrier counter)
    } else {
      if (int e = run_body(h, s, n_utts)) return e;
    }
    if (int e = sample_step(h, s, n_utts)) return e;
  }
  return MVB_OK;
}


// Parity hook for the persistent decode kernel: one fused position for utterances [0, n_utts) from the current
// decode state (tokens/positions installed with mvb_s1_begin), logits copied out, no sampling, state untouched
// except for the KV-cache append at the current positions.
extern "C" int mvb_s1_step_logits(mvb_s1* h, int32_t n_utts, float* d_logits, void* stream) {
  if (!h || !d_logits) return fail(MVB_ERR_ARG, "null argument");
  if (n_utts < 1 || n_utts > h->cfg.max_utts || 2 * n_utts > PC_RPAD) return fail(MVB_ERR_ARG, "n_utts %d out of range", n_utts);
  if (!h->pc_ok) return fail(MVB_ERR_UNSUPPORTED, "persistent decode kernel unavailable for this configuration");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t bytes = sizeof(float) * 2 * (size_t)n_utts * h->cfg.vocab;
  k_identity_slots<<<1, 64, 0, s>>>(h->st, n_utts);
  h->launches++;
  CK(cudaGetLastError());
  CK(cudaMemsetAsync(h->wsp<float>(h->L.logits), 0, bytes, s));
  CK(cudaMemsetAsync(h->wsp<unsigned>(h->L.c_bar), 0, 4, s));
  if (int e = launch_persistent(h, s, n_utts, 1, false)) return e;
  CK(cudaMemcpyAsync(d_logits, h->wsp<float>(h->L.logits), bytes, cudaMemcpyDeviceToDevice, s));
  CK(cudaMemsetAsync(h->wsp<float>(h->L.logits), 0, bytes, s));
  CK(cudaMemsetAsync(h->wsp<unsigned>(h->L.c_bar), 0, 4, s));
  return MVB_OK;
}

// Debug: copy the persistent kernel's per-CTA phase time stamps (clock64) of the last step to the host.
extern "C" int mvb_s1_trace_fetch(mvb_s1* h, long long* out, int32_t max_ctas, void* stream) {
  if (!h || !out) return fail(MVB_ERR_ARG, "null argument");
  const int n = h->n_sm < max_ctas ? h->n_sm : max_ctas;
  CK(cudaMemcpyAsync(out, h->wsp<long long>(h->L.c_trace), sizeof(long long) * (size_t)n * PC_TRACE_EVENTS, cudaMemcpyDeviceToHost,
                     (cudaStream_t)stream));
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  return n;
}

extern "C" int mvb_s1_fetch(mvb_s1* h, int32_t utt, int32_t* out_tokens, int32_t cap, int32_t* n_out, int32_t* done,
                            void* stream) {
  if (!h) return fail(MVB_ERR_ARG, "null handle");
  if (utt < 0 || utt >= h->cfg.max_utts) return fail(MVB_ERR_ARG, "utterance slot %d out of range", utt);
  cudaStream_t s = (cudaStream_t)stream;
  CK(cudaMemcpyAsync(h->h_flags, h->st.n_gen + utt, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(h->h_flags + 1, h->st.done + utt, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  const int n = h->h_flags[0];
  if (n_out) *n_out = n;
  if (done) *done = h->h_flags[1];
  if (out_tokens && cap > 0) {
    const int m = n < cap ? n : cap;
    CK(cudaMemcpyAsync(out_tokens, h->st.gen_tokens + (size_t)utt * h->cfg.max_new, sizeof(int) * m, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  return MVB_OK;
}

// Sampled (pre-teacher-forcing) tokens, for the parity tests.
extern "C" int mvb_s1_fetch_sampled(mvb_s1* h, int32_t utt, int32_t* out_tokens, int32_t cap, void* stream) {
  if (!h || !out_tokens) return fail(MVB_ERR_ARG, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  CK(cudaMemcpyAsync(out_tokens, h->st.sampled_tokens + (size_t)utt * h->cfg.max_new, sizeof(int) * cap, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return MVB_OK;
}

extern "C" int mvb_s1_generate(mvb_s1* h, int32_t n_utts, const int32_t* prompts, const int32_t* prompt_lens,
                               const float* spk_embs, const mvb_sampling* params, int32_t max_new_tokens,
                               const float* noise, int32_t noise_on_device, const int32_t* forced, int32_t* out_tokens,
                               int32_t* out_lens, void* stream) {
  if (!h || !prompts || !prompt_lens || !spk_embs || !params || !out_tokens || !out_lens)
    return fail(MVB_ERR_ARG, "null argument");
  const mvb_s1_config& c = h->cfg;
  if (n_utts < 1 || n_utts > c.max_utts) return fail(MVB_ERR_ARG, "n_utts %d exceeds the %d slots", n_utts, c.max_utts);
  if (max_new_tokens < 1 || max_new_tokens > c.max_new) return fail(MVB_ERR_ARG, "max_new_tokens %d out of range", max_new_tokens);
  cudaStream_t s = (cudaStream_t)stream;
  // generate(): max_seq = min(T + max_new, block_size); raise if no room (utils:196-204)
  std::vector<int> budget(n_utts);
  for (int i = 0; i < n_utts; ++i) {
    const int T = prompt_lens[i];
    if (T < 1) return fail(MVB_ERR_ARG, "empty prompt for utterance %d", i);
    const int room = (T + max_new_tokens < c.block_size ? T + max_new_tokens : c.block_size) - T;
    if (room <= 0) return fail(MVB_ERR_PROMPT_TOO_LONG, "Prompt is too long to generate more tokens");
    budget[i] = room;
  }
  const int V = c.vocab;
  // Exp(1) draws: a device buffer is used in place; a host buffer is staged one burst of decode steps at a time
  // into the workspace (no allocation inside the call).
  const bool stage_noise = noise != nullptr && !noise_on_device;
  float* d_stage = h->wsp<float>(h->L.stage_noise);
  auto stage_rows = [&](int step0, int rows) -> int {
    for (int i = 0; i < n_utts; ++i)
      CK(cudaMemcpyAsync(d_stage + (size_t)i * GEN_BURST * V, noise + ((size_t)i * max_new_tokens + step0) * V,
                         sizeof(float) * (size_t)rows * V, cudaMemcpyHostToDevice, s));
    k_set_noise_base<<<1, 64, 0, s>>>(h->st, n_utts, step0);
    h->launches++;
    CK(cudaGetLastError());
    return MVB_OK;
  };
  int* d_forced = h->wsp<int>(h->L.stage_forced);
  if (forced) CK(cudaMemcpyAsync(d_forced, forced, sizeof(int) * (size_t)n_utts * max_new_tokens, cudaMemcpyHostToDevice, s));
  int* d_idx = h->wsp<int>(h->L.stage_idx);
  float* d_spk = h->wsp<float>(h->L.stage_spk);
  CK(cudaMemcpyAsync(d_spk, spk_embs, sizeof(float) * (size_t)n_utts * c.spk_dim, cudaMemcpyHostToDevice, s));
  int rc = MVB_OK;
  size_t poff = 0;
  int max_budget = 0;
  for (int i = 0; i < n_utts && rc == MVB_OK; ++i) {
    const int T = prompt_lens[i];
    int* di = d_idx + (size_t)i * 2 * c.block_size;
    // prompt.view(1,-1).repeat(2,1): both CFG rows see the same tokens (utils:211)
    CK(cudaMemcpyAsync(di, prompts + poff, sizeof(int) * T, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(di + T, prompts + poff, sizeof(int) * T, cudaMemcpyHostToDevice, s));
    poff += T;
    if ((rc = mvb_s1_set_speaker(h, i, d_spk + (size_t)i * c.spk_dim, s))) break;
    mvb_sampling sp = params[i];
    const float* dn = nullptr;
    if (stage_noise) dn = d_stage + (size_t)i * GEN_BURST * V;
    else if (noise) dn = noise + (size_t)i * max_new_tokens * V;
    if ((rc = mvb_s1_begin(h, i, -1, 0, &sp, dn, forced ? d_forced + (size_t)i * max_new_tokens : nullptr, s))) break;
    k_set_budget<<<1, 32, 0, s>>>(h->st, i, budget[i]);
    h->launches++;
    if (budget[i] > max_budget) max_budget = budget[i];
  }
  if (rc == MVB_OK && stage_noise) rc = stage_rows(0, 1);      // row 0 feeds the post-prefill samples below
  for (int i = 0; i < n_utts && rc == MVB_OK; ++i) {
    const int T = prompt_lens[i];
    int* di = d_idx + (size_t)i * 2 * c.block_size;
    if ((rc = mvb_s1_forward(h, i, di, T, 0, nullptr, 0, s))) break;  // prefill (utils:123-132)
    // first token sampled from the last prefill position (utils:211-212)
    SampleP spp{};
    spp.logits = h->wsp<float>(h->L.logits);
    spp.V = V;
    spp.decode_mode = 1;
    k_sample<<<1, SAMP_THREADS, 0, s>>>(spp, h->st);  // slot_map[0] == i after the prefill
    h->launches++;
    if (cudaGetLastError() != cudaSuccess) { rc = fail(MVB_ERR_CUDA, "sampler launch failed"); break; }
  }
  // decode_n_tokens: at most budget-1 further steps, all utterances advance together with their own positions;
  // per-utterance budgets and the end of the context are latched on the device (done flag).
  if (rc == MVB_OK) {
    int remaining = max_budget - 1, step0 = 1;
    while (remaining > 0 && rc == MVB_OK) {
      const int burst = remaining < GEN_BURST ? remaining : GEN_BURST;
      if (stage_noise && (rc = stage_rows(step0, burst))) break;
      rc = mvb_s1_decode(h, n_utts, burst, s);
      remaining -= burst;
      step0 += burst;
      if (rc) break;
      CK(cudaMemcpyAsync(h->h_flags, h->st.done, sizeof(int) * n_utts, cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
      bool all = true;
      for (int i = 0; i < n_utts; ++i) all = all && h->h_flags[i];
      if (all) break;
    }
  }
  if (rc == MVB_OK) {
    CK(cudaMemcpyAsync(h->h_flags, h->st.n_gen, sizeof(int) * n_utts, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (int i = 0; i < n_utts; ++i) {
      int n = h->h_flags[i];
      if (n > budget[i]) n = budget[i];
      out_lens[i] = n;
      CK(cudaMemcpyAsync(out_tokens + (size_t)i * max_new_tokens, h->st.gen_tokens + (size_t)i * c.max_new, sizeof(int) * n,
                         cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
  }
  return rc;
}
