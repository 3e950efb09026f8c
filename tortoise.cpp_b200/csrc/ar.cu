// ar.cu -- AR stage driver: loader, prefill, KV-cached decode step, latent pass.
// Reference: autoregressive_model_load (main.cpp:482-897), autoregressive_graph
// (main.cpp:2545-3040), autoregressive_latent_graph (main.cpp:2053-2519) and the graph
// launches inside autoregressive() (main.cpp:5186, 5247, 5342).
#include <set>

#include "ar_kernels.cuh"
#include "engine.h"
#include "gemm_launch.cuh"
#include "wsgemv.cuh"
#include "ar_mega2.cuh"
#include "ar_mega3.cuh"
#include "ar_mega4.cuh"

namespace tts {

static size_t wbytes(int dtype) { return dtype == TTS_DTYPE_F16 ? 2 : 4; }

// file tensor [K][N] (GPT-2 Conv1D, "in x out") or [N][K] -> device [N][K] in the AR dtype
static void *upload_matrix(tts_ctx *c, const Container &ct, const std::string &name, int N, int K,
                           bool file_is_kn) {
  auto it = ct.tensors.find(name);
  if (it == ct.tensors.end()) throw ArgError("tensor '" + name + "' missing from " + ct.path, TTS_EIO);
  const auto &ne = it->second.ne;
  const int want0 = file_is_kn ? N : K, want1 = file_is_kn ? K : N;  // ne[0] is the fastest dim
  if (it->second.nelem != size_t(N) * K || ne[0] != want0 || (ne.size() > 1 ? ne[1] : 1) != want1)
    throw ArgError("tensor '" + name + "' has wrong shape in model file", TTS_EIO);
  size_t n = 0;
  read_tensor_to_staging(c, ct, name, &n);
  void *d = nullptr;
  TTS_CUDA_TRY(ctx_malloc(c, &d, n * wbytes(c->cfg.dtype)));
  if (file_is_kn) {
    dim3 grid((N + 31) / 32, (K + 31) / 32), block(32, 8);
    if (c->cfg.dtype == TTS_DTYPE_F16)
      transpose_convert_kernel<__half><<<grid, block, 0, c->stream>>>(c->d_scratch, (__half *)d, K, N);
    else
      transpose_convert_kernel<float><<<grid, block, 0, c->stream>>>(c->d_scratch, (float *)d, K, N);
  } else {
    if (c->cfg.dtype == TTS_DTYPE_F16)
      convert_kernel<__half><<<1024, 256, 0, c->stream>>>(c->d_scratch, (__half *)d, n);
    else
      convert_kernel<float><<<1024, 256, 0, c->stream>>>(c->d_scratch, (float *)d, n);
  }
  TTS_CUDA_TRY(cudaGetLastError());
  TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
  c->ar.decode_weight_bytes += n * wbytes(c->cfg.dtype);
  return d;
}

// hi/lo f16 planes of an f32 [N][K] device matrix (parity mode only)
static void make_planes(tts_ctx *c, const void *w, size_t n, __half **hi, __half **lo) {
  if (c->cfg.dtype == TTS_DTYPE_F16) {
    *hi = (__half *)w;
    *lo = nullptr;
    return;
  }
  TTS_CUDA_TRY(ctx_malloc(c, hi, n * 2));
  TTS_CUDA_TRY(ctx_malloc(c, lo, n * 2));
  split_f16_kernel<<<1024, 256, 0, c->stream>>>((const float *)w, *hi, *lo, n);
  TTS_CUDA_TRY(cudaGetLastError());
  TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
}

void ar_load(tts_ctx *c, const char *path) {
  Container ct;
  std::string err;
  if (!ct.open(path, err)) throw ArgError(err, TTS_EIO);
  ArModel &m = c->ar;
  if (m.loaded) throw ArgError("AR model already loaded in this context");
  m.dtype = c->cfg.dtype;
  m.decode_weight_bytes = 0;
  std::set<std::string> known;
  auto f32 = [&](const std::string &n, std::vector<int> ne) {
    known.insert(n);
    return upload_f32(c, ct, n, ne);
  };
  auto mat = [&](const std::string &n, int N, int K, bool kn) {
    known.insert(n);
    return upload_matrix(c, ct, n, N, K, kn);
  };
  for (int i = 0; i < kLayers; ++i) {
    const std::string p = "inference_model.transformer.h." + std::to_string(i) + ".";
    ArLayer &l = m.layers[i];
    l.ln1_w = f32(p + "ln_1.weight", {1024});
    l.ln1_b = f32(p + "ln_1.bias", {1024});
    l.w_qkv = mat(p + "attn.c_attn.weight", 3072, 1024, true);
    l.b_qkv = f32(p + "attn.c_attn.bias", {3072});
    l.w_proj = mat(p + "attn.c_proj.weight", 1024, 1024, true);
    l.b_proj = f32(p + "attn.c_proj.bias", {1024});
    l.ln2_w = f32(p + "ln_2.weight", {1024});
    l.ln2_b = f32(p + "ln_2.bias", {1024});
    l.w_fc = mat(p + "mlp.c_fc.weight", 4096, 1024, true);
    l.b_fc = f32(p + "mlp.c_fc.bias", {4096});
    l.w_proj2 = mat(p + "mlp.c_proj.weight", 1024, 4096, true);
    l.b_proj2 = f32(p + "mlp.c_proj.bias", {1024});
    make_planes(c, l.w_qkv, size_t(3072) * 1024, &l.qkv_hi, &l.qkv_lo);
    make_planes(c, l.w_proj, size_t(1024) * 1024, &l.proj_hi, &l.proj_lo);
    make_planes(c, l.w_fc, size_t(4096) * 1024, &l.fc_hi, &l.fc_lo);
    make_planes(c, l.w_proj2, size_t(1024) * 4096, &l.proj2_hi, &l.proj2_lo);
  }
  m.lnf_w = f32("inference_model.transformer.ln_f.weight", {1024});
  m.lnf_b = f32("inference_model.transformer.ln_f.bias", {1024});
  m.lm0_w = f32("inference_model.lm_head.0.weight", {1024});
  m.lm0_b = f32("inference_model.lm_head.0.bias", {1024});
  m.lm_w = mat("inference_model.lm_head.1.weight", kMelVocab, 1024, false);
  m.lm_b = f32("inference_model.lm_head.1.bias", {kMelVocab});
  m.text_emb = f32("text_embedding.weight", {1024, 256});
  m.text_pos = f32("text_pos_embedding.emb.weight", {1024, 404});
  m.mel_emb = f32("mel_embedding.weight", {1024, kMelVocab});
  m.mel_pos = f32("mel_pos_embedding.emb.weight", {1024, 608});
  for (const auto &n : ct.order)
    if (!known.count(n)) throw ArgError("unknown tensor '" + n + "' in model file", TTS_EIO);  // main.cpp:834-838

  // decode-state buffers
  ArState &s = c->ars;
  s.Bmax = c->cfg.max_batch;
  s.P = c->cfg.max_positions;
  const size_t B = s.Bmax;
  TTS_CUDA_TRY(ctx_malloc(c, &s.h, B * kDim * 4));
  TTS_CUDA_TRY(ctx_malloc(c, &s.q, B * kDim * 4));
  TTS_CUDA_TRY(ctx_malloc(c, &s.attn, B * kDim * 4));
  TTS_CUDA_TRY(ctx_malloc(c, &s.m, B * kFF * 4));
  TTS_CUDA_TRY(ctx_malloc(c, &s.logits, B * kMelVocab * 4));
  const size_t kv = size_t(kLayers) * B * kHeads * s.P * kHeadDim;
  TTS_CUDA_TRY(ctx_malloc(c, &s.kc, kv * 2));
  TTS_CUDA_TRY(ctx_malloc(c, &s.vc, kv * 2));
  // (utterance batching leaves padding rows in front of a right-aligned prompt: never read, kept finite anyway)
  TTS_CUDA_TRY(cudaMemsetAsync(s.kc, 0, kv * 2, c->stream));
  TTS_CUDA_TRY(cudaMemsetAsync(s.vc, 0, kv * 2, c->stream));
  TTS_CUDA_TRY(ctx_malloc(c, &s.d_tokens, B * 4));
  TTS_CUDA_TRY(ctx_malloc(c, &s.d_state, 16));
  TTS_CUDA_TRY(ctx_malloc_host(c, &s.h_tokens, B * 4));
  TTS_CUDA_TRY(ctx_malloc_host(c, &s.h_state, 16));
  TTS_CUDA_TRY(ctx_malloc_host(c, &s.h_logits, B * kMelVocab * 4));
  TTS_CUDA_TRY(ctx_malloc(c, &s.d_text, 1024 * 4));
  TTS_CUDA_TRY(ctx_malloc(c, &s.d_voice, kDim * 4));
  {
    std::vector<MegaLayer> ml(kLayers);
    for (int i = 0; i < kLayers; ++i) {
      const ArLayer &l = m.layers[i];
      ml[i] = MegaLayer{l.ln1_w, l.ln1_b, l.ln2_w, l.ln2_b, l.w_qkv, l.w_proj, l.w_fc, l.w_proj2,
                        l.b_qkv,  l.b_proj, l.b_fc,  l.b_proj2};
    }
    TTS_CUDA_TRY(ctx_malloc(c, &m.mega_layers, sizeof(MegaLayer) * kLayers));
    TTS_CUDA_TRY(cudaMemcpy(m.mega_layers, ml.data(), sizeof(MegaLayer) * kLayers, cudaMemcpyHostToDevice));
    {  // exchange buffers of ar_mega2.cuh: zero tags never match (launch generations start at 1)
      const size_t nb = std::min<size_t>(B, 4);
      const size_t sizes[5] = {M2_REP * nb * kDim, M2_REP * nb * kDim, nb * 3072, M2_REP * nb * (kFF / 2),
                               M2_REP * nb * kHeads * M2_SMAX * M2_REC};
      uint2 **ptrs[5] = {&m.ll_h, &m.ll_h2, &m.ll_qkv, &m.ll_m, &m.ll_att};
      for (int i = 0; i < 5; ++i) {
        m.ll_bytes[i] = sizes[i] * sizeof(uint2);
        TTS_CUDA_TRY(ctx_malloc(c, ptrs[i], m.ll_bytes[i]));
        TTS_CUDA_TRY(cudaMemset(*ptrs[i], 0, m.ll_bytes[i]));
      }
      m.mega_epoch = 0;
    }
    if (B > 4 && m.dtype == TTS_DTYPE_F16) {  // 5..16 candidates on one weight stream (ar_mega4.cuh)
      const size_t sizes[5] = {size_t(2) * 16 * kDim, size_t(2) * 16 * kDim, size_t(16) * 3072, size_t(2) * 16 * (kFF / 2),
                               size_t(2) * 16 * kHeads * M4_REC};
      uint2 **ptrs[5] = {&m.l4_h, &m.l4_h2, &m.l4_qkv, &m.l4_m, &m.l4_att};
      for (int i = 0; i < 5; ++i) {
        m.l4_bytes[i] = sizes[i] * sizeof(uint2);
        TTS_CUDA_TRY(ctx_malloc(c, ptrs[i], m.l4_bytes[i]));
        TTS_CUDA_TRY(cudaMemset(*ptrs[i], 0, m.l4_bytes[i]));
      }
    }
    TTS_CUDA_TRY(ctx_malloc(c, &s.d_topv, B * AR_TOPK * 4));
    TTS_CUDA_TRY(ctx_malloc(c, &s.d_topi, B * AR_TOPK * 4));
    TTS_CUDA_TRY(ctx_malloc(c, &s.d_topf, B * 4));
    TTS_CUDA_TRY(ctx_malloc_host(c, &s.h_topv, B * AR_TOPK * 4));
    TTS_CUDA_TRY(ctx_malloc_host(c, &s.h_topi, B * AR_TOPK * 4));
    TTS_CUDA_TRY(ctx_malloc_host(c, &s.h_topf, B * 4));
    const char *tr = getenv("TTS_MEGA_TRACE");
    if (tr && (tr[0] == '1' || tr[0] == '2')) {
      m.mega_dbg_mode = tr[0] - '0';
      TTS_CUDA_TRY(ctx_malloc(c, &m.mega_dbg, kMegaDbgWords * sizeof(long long)));
      TTS_CUDA_TRY(cudaMemset(m.mega_dbg, 0, kMegaDbgWords * sizeof(long long)));
    }
  }
  m.loaded = true;
}

void ar_free(tts_ctx *c) {
  ArState &s = c->ars;
  if (s.step_graph) cudaGraphExecDestroy(s.step_graph);
  s.step_graph = nullptr;
  // (device / pinned buffers belong to the context's allocation registry: tts_free -> ctx_free_all)
}

// ------------------------------------------------------------------------------------------
template <typename WT>
static void launch_gemv_t(tts_ctx *c, const Launcher &L, const GemvArgs &a) {
  const size_t smem = gemv_smem_bytes();
  auto k1 = wsgemv_kernel<WT, 1>;
  auto k2 = wsgemv_kernel<WT, 2>;
  ensure_smem_attr(k1, smem);
  ensure_smem_attr(k2, smem);
  const int G = c->num_sms;
  if ((a.N + G - 1) / G > GV_MAX_ROWS_PER_CTA) throw ArgError("gemv: too many rows per CTA");
  if (a.B == 1)
    L(k1, dim3(G), dim3(GV_THREADS), smem, a);
  else
    L(k2, dim3(G), dim3(GV_THREADS), smem, a);
}
static void launch_gemv(tts_ctx *c, const Launcher &L, const GemvArgs &a) {
  if (c->ar.dtype == TTS_DTYPE_F16) launch_gemv_t<__half>(c, L, a);
  else launch_gemv_t<float>(c, L, a);
}

static GemvArgs gemv_args(const void *W, const float *bias, const float *in, float *out, int N, int K, int B,
                          int pro, int epi) {
  GemvArgs a{};
  a.W = W; a.bias = bias; a.in = in; a.out = out; a.N = N; a.K = K; a.B = B; a.pro = pro; a.epi = epi;
  return a;
}

static void launch_tgemm(tts_ctx *c, const Launcher &L, const __half *Ahi, const __half *Alo, const __half *Whi,
                         const __half *Wlo, const float *bias, float *C, __half *Chi, __half *Clo, int M, int N,
                         int K, int lda, int ldc, int ldh, int epi) {
  TGemmArgs g{Ahi, Alo, Whi, Wlo, bias, C, Chi, Clo, M, N, K, lda, ldc, ldh, epi, 1, 1, 0, 0, M};
  launch_gemm(L, g);
}

// The lm-head on decode-shaped activations: logits = W . LN(LN(h) wf + bf) w0 + b0 (A-1)
static void enqueue_lm_head(tts_ctx *c, const Launcher &L, int B) {
  ArModel &m = c->ar;
  ArState &s = c->ars;
  GemvArgs a = gemv_args(m.lm_w, m.lm_b, s.h, s.logits, kMelVocab, kDim, B, PRO_LN2, EPI_STORE);
  a.ln_w = m.lnf_w; a.ln_b = m.lnf_b; a.ln2_w = m.lm0_w; a.ln2_b = m.lm0_b;
  launch_gemv(c, L, a);
}

// All launches of one decode step (after tokens/state are on the device).
static void enqueue_step(tts_ctx *c, const Launcher &L, int B) {
  ArModel &m = c->ar;
  ArState &s = c->ars;
  L(ar_embed_decode_kernel, dim3(B), dim3(256), 0, (const int *)s.d_tokens, (const int *)s.d_state,
    (const float *)m.mel_emb, (const float *)m.mel_pos, s.h);
  const int kvb = kHeads * s.P * kHeadDim;
  const size_t layer_kv = size_t(s.Bmax) * kvb;
  const size_t attn_smem = size_t(s.P + 128) * sizeof(float);
  for (int i = 0; i < kLayers; ++i) {
    ArLayer &l = m.layers[i];
    GemvArgs a = gemv_args(l.w_qkv, l.b_qkv, s.h, s.q, 3072, kDim, B, PRO_LN, EPI_QKV);
    a.ln_w = l.ln1_w; a.ln_b = l.ln1_b;
    a.kcache = s.kc + i * layer_kv; a.vcache = s.vc + i * layer_kv;
    a.state = s.d_state; a.kv_b_stride = kvb;
    launch_gemv(c, L, a);
    L(ar_attn_decode_kernel, dim3(kHeads, B), dim3(128), attn_smem, (const float *)s.q,
      (const __half *)(s.kc + i * layer_kv), (const __half *)(s.vc + i * layer_kv), s.attn,
      (const int *)s.d_state, s.P);
    launch_gemv(c, L, gemv_args(l.w_proj, l.b_proj, s.attn, s.h, kDim, kDim, B, PRO_NONE, EPI_RESID));
    GemvArgs f = gemv_args(l.w_fc, l.b_fc, s.h, s.m, kFF, kDim, B, PRO_LN, EPI_GELU16);
    f.ln_w = l.ln2_w; f.ln_b = l.ln2_b;
    launch_gemv(c, L, f);
    launch_gemv(c, L, gemv_args(l.w_proj2, l.b_proj2, s.m, s.h, kDim, kFF, B, PRO_NONE, EPI_RESID));
  }
  enqueue_lm_head(c, L, B);
}

// Second generation (ar_mega2.cuh): tag-fused activation exchange, up to 4 candidates per weight stream.
template <typename WT, int BT>
static void launch_mega2_bt(tts_ctx *c, Mega2Args &a) {
  auto k = ar_decode_mega2_kernel<WT, BT>;
  const size_t smem = mega2_smem_bytes<BT>();
  ensure_smem_attr(k, smem);
  void *args[] = {&a};
  TTS_CUDA_TRY(cudaLaunchCooperativeKernel((void *)k, dim3(c->num_sms), dim3(M2_THREADS), args, smem, c->stream));
  c->launches += 1;
}

// Third generation (ar_mega3.cuh): same protocol, GEMV phases on the tensor cores (f16 weights only).
template <int BT>
static void launch_mega3_bt(tts_ctx *c, Mega2Args &a) {
  auto k = ar_decode_mega3_kernel<BT>;
  const size_t smem = mega3_smem_bytes<BT>();
  ensure_smem_attr(k, smem);
  void *args[] = {&a};
  TTS_CUDA_TRY(cudaLaunchCooperativeKernel((void *)k, dim3(c->num_sms), dim3(M2_THREADS), args, smem, c->stream));
  c->launches += 1;
}

template <typename WT>
static void launch_mega2_t(tts_ctx *c, int B, int n_past, int pos_id) {
  ArModel &m = c->ar;
  ArState &s = c->ars;
  if (m.mega_epoch >= (1u << 24) - 64) {  // tag generations exhausted: start over with clean buffers
    uint2 *ptrs[5] = {m.ll_h, m.ll_h2, m.ll_qkv, m.ll_m, m.ll_att};
    for (int i = 0; i < 5; ++i) TTS_CUDA_TRY(cudaMemsetAsync(ptrs[i], 0, m.ll_bytes[i], c->stream));
    m.mega_epoch = 0;
  }
  Mega2Args a{};
  a.layers = (const MegaLayer *)m.mega_layers;
  a.lnf_w = m.lnf_w; a.lnf_b = m.lnf_b; a.lm0_w = m.lm0_w; a.lm0_b = m.lm0_b; a.lm_b = m.lm_b; a.lm_w = m.lm_w;
  a.mel_emb = m.mel_emb; a.mel_pos = m.mel_pos; a.tokens = s.d_tokens;
  a.ll_h = m.ll_h; a.ll_h2 = m.ll_h2; a.ll_qkv = m.ll_qkv; a.ll_m = m.ll_m; a.ll_att = m.ll_att;
  a.logits = s.logits; a.kc = s.kc; a.vc = s.vc;
  a.Bmax = s.Bmax; a.P = s.P; a.n_past = n_past; a.pos_id = pos_id;
  a.dbg = m.mega_dbg;
  a.dbg_mode = m.mega_dbg_mode;
  // settled by same-box A/B runs (profiles/r01_decode_*_ab.txt): 2 replicas of the all-to-all vectors,
  // deferred ring-slot release, plain re-polls, one attention item per head, no L2 eviction hint
  a.nrep = 2;
  a.defer = 1;
  a.poll_spin = 0;
  a.keys_per_split = 128;
  a.evict_first = 0;
  // up to 4 candidates ride on one weight stream; more candidates = one launch per group of 4
  // (the exchange buffers are reused: tags are unique per launch)
  for (int b0 = 0; b0 < B; b0 += 4) {
    const int Bg = std::min(4, B - b0);
    a.b0 = b0;
    a.B = Bg;
    a.epoch = ++m.mega_epoch;
    if (sizeof(WT) == 2 && !c->use_mega_v2) {
      if (Bg == 1) launch_mega3_bt<1>(c, a);
      else if (Bg == 2) launch_mega3_bt<2>(c, a);
      else launch_mega3_bt<4>(c, a);
    } else {
      if (Bg == 1) launch_mega2_bt<WT, 1>(c, a);
      else if (Bg == 2) launch_mega2_bt<WT, 2>(c, a);
      else launch_mega2_bt<WT, 4>(c, a);
    }
  }
}

// 5..16 candidates: ONE launch, one weight stream (ar_mega4.cuh; f16 weights)
template <int BT>
static void launch_mega4_bt(tts_ctx *c, Mega4Args &a) {
  auto k = ar_decode_mega4_kernel<BT>;
  const size_t smem = mega4_smem_bytes<BT>();
  ensure_smem_attr(k, smem);
  void *args[] = {&a};
  TTS_CUDA_TRY(cudaLaunchCooperativeKernel((void *)k, dim3(c->num_sms), dim3(M2_THREADS), args, smem, c->stream));
  c->launches += 1;
}
static void launch_mega4(tts_ctx *c, int B, int n_past, int pos_id) {
  ArModel &m = c->ar;
  ArState &s = c->ars;
  if (m.mega_epoch >= (1u << 24) - 64) {
    uint2 *ptrs[5] = {m.l4_h, m.l4_h2, m.l4_qkv, m.l4_m, m.l4_att};
    for (int i = 0; i < 5; ++i) TTS_CUDA_TRY(cudaMemsetAsync(ptrs[i], 0, m.l4_bytes[i], c->stream));
    uint2 *p2[5] = {m.ll_h, m.ll_h2, m.ll_qkv, m.ll_m, m.ll_att};
    for (int i = 0; i < 5; ++i) TTS_CUDA_TRY(cudaMemsetAsync(p2[i], 0, m.ll_bytes[i], c->stream));
    m.mega_epoch = 0;
  }
  Mega4Args a{};
  a.layers = (const MegaLayer *)m.mega_layers;
  a.lnf_w = m.lnf_w; a.lnf_b = m.lnf_b; a.lm0_w = m.lm0_w; a.lm0_b = m.lm0_b; a.lm_b = m.lm_b; a.lm_w = m.lm_w;
  a.mel_emb = m.mel_emb; a.mel_pos = m.mel_pos; a.tokens = s.d_tokens;
  a.ll_h = m.l4_h; a.ll_h2 = m.l4_h2; a.ll_qkv = m.l4_qkv; a.ll_m = m.l4_m; a.ll_att = m.l4_att;
  a.logits = s.logits; a.kc = s.kc; a.vc = s.vc;
  a.B = B; a.Bmax = s.Bmax; a.P = s.P; a.n_past = n_past; a.pos_id = pos_id;
  a.n_prefix = s.n_prefix;
  if (s.multi)
    for (int b = 0; b < 16; ++b) a.start[b] = s.start[b];
  a.epoch = ++m.mega_epoch;
  a.nrep = 2;
  if (B <= 8) launch_mega4_bt<8>(c, a);
  else launch_mega4_bt<16>(c, a);
}

// which decode path serves B candidates: the shared-prefix single-launch kernel needs f16 weights
static bool use_mega4(const tts_ctx *c, int B) {
  return c->use_mega && c->ar.dtype == TTS_DTYPE_F16 && B > 4 && B <= 16 && c->ars.P <= 1024 && c->ar.l4_h != nullptr;
}

static void ensure_rows(tts_ctx *c, size_t rows) {
  ArState &s = c->ars;
  if (rows <= s.rows_cap) return;
  for (void *p : {(void *)s.H, (void *)s.QKV, (void *)s.Z, (void *)s.Ahi, (void *)s.Alo, (void *)s.ATThi,
                  (void *)s.ATTlo, (void *)s.Mhi})
    if (p) ctx_free(c, p);
  TTS_CUDA_TRY(ctx_malloc(c, &s.H, rows * kDim * 4));
  TTS_CUDA_TRY(ctx_malloc(c, &s.QKV, rows * 3072 * 4));
  TTS_CUDA_TRY(ctx_malloc(c, &s.Z, rows * kDim * 4));
  TTS_CUDA_TRY(ctx_malloc(c, &s.Ahi, rows * kDim * 2));
  TTS_CUDA_TRY(ctx_malloc(c, &s.Alo, rows * kDim * 2));
  TTS_CUDA_TRY(ctx_malloc(c, &s.ATThi, rows * kDim * 2));
  TTS_CUDA_TRY(ctx_malloc(c, &s.ATTlo, rows * kDim * 2));
  TTS_CUDA_TRY(ctx_malloc(c, &s.Mhi, rows * kFF * 2));
  s.rows_cap = rows;
}

// 30 transformer layers over nb sequences of R rows each (H in/out).  If kv_B > 0 the
// K/V rows (single sequence) are also scattered into the decode cache of kv_B candidates.
static void enqueue_rows_layers(tts_ctx *c, const Launcher &L, int nb, int R, int kv_B, int kv_slot0 = 0, int kv_pos_off = 0) {
  ArModel &m = c->ar;
  ArState &s = c->ars;
  const int rows = nb * R;
  const size_t layer_kv = size_t(s.Bmax) * kHeads * s.P * kHeadDim;
  for (int i = 0; i < kLayers; ++i) {
    ArLayer &l = m.layers[i];
    L(ln_rows_kernel, dim3(rows), dim3(256), 0, (const float *)s.H, (float *)nullptr, s.Ahi, s.Alo,
      (const float *)l.ln1_w, (const float *)l.ln1_b, (const float *)nullptr, (const float *)nullptr, kDim, kDim);
    launch_tgemm(c, L, s.Ahi, s.Alo, l.qkv_hi, l.qkv_lo, l.b_qkv, s.QKV, nullptr, nullptr, rows, 3072, kDim, kDim,
                 3072, 0, E_BIAS_H16);
    if (kv_B > 0)
      L(ar_kv_scatter_kernel, dim3(R), dim3(256), 0, (const float *)s.QKV, s.kc + i * layer_kv,
        s.vc + i * layer_kv, R, kv_B, s.P, kv_slot0, kv_pos_off);
    L(ar_attn_causal_kernel, dim3((R + 15) / 16, kHeads, nb), dim3(128), 0, (const float *)s.QKV, s.ATThi,
      s.ATTlo, R);
    launch_tgemm(c, L, s.ATThi, s.ATTlo, l.proj_hi, l.proj_lo, l.b_proj, s.H, nullptr, nullptr, rows, kDim, kDim,
                 kDim, kDim, 0, E_BIAS_RESID);
    L(ln_rows_kernel, dim3(rows), dim3(256), 0, (const float *)s.H, (float *)nullptr, s.Ahi, s.Alo,
      (const float *)l.ln2_w, (const float *)l.ln2_b, (const float *)nullptr, (const float *)nullptr, kDim, kDim);
    // gelu16 outputs are f16-exact: a single plane carries them without loss
    launch_tgemm(c, L, s.Ahi, s.Alo, l.fc_hi, l.fc_lo, l.b_fc, nullptr, s.Mhi, nullptr, rows, kFF, kDim, kDim, 0,
                 kFF, E_BIAS_GELU16);
    launch_tgemm(c, L, s.Mhi, nullptr, l.proj2_hi, l.proj2_lo, l.b_proj2, s.H, nullptr, nullptr, rows, kDim, kFF,
                 kFF, kDim, 0, E_BIAS_RESID);
  }
}

void ar_prefill(tts_ctx *c, const int32_t *text, int T, const float *voice, int B, float *logits_out) {
  ArModel &m = c->ar;
  ArState &s = c->ars;
  if (!m.loaded) throw ArgError("AR model not loaded");
  if (B < 1 || B > s.Bmax) throw ArgError("batch exceeds max_batch", TTS_ELIMIT);
  if (T < 1 || T > 404) throw ArgError("text longer than the 404 text positions (main.cpp:685-689)", TTS_ELIMIT);
  const int R = T + 2;
  if (R >= s.P) throw ArgError("prompt does not fit max_positions", TTS_ELIMIT);
  for (int j = 0; j < T; ++j)
    if (text[j] < 0 || text[j] > 255) throw ArgError("text token out of range");
  ensure_rows(c, R);
  Launcher L{c->stream, c->use_pdl, &c->launches};
  TTS_CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
  TTS_CUDA_TRY(cudaMemcpyAsync(s.d_text, text, T * 4, cudaMemcpyHostToDevice, c->stream));
  TTS_CUDA_TRY(cudaMemcpyAsync(s.d_voice, voice, kDim * 4, cudaMemcpyHostToDevice, c->stream));
  // mel part of the prefill row set: one start token (8192) at mel position 0 (A-4)
  s.h_tokens[0] = TTS_MEL_START;
  s.h_state[0] = 0;
  TTS_CUDA_TRY(cudaMemcpyAsync(s.d_tokens, s.h_tokens, 4, cudaMemcpyHostToDevice, c->stream));
  TTS_CUDA_TRY(cudaMemcpyAsync(s.d_state, s.h_state, 4, cudaMemcpyHostToDevice, c->stream));
  // (d_state[0] doubles as the all-zero mel position table for the single start token)
  L(ar_embed_rows_kernel, dim3(R, 1), dim3(256), 0, (const int *)s.d_text, T, (const float *)s.d_voice,
    (const int *)s.d_tokens, (const int *)s.d_state, 1, (const float *)m.text_emb, (const float *)m.text_pos,
    (const float *)m.mel_emb, (const float *)m.mel_pos, s.H);
  // 5..16 candidates on the single-launch path: the prompt's K/V rows are stored once (slot 0)
  const bool shared = use_mega4(c, B);
  enqueue_rows_layers(c, L, 1, R, shared ? 1 : B);
  L(bcast_row_kernel, dim3(B), dim3(256), 0, (const float *)(s.H + size_t(R - 1) * kDim), s.h, kDim);
  enqueue_lm_head(c, L, B);
  if (logits_out)
    TTS_CUDA_TRY(cudaMemcpyAsync(s.h_logits, s.logits, size_t(B) * kMelVocab * 4, cudaMemcpyDeviceToHost, c->stream));
  TTS_CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
  TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
  TTS_CUDA_TRY(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  c->total_ms += c->last_ms;
  if (logits_out) memcpy(logits_out, s.h_logits, size_t(B) * kMelVocab * 4);
  s.B = B;
  s.T = T;
  s.n_past = R;
  s.n_prefix = shared ? R : 0;
  s.multi = false;
}

// Utterance batching of the decode loop (BASELINE configs[4]): U DIFFERENT prompts occupy the candidate slots of one
// batched decode launch (ar_mega4.cuh).  Each prompt is prefilled on its own (the reference's prefill graph, one
// sequence); its K/V rows go to slot u RIGHT-ALIGNED at rows [Rmax - R_u, Rmax), so that every slot appends its
// generated rows at the same index and only the attention's first key differs per slot (Mega4Args::start).
// logits_out [U][8194]: each prompt's first-step logits.  The reference has no counterpart (one prompt per run).
void ar_prefill_multi(tts_ctx *c, int U, const int32_t *const *text, const int32_t *T, const float *voice, float *logits_out) {
  ArModel &m = c->ar;
  ArState &s = c->ars;
  if (!m.loaded) throw ArgError("AR model not loaded");
  if (U < 1 || U > s.Bmax || U > 16) throw ArgError("utterance batch exceeds max_batch (<= 16)", TTS_ELIMIT);
  if (!(c->use_mega && m.dtype == TTS_DTYPE_F16 && s.P <= 1024 && m.l4_h != nullptr))
    throw ArgError("utterance-batched decode needs the f16 batched decode kernel (dtype f16, max_batch >= 5, max_positions <= 1024)");
  int Rmax = 0;
  for (int u = 0; u < U; ++u) {
    if (T[u] < 1 || T[u] > 404) throw ArgError("text longer than the 404 text positions (main.cpp:685-689)", TTS_ELIMIT);
    for (int j = 0; j < T[u]; ++j)
      if (text[u][j] < 0 || text[u][j] > 255) throw ArgError("text token out of range");
    Rmax = std::max(Rmax, T[u] + 2);
  }
  if (Rmax >= s.P) throw ArgError("prompt does not fit max_positions", TTS_ELIMIT);
  ensure_rows(c, Rmax);
  Launcher L{c->stream, c->use_pdl, &c->launches};
  TTS_CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
  TTS_CUDA_TRY(cudaMemcpyAsync(s.d_voice, voice, kDim * 4, cudaMemcpyHostToDevice, c->stream));
  s.h_tokens[0] = TTS_MEL_START;
  s.h_state[0] = 0;
  TTS_CUDA_TRY(cudaMemcpyAsync(s.d_tokens, s.h_tokens, 4, cudaMemcpyHostToDevice, c->stream));
  TTS_CUDA_TRY(cudaMemcpyAsync(s.d_state, s.h_state, 4, cudaMemcpyHostToDevice, c->stream));
  for (int u = 0; u < U; ++u) {
    const int R = T[u] + 2;
    TTS_CUDA_TRY(cudaMemcpyAsync(s.d_text, text[u], T[u] * 4, cudaMemcpyHostToDevice, c->stream));
    L(ar_embed_rows_kernel, dim3(R, 1), dim3(256), 0, (const int *)s.d_text, T[u], (const float *)s.d_voice,
      (const int *)s.d_tokens, (const int *)s.d_state, 1, (const float *)m.text_emb, (const float *)m.text_pos,
      (const float *)m.mel_emb, (const float *)m.mel_pos, s.H);
    enqueue_rows_layers(c, L, 1, R, 1, u, Rmax - R);
    L(bcast_row_kernel, dim3(1), dim3(256), 0, (const float *)(s.H + size_t(R - 1) * kDim), s.h + size_t(u) * kDim, kDim);
    s.start[u] = Rmax - R;
  }
  for (int u = U; u < 16; ++u) s.start[u] = 0;
  enqueue_lm_head(c, L, U);
  if (logits_out)
    TTS_CUDA_TRY(cudaMemcpyAsync(s.h_logits, s.logits, size_t(U) * kMelVocab * 4, cudaMemcpyDeviceToHost, c->stream));
  TTS_CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
  TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
  TTS_CUDA_TRY(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  c->total_ms += c->last_ms;
  if (logits_out) memcpy(logits_out, s.h_logits, size_t(U) * kMelVocab * 4);
  s.B = U;
  s.T = Rmax - 2;
  s.n_past = Rmax;
  s.n_prefix = 0;
  s.multi = true;
}

static void build_step_graph(tts_ctx *c, int B) {
  ArState &s = c->ars;
  if (s.step_graph) {
    cudaGraphExecDestroy(s.step_graph);
    s.step_graph = nullptr;
  }
  cudaGraph_t graph;
  int64_t dummy = 0;
  Launcher L{c->stream, c->use_pdl, &dummy};
  TTS_CUDA_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
  try {
    TTS_CUDA_TRY(cudaMemcpyAsync(s.d_tokens, s.h_tokens, B * 4, cudaMemcpyHostToDevice, c->stream));
    TTS_CUDA_TRY(cudaMemcpyAsync(s.d_state, s.h_state, 8, cudaMemcpyHostToDevice, c->stream));
    enqueue_step(c, L, B);
    TTS_CUDA_TRY(cudaMemcpyAsync(s.h_logits, s.logits, size_t(B) * kMelVocab * 4, cudaMemcpyDeviceToHost, c->stream));
  } catch (...) {
    cudaGraph_t g2;
    cudaStreamEndCapture(c->stream, &g2);
    throw;
  }
  TTS_CUDA_TRY(cudaStreamEndCapture(c->stream, &graph));
  {  // kernel nodes of the captured step: what one replay launches (the launch counter is a count)
    size_t n = 0;
    TTS_CUDA_TRY(cudaGraphGetNodes(graph, nullptr, &n));
    std::vector<cudaGraphNode_t> nodes(n);
    if (n) TTS_CUDA_TRY(cudaGraphGetNodes(graph, nodes.data(), &n));
    s.step_graph_kernels = 0;
    for (size_t i = 0; i < n; ++i) {
      cudaGraphNodeType t;
      TTS_CUDA_TRY(cudaGraphNodeGetType(nodes[i], &t));
      if (t == cudaGraphNodeTypeKernel) ++s.step_graph_kernels;
    }
  }
  TTS_CUDA_TRY(cudaGraphInstantiate(&s.step_graph, graph, 0));
  cudaGraphDestroy(graph);
  s.step_graph_B = B;
}

void ar_step(tts_ctx *c, const int32_t *tokens, int pos_id, float *logits_out, bool sync_out) {
  ArState &s = c->ars;
  if (!c->ar.loaded || s.B == 0) throw ArgError("tts_ar_step before tts_ar_prefill");
  if (s.n_past + 1 > s.P) throw ArgError("KV cache full (reference limit: 404 slots, main.cpp:794-797)", TTS_ELIMIT);
  if (pos_id < 0 || pos_id >= 608) throw ArgError("mel position id out of range (608 rows)", TTS_ELIMIT);
  const int B = s.B;
  for (int b = 0; b < B; ++b)
    if (tokens[b] < 0 || tokens[b] >= kMelVocab) throw ArgError("mel token out of range");
  // the pinned token / state words are read by copies queued on the stream (captured in the per-op
  // graph): a back-to-back asynchronous step must not overwrite them before the previous one ran
  const bool mega = c->use_mega && s.P <= 1024;
  if ((s.n_prefix > 0 || s.multi) && !mega) throw ArgError("internal: shared-prefix / multi-prompt KV needs the persistent decode kernel");
  if (!sync_out && !mega) TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
  for (int b = 0; b < B; ++b) s.h_tokens[b] = tokens[b];
  s.h_state[0] = s.n_past;
  s.h_state[1] = pos_id;
  TTS_CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
  if (mega) {
    {  // tokens travel as kernel parameters (no pinned staging to race with an asynchronous caller)
      TokenArgs ta{};
      for (int b = 0; b < B; ++b) ta.tok[b] = tokens[b];
      set_tokens_kernel<<<1, 64, 0, c->stream>>>(s.d_tokens, ta, B);
      TTS_CUDA_TRY(cudaGetLastError());
      c->launches += 1;
    }
    if (s.n_prefix > 0 || s.multi) launch_mega4(c, B, s.n_past, pos_id);
    else if (c->ar.dtype == TTS_DTYPE_F16) launch_mega2_t<__half>(c, B, s.n_past, pos_id);
    else launch_mega2_t<float>(c, B, s.n_past, pos_id);
    if (logits_out)
      TTS_CUDA_TRY(cudaMemcpyAsync(s.h_logits, s.logits, size_t(B) * kMelVocab * 4, cudaMemcpyDeviceToHost, c->stream));
  } else if (c->use_graph) {
    if (!s.step_graph || s.step_graph_B != B) build_step_graph(c, B);
    TTS_CUDA_TRY(cudaGraphLaunch(s.step_graph, c->stream));
    c->launches += s.step_graph_kernels;
  } else {
    Launcher L{c->stream, c->use_pdl, &c->launches};
    TTS_CUDA_TRY(cudaMemcpyAsync(s.d_tokens, s.h_tokens, B * 4, cudaMemcpyHostToDevice, c->stream));
    TTS_CUDA_TRY(cudaMemcpyAsync(s.d_state, s.h_state, 8, cudaMemcpyHostToDevice, c->stream));
    enqueue_step(c, L, B);
    TTS_CUDA_TRY(cudaMemcpyAsync(s.h_logits, s.logits, size_t(B) * kMelVocab * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  TTS_CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
  s.n_past += 1;
  if (sync_out) {
    TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
    TTS_CUDA_TRY(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  c->total_ms += c->last_ms;
    if (logits_out) memcpy(logits_out, s.h_logits, size_t(B) * kMelVocab * 4);
    if (c->ar.mega_dbg && getenv("TTS_MEGA_TRACE_DUMP")) {
      std::vector<long long> t(kMegaDbgWords);
      cudaMemcpy(t.data(), c->ar.mega_dbg, kMegaDbgWords * sizeof(long long), cudaMemcpyDeviceToHost);
      if (c->ar.mega_dbg_mode == 2) {  // raw [cta][128][4] globaltimer stamps -> file named by the variable
        if (FILE *f = fopen(getenv("TTS_MEGA_TRACE_DUMP"), "wb")) {
          fwrite(t.data(), sizeof(long long), size_t(c->num_sms) * 128 * 4, f);
          fclose(f);
        }
      } else {
        for (int i = 1; i < 4000 && t[2 * i] != 0; ++i)
          fprintf(stderr, "trace %3d tag %2lld dt %6lld\n", i, t[2 * i], t[2 * i + 1] - t[2 * i - 1]);
      }
    }
  }
}

// One decode step whose result crosses PCIe as AR_TOPK (value, index) pairs per candidate instead of
// 8194 floats (reference D2H: main.cpp:4767).  flags_out[b] != 0: ties overflowed, fetch the row.
void ar_step_topk(tts_ctx *c, const int32_t *tokens, int pos_id, float *vals_out, int32_t *idx_out, int32_t *flags_out) {
  ArState &s = c->ars;
  ar_step(c, tokens, pos_id, nullptr, false);
  const int B = s.B;
  ar_topk_kernel<<<B, 256, 0, c->stream>>>(s.logits, s.d_topv, s.d_topi, s.d_topf);
  TTS_CUDA_TRY(cudaGetLastError());
  c->launches += 1;
  TTS_CUDA_TRY(cudaMemcpyAsync(s.h_topv, s.d_topv, size_t(B) * AR_TOPK * 4, cudaMemcpyDeviceToHost, c->stream));
  TTS_CUDA_TRY(cudaMemcpyAsync(s.h_topi, s.d_topi, size_t(B) * AR_TOPK * 4, cudaMemcpyDeviceToHost, c->stream));
  TTS_CUDA_TRY(cudaMemcpyAsync(s.h_topf, s.d_topf, size_t(B) * 4, cudaMemcpyDeviceToHost, c->stream));
  TTS_CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
  TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
  TTS_CUDA_TRY(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  c->total_ms += c->last_ms;
  memcpy(vals_out, s.h_topv, size_t(B) * AR_TOPK * 4);
  memcpy(idx_out, s.h_topi, size_t(B) * AR_TOPK * 4);
  memcpy(flags_out, s.h_topf, size_t(B) * 4);
}

// full logits [B][8194] of the last prefill / step (fallback of the top-k path, tests)
void ar_logits(tts_ctx *c, float *logits_out) {
  ArState &s = c->ars;
  if (!c->ar.loaded || s.B == 0) throw ArgError("tts_ar_logits before tts_ar_prefill");
  TTS_CUDA_TRY(cudaMemcpyAsync(s.h_logits, s.logits, size_t(s.B) * kMelVocab * 4, cudaMemcpyDeviceToHost, c->stream));
  TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
  memcpy(logits_out, s.h_logits, size_t(s.B) * kMelVocab * 4);
}

// Latent pass.  Reference quirk A-4: the mel position table is only written for
// index < 502*B/4 per candidate slot (main.cpp:5326-5333); the rest stays zero.
void ar_latents(tts_ctx *c, const int32_t *text, int T, const float *voice, const int32_t *codes, int B,
                int n_keep, float *out) {
  ArModel &m = c->ar;
  ArState &s = c->ars;
  if (!m.loaded) throw ArgError("AR model not loaded");
  if (B < 1 || B > s.Bmax) throw ArgError("batch exceeds max_batch", TTS_ELIMIT);
  if (n_keep < 1 || n_keep > 500) throw ArgError("n_keep must be in [1,500]");
  if (T < 1 || T > 404) throw ArgError("text too long", TTS_ELIMIT);
  const int n_mel = n_keep;  // causal: mel rows beyond n_keep-1 cannot influence kept latents
  const int R = 1 + T + n_mel;
  // process candidates in chunks so the row buffers stay bounded
  const int chunk = std::max(1, std::min(B, 8192 / R));
  ensure_rows(c, size_t(chunk) * R);
  std::vector<int> pos(size_t(B) * 502, 0), h_codes(size_t(B) * n_mel), h_pos(size_t(B) * n_mel);
  // The quirk only exists where the reference itself runs (B <= 4: its KV / position buffers are
  // sized for 4 candidates, main.cpp:794-797); beyond that every candidate gets positions 0..501.
  if (c->cfg.parity_quirks && B <= 4) {
    const int per = 502 * B / 4;
    for (int i = 0; i < B; ++i)
      for (int cc = 0; cc < per; ++cc) {
        const size_t idx = size_t(i) * per + cc;
        if (idx < pos.size()) pos[idx] = cc;
      }
  } else {
    for (int b = 0; b < B; ++b)
      for (int j = 0; j < 502; ++j) pos[size_t(b) * 502 + j] = j;
  }
  for (int b = 0; b < B; ++b)
    for (int j = 0; j < n_mel; ++j) {
      const int code = codes[size_t(b) * 502 + j];
      if (code < 0 || code >= kMelVocab) throw ArgError("mel code out of range");
      h_codes[size_t(b) * n_mel + j] = code;
      const int pj = pos[size_t(b) * 502 + j];
      if (pj < 0 || pj >= 608) throw ArgError("mel position id out of range (608 rows)", TTS_ELIMIT);
      h_pos[size_t(b) * n_mel + j] = pj;
    }
  if (!s.d_codes) {
    TTS_CUDA_TRY(ctx_malloc(c, &s.d_codes, size_t(s.Bmax) * 502 * 4));
    TTS_CUDA_TRY(ctx_malloc(c, &s.d_pos, size_t(s.Bmax) * 502 * 4));
  }
  Launcher L{c->stream, c->use_pdl, &c->launches};
  TTS_CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
  TTS_CUDA_TRY(cudaMemcpyAsync(s.d_text, text, T * 4, cudaMemcpyHostToDevice, c->stream));
  TTS_CUDA_TRY(cudaMemcpyAsync(s.d_voice, voice, kDim * 4, cudaMemcpyHostToDevice, c->stream));
  TTS_CUDA_TRY(cudaMemcpyAsync(s.d_codes, h_codes.data(), h_codes.size() * 4, cudaMemcpyHostToDevice, c->stream));
  TTS_CUDA_TRY(cudaMemcpyAsync(s.d_pos, h_pos.data(), h_pos.size() * 4, cudaMemcpyHostToDevice, c->stream));
  TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
  memset(out, 0, size_t(B) * 500 * kDim * 4);
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int nb = std::min(chunk, B - b0);
    L(ar_embed_rows_kernel, dim3(R, nb), dim3(256), 0, (const int *)s.d_text, T, (const float *)s.d_voice,
      (const int *)(s.d_codes + size_t(b0) * n_mel), (const int *)(s.d_pos + size_t(b0) * n_mel), n_mel,
      (const float *)m.text_emb, (const float *)m.text_pos, (const float *)m.mel_emb,
      (const float *)m.mel_pos, s.H);
    enqueue_rows_layers(c, L, nb, R, 0);
    // z = LN(LN(h) ln_f) lm_head.0  (main.cpp:2475-2499), all rows; mel rows are copied out
    L(ln_rows_kernel, dim3(nb * R), dim3(256), 0, (const float *)s.H, s.Z, (__half *)nullptr, (__half *)nullptr,
      (const float *)m.lnf_w, (const float *)m.lnf_b, (const float *)m.lm0_w, (const float *)m.lm0_b, kDim, kDim);
    for (int b = 0; b < nb; ++b)
      TTS_CUDA_TRY(cudaMemcpyAsync(out + size_t(b0 + b) * 500 * kDim, s.Z + (size_t(b) * R + 1 + T) * kDim,
                                   size_t(n_mel) * kDim * 4, cudaMemcpyDeviceToHost, c->stream));
    TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
  }
  TTS_CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
  TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
  TTS_CUDA_TRY(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  c->total_ms += c->last_ms;
}

// Decode-step benchmark: `iters` consecutive steps (persistent kernel or per-op graph, whatever
// the context is configured for) between two CUDA events, no host round trip in between.
// bytes = algorithmic bytes of one step at the mean KV length: streamed weights + embeddings
// + KV read/append + logits.
void ar_bench_step(tts_ctx *c, int iters, float *ms, double *bytes) {
  ArState &s = c->ars;
  if (!c->ar.loaded || s.B == 0) throw ArgError("tts_bench_decode_step before tts_ar_prefill");
  if (iters < 1 || s.n_past + iters > s.P) throw ArgError("not enough KV slots for the requested iterations", TTS_ELIMIT);
  std::vector<int32_t> toks(s.B, 100);
  const int n0 = s.n_past;
  ar_step(c, toks.data(), 2, nullptr, false);  // warm-up (graph build / attribute setup)
  TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
  cudaEvent_t e0, e1;
  TTS_CUDA_TRY(cudaEventCreate(&e0));
  TTS_CUDA_TRY(cudaEventCreate(&e1));
  TTS_CUDA_TRY(cudaEventRecord(e0, c->stream));
  for (int i = 1; i < iters; ++i) ar_step(c, toks.data(), 2 + (i % 500), nullptr, false);
  TTS_CUDA_TRY(cudaEventRecord(e1, c->stream));
  TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
  float t = 0;
  TTS_CUDA_TRY(cudaEventElapsedTime(&t, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const int timed = iters - 1;
  *ms = timed > 0 ? t / timed : 0.f;
  const double n_mean = n0 + 1 + 0.5 * iters;
  const double B = s.B;
  // (shared-prefix KV, SURVEY 8d: the prompt rows are read once per step, not once per candidate)
  const double kv_rows = s.n_prefix > 0 ? s.n_prefix + B * (n_mean + 1 - s.n_prefix) : B * (n_mean + 1);
  *bytes = double(c->ar.decode_weight_bytes) + 2.0 * kLayers * kDim * kv_rows * 2.0 /*KV read, f16*/ +
           2.0 * kLayers * kDim * B * 2.0 /*KV append*/ + (B + 1) * kDim * 4.0 /*embedding rows*/ +
           double(kMelVocab) * B * 4.0 /*logits*/;
}

// Streaming-GEMV micro-benchmark over the 30 layers' weights (larger than L2 in total, so
// every launch streams from HBM): average launch time from CUDA events on the stream.
void ar_bench_gemv(tts_ctx *c, int op, int B, int iters, float *ms, double *bytes) {
  ArModel &m = c->ar;
  ArState &s = c->ars;
  if (!m.loaded) throw ArgError("AR model not loaded");
  if (B < 1 || B > s.Bmax) throw ArgError("bad B");
  int64_t dummy = 0;
  Launcher L{c->stream, c->use_pdl, &dummy};
  TTS_CUDA_TRY(cudaMemsetAsync(s.h, 0, size_t(B) * kDim * 4, c->stream));
  TTS_CUDA_TRY(cudaMemsetAsync(s.m, 0, size_t(B) * kFF * 4, c->stream));
  TTS_CUDA_TRY(cudaMemsetAsync(s.d_state, 0, 16, c->stream));
  int N = 0, K = 0;
  auto one = [&](int layer) {
    ArLayer &l = m.layers[layer % kLayers];
    GemvArgs a{};
    switch (op) {
      case 0: a = gemv_args(l.w_qkv, l.b_qkv, s.h, s.q, 3072, kDim, B, PRO_LN, EPI_QKV);
        a.ln_w = l.ln1_w; a.ln_b = l.ln1_b; a.kcache = s.kc; a.vcache = s.vc; a.state = s.d_state;
        a.kv_b_stride = kHeads * s.P * kHeadDim; break;
      case 1: a = gemv_args(l.w_proj, l.b_proj, s.attn, s.q, kDim, kDim, B, PRO_NONE, EPI_STORE); break;
      case 2: a = gemv_args(l.w_fc, l.b_fc, s.h, s.m, kFF, kDim, B, PRO_LN, EPI_GELU16);
        a.ln_w = l.ln2_w; a.ln_b = l.ln2_b; break;
      case 3: a = gemv_args(l.w_proj2, l.b_proj2, s.m, s.q, kDim, kFF, B, PRO_NONE, EPI_STORE); break;
      default: throw ArgError("bad op");
    }
    N = a.N; K = a.K;
    launch_gemv(c, L, a);
  };
  for (int i = 0; i < kLayers; ++i) one(i);  // warm-up, also evicts L2 with 30 distinct matrices
  TTS_CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
  for (int i = 0; i < iters; ++i) one(i);
  TTS_CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
  TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
  float t = 0;
  TTS_CUDA_TRY(cudaEventElapsedTime(&t, c->ev0, c->ev1));
  *ms = t / iters;
  // algorithmic bytes of one launch: K*N*s_w + B*K*4 (activations) + B*N*4 (outputs) + N*4 (bias)
  *bytes = double(N) * K * wbytes(m.dtype) + double(B) * K * 4 + double(B) * N * 4 + double(N) * 4;
  c->launches += iters + kLayers;
}

}  // namespace tts
