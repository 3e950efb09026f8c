// diffusion.cu -- diffusion stage: loader, conditioning pre-pass, denoiser pass, sampler.
// Reference: diffusion_model_load (main.cpp:931-1634), diffusion_graph (main.cpp:3066-4044),
// diffusion() (main.cpp:5614-6042).  B200-first restructuring (all exact w.r.t. the
// reference's math): time-major f16 conv operands (implicit GEMM on tensor cores, no im2col),
// conditioned + unconditioned passes batched as two sequences of one launch set, the
// timestep-invariant conditioning branch hoisted out of the step loop, every timestep's
// embedding MLP + the 16 emb_layers evaluated once per utterance as three GEMMs, the DDPM
// update fused on the device (noise still drawn by the host in the reference's RNG order).
#include <set>

#include "diff_kernels.cuh"
#include "engine.h"
#include "gemm_launch.cuh"
#include "host/host_math.h"

namespace tts {

struct DRes {
  float *gn1_w, *gn1_b, *b_in2, *gn2_w, *gn2_b, *b_out3;
  __half *w_in2, *w_out3;
};
struct DAttn {
  float *n_w, *n_b, *b_qkv, *b_proj, *relbias;
  __half *w_qkv, *proj_hi, *proj_lo;
};
struct DiffModel {
  bool loaded = false;
  DAttn lc[4];
  __half *lc_conv_w;
  float *lc_conv_b, *code_norm_w, *code_norm_b, *cond_latent, *uncond_emb;
  __half *te0_hi, *te0_lo, *te2_hi, *te2_lo;
  float *te0_b, *te2_b;
  DRes res[16];    // 0-2 integrator, 3-12 main (with attention), 13-15 main tail
  DAttn attn[13];  // 0-2 integrator, 3-12 main
  __half *emb_hi, *emb_lo;  // [16*2048][1024] stacked emb_layers.1
  float *emb_b;
  __half *w_inp, *w_integ, *w_out;
  float *b_inp, *b_integ, *b_out, *out_gn_w, *out_gn_b;
  // work buffers
  int capS = 0, capSteps = 0, capU = 0;  // capacities: frames per sequence, sampling steps, utterances per batch
  float *X = nullptr, *CW = nullptr, *CE = nullptr, *H1 = nullptr, *OUT = nullptr, *INP = nullptr;
  float *stats = nullptr, *x_dev = nullptr, *noise_dev = nullptr, *lat_dev = nullptr;
  __half *A16 = nullptr, *CAT16 = nullptr, *XIN16 = nullptr, *ATThi = nullptr, *ATTlo = nullptr, *QKV16 = nullptr;
  int *rpb = nullptr, *up_idx = nullptr;
  float *TE = nullptr, *T0 = nullptr, *TEMB = nullptr, *EMB = nullptr;
  __half *P_hi = nullptr, *P_lo = nullptr;  // [steps][1024] planes scratch
  DdpmCoef *coefs = nullptr;
  int *d_step = nullptr;           // device-side sampling-step counter (graph replay)
  cudaGraphExec_t step_graph = nullptr;
  int graph_S = -1;
  int graph_kernels = 0;           // kernel nodes of step_graph (what one replay launches)
  int graph_U = 0;
  size_t noise_cap = 0;
  int run_S = 0, run_steps = 0, run_i = -1;  // streaming sampler state (run_S = common row stride = longest utterance)
  // utterance batching: U utterances = 2 U sequences (2u cond, 2u + 1 uncond) of different lengths on one launch set
  int run_U = 0;
  std::vector<int> run_Su;            // frames per utterance
  std::vector<long long> run_xoff;    // offset of utterance u in x_dev (its own [100][S_u] layout)
  std::vector<long long> run_noff;    // offset of utterance u's noise blocks in noise_dev
  int *d_Tseq = nullptr, *d_Sutt = nullptr;       // device: valid frames per sequence [2U] / per utterance [U]
  long long *d_xoff = nullptr, *d_noff = nullptr; // device copies of run_xoff / run_noff
  const int *tseq = nullptr;          // what the launch helpers pass as per-sequence lengths (null: all sequences full)
  size_t x_cap = 0;
  double *gn_partial = nullptr;    // fused GroupNorm statistics written by the last GEMM epilogue
  const float *partial_src = nullptr;  // ... and the tensor they describe (null = stale)
  int partial_mtiles = 0;
  float *h_pin = nullptr;  // pinned staging for x / outputs
  size_t h_pin_bytes = 0;
  // cached conditioning
  int cond_L = -1, cond_S = -1;
};

static __half *upload_conv_w(tts_ctx *c, const Container &ct, const std::string &name, int OC, int IC, int K,
                             int ICpad) {
  auto it = ct.tensors.find(name);
  if (it == ct.tensors.end()) throw ArgError("tensor '" + name + "' missing from " + ct.path, TTS_EIO);
  if (it->second.nelem != size_t(OC) * IC * K) throw ArgError("tensor '" + name + "' has wrong size in model file", TTS_EIO);
  // reference check (main.cpp:1585-1592): ne[0], ne[1] of the ggml shape [K, IC, OC] (2-D for 1x1)
  const auto &ne = it->second.ne;
  const int e0 = K > 1 ? K : IC, e1 = K > 1 ? IC : OC;
  if (ne[0] != e0 || (ne.size() > 1 ? ne[1] : 1) != e1)
    throw ArgError("tensor '" + name + "' has wrong shape in model file", TTS_EIO);
  size_t n = 0;
  read_tensor_to_staging(c, ct, name, &n);
  __half *d = nullptr;
  TTS_CUDA_TRY(ctx_malloc(c, &d, size_t(K) * OC * ICpad * 2));
  conv_weight_kernel<<<1024, 256, 0, c->stream>>>(c->d_scratch, d, OC, IC, K, ICpad);
  TTS_CUDA_TRY(cudaGetLastError());
  TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return d;
}

static void upload_planes(tts_ctx *c, const Container &ct, const std::string &name, int N, int K, __half *hi,
                          __half *lo) {
  auto it = ct.tensors.find(name);
  if (it == ct.tensors.end()) throw ArgError("tensor '" + name + "' missing from " + ct.path, TTS_EIO);
  const auto &ne = it->second.ne;
  if (it->second.nelem != size_t(N) * K || ne[0] != K || (ne.size() > 1 ? ne[1] : 1) != N)
    throw ArgError("tensor '" + name + "' has wrong shape in model file", TTS_EIO);
  size_t n = 0;
  read_tensor_to_staging(c, ct, name, &n);
  silu_split_kernel<<<1024, 256, 0, c->stream>>>(c->d_scratch, hi, lo, n, 0);
  TTS_CUDA_TRY(cudaGetLastError());
  TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
}

void diff_load(tts_ctx *c, const char *path) {
  Container ct;
  std::string err;
  if (!ct.open(path, err)) throw ArgError(err, TTS_EIO);
  if (c->diff && c->diff->loaded) throw ArgError("diffusion model already loaded in this context");
  if (!c->diff) c->diff = new DiffModel();
  DiffModel &m = *c->diff;
  std::set<std::string> known;
  auto f32 = [&](const std::string &n, std::vector<int> ne) { known.insert(n); return upload_f32(c, ct, n, ne); };
  auto convw = [&](const std::string &n, int OC, int IC, int K, int ICpad) {
    known.insert(n);
    return upload_conv_w(c, ct, n, OC, IC, K, ICpad);
  };
  auto planes = [&](const std::string &n, int N, int K, __half **hi, __half **lo) {
    known.insert(n);
    TTS_CUDA_TRY(ctx_malloc(c, hi, size_t(N) * K * 2));
    TTS_CUDA_TRY(ctx_malloc(c, lo, size_t(N) * K * 2));
    upload_planes(c, ct, n, N, K, *hi, *lo);
  };
  auto load_attn = [&](const std::string &p, DAttn &a) {
    a.n_w = f32(p + "norm.weight", {1024});
    a.n_b = f32(p + "norm.bias", {1024});
    a.w_qkv = convw(p + "qkv.weight", 3072, 1024, 1, 1024);
    a.b_qkv = f32(p + "qkv.bias", {3072});
    planes(p + "proj_out.weight", 1024, 1024, &a.proj_hi, &a.proj_lo);
    a.b_proj = f32(p + "proj_out.bias", {1024});
    a.relbias = f32(p + "relative_pos_embeddings.relative_attention_bias.weight", {16, 32});
  };
  TTS_CUDA_TRY(ctx_malloc(c, &m.emb_hi, size_t(16) * 2048 * 1024 * 2));
  TTS_CUDA_TRY(ctx_malloc(c, &m.emb_lo, size_t(16) * 2048 * 1024 * 2));
  TTS_CUDA_TRY(ctx_malloc(c, &m.emb_b, size_t(16) * 2048 * 4));
  auto load_res = [&](const std::string &p, DRes &r, int idx) {
    r.gn1_w = f32(p + "in_layers.0.weight", {1024});
    r.gn1_b = f32(p + "in_layers.0.bias", {1024});
    r.w_in2 = convw(p + "in_layers.2.weight", 1024, 1024, 1, 1024);
    r.b_in2 = f32(p + "in_layers.2.bias", {1024});
    known.insert(p + "emb_layers.1.weight");
    upload_planes(c, ct, p + "emb_layers.1.weight", 2048, 1024, m.emb_hi + size_t(idx) * 2048 * 1024,
                  m.emb_lo + size_t(idx) * 2048 * 1024);
    float *eb = f32(p + "emb_layers.1.bias", {2048});
    TTS_CUDA_TRY(cudaMemcpy(m.emb_b + size_t(idx) * 2048, eb, 2048 * 4, cudaMemcpyDeviceToDevice));
    ctx_free(c, eb);
    r.gn2_w = f32(p + "out_layers.0.weight", {1024});
    r.gn2_b = f32(p + "out_layers.0.bias", {1024});
    r.w_out3 = convw(p + "out_layers.3.weight", 1024, 1024, 3, 1024);
    r.b_out3 = f32(p + "out_layers.3.bias", {1024});
  };
  m.cond_latent = f32("diffusion_conditioning_latent", {2048});
  m.lc_conv_w = convw("latent_conditioner.0.weight", 1024, 1024, 3, 1024);
  m.lc_conv_b = f32("latent_conditioner.0.bias", {1024});
  for (int i = 0; i < 4; ++i) load_attn("latent_conditioner." + std::to_string(i + 1) + ".", m.lc[i]);
  m.code_norm_w = f32("code_norm.weight", {1024});
  m.code_norm_b = f32("code_norm.bias", {1024});
  planes("time_embed.0.weight", 1024, 1024, &m.te0_hi, &m.te0_lo);
  m.te0_b = f32("time_embed.0.bias", {1024});
  planes("time_embed.2.weight", 1024, 1024, &m.te2_hi, &m.te2_lo);
  m.te2_b = f32("time_embed.2.bias", {1024});
  for (int i = 0; i < 3; ++i) {
    const std::string p = "conditioning_timestep_integrator." + std::to_string(i) + ".";
    load_res(p + "resblk.", m.res[i], i);
    load_attn(p + "attn.", m.attn[i]);
  }
  for (int i = 0; i < 10; ++i) {
    const std::string p = "layers." + std::to_string(i) + ".";
    load_res(p + "resblk.", m.res[3 + i], 3 + i);
    load_attn(p + "attn.", m.attn[3 + i]);
  }
  for (int i = 0; i < 3; ++i) load_res("layers." + std::to_string(10 + i) + ".", m.res[13 + i], 13 + i);
  m.w_inp = convw("inp_block.weight", 1024, 100, 3, 128);
  m.b_inp = f32("inp_block.bias", {1024});
  m.w_integ = convw("integrating_conv.weight", 1024, 2048, 1, 2048);
  m.b_integ = f32("integrating_conv.bias", {1024});
  m.out_gn_w = f32("out.0.weight", {1024});
  m.out_gn_b = f32("out.0.bias", {1024});
  m.w_out = convw("out.2.weight", 200, 1024, 3, 1024);
  m.b_out = f32("out.2.bias", {200});
  m.uncond_emb = f32("unconditioned_embedding", {1024});
  for (const auto &n : ct.order)
    if (!known.count(n)) throw ArgError("unknown tensor '" + n + "' in model file", TTS_EIO);
  m.loaded = true;
}

void diff_free(tts_ctx *c) {
  if (c->diff) {
    if (c->diff->step_graph) cudaGraphExecDestroy(c->diff->step_graph);
    delete c->diff;  // its buffers belong to the context's allocation registry (tts_free)
    c->diff = nullptr;
  }
}

template <typename T>
static void grow(tts_ctx *c, T **p, size_t n) {
  if (*p) ctx_free(c, *p);
  TTS_CUDA_TRY(ctx_malloc(c, p, n * sizeof(T)));
}

static void ensure_buffers(tts_ctx *c, int S, int steps, int U = 1) {
  DiffModel &m = *c->diff;
  if (S > m.capS || U > m.capU) {
    S = std::max(S, m.capS);
    U = std::max(U, m.capU);
    const size_t nseq = size_t(2) * U, s2 = nseq * S;
    grow(c, &m.X, s2 * kDim);
    grow(c, &m.CW, s2 * kDim);
    grow(c, &m.CE, s2 * kDim);
    grow(c, &m.H1, s2 * kDim);
    grow(c, &m.QKV16, s2 * 3072);
    grow(c, &m.OUT, s2 * 200);
    grow(c, &m.INP, size_t(U) * S * kDim);
    grow(c, &m.stats, nseq * 32 * 2);
    grow(c, &m.lat_dev, size_t(S) * kDim);
    grow(c, &m.A16, nseq * (S + 2) * kDim);
    grow(c, &m.CAT16, nseq * (S + 2) * 2048);
    grow(c, &m.XIN16, size_t(U) * (S + 2) * 128);
    grow(c, &m.ATThi, s2 * kDim);
    grow(c, &m.ATTlo, s2 * kDim);
    grow(c, &m.rpb, size_t(S) + 8);
    grow(c, &m.up_idx, size_t(S));
    grow(c, &m.gn_partial, nseq * 32 * 64 * 2);
    grow(c, &m.d_Tseq, nseq);
    grow(c, &m.d_Sutt, size_t(U));
    grow(c, &m.d_xoff, size_t(U));
    grow(c, &m.d_noff, size_t(U));
    if (!m.d_step) TTS_CUDA_TRY(ctx_malloc(c, &m.d_step, 4));
    m.partial_src = nullptr;  // buffers moved: no tensor has fused statistics
    if (m.step_graph) { cudaGraphExecDestroy(m.step_graph); m.step_graph = nullptr; m.graph_S = -1; }
    m.capS = S;
    m.capU = U;
    m.cond_L = m.cond_S = -1;
  }
  if (steps > m.capSteps) {
    grow(c, &m.TE, size_t(steps) * kDim);
    grow(c, &m.T0, size_t(steps) * kDim);
    grow(c, &m.TEMB, size_t(steps) * kDim);
    grow(c, &m.EMB, size_t(steps) * 16 * 2048);
    grow(c, &m.P_hi, size_t(steps) * kDim);
    grow(c, &m.P_lo, size_t(steps) * kDim);
    grow(c, &m.coefs, size_t(steps));
    if (m.step_graph) { cudaGraphExecDestroy(m.step_graph); m.step_graph = nullptr; m.graph_S = -1; }
    m.capSteps = steps;
  }
}

static void tg(tts_ctx *c, const Launcher &L, const __half *Ahi, const __half *Alo, const __half *Whi, const __half *Wlo,
               const float *bias, float *C, int M, int N, int K, int lda, int ldc, int epi, int taps = 1, int T = 0,
               int halo = 0, __half *C16 = nullptr) {
  DiffModel &m = *c->diff;
  const int Tt = T > 0 ? T : M;
  TGemmArgs g{Ahi, Alo, Whi, Wlo, bias, C, C16, nullptr, M, N, K, lda, ldc, C16 ? ldc : 0, epi, taps, 1, taps / 2, halo, Tt};
  g.Tseq = T > 0 ? m.tseq : nullptr;
  // outputs that feed a GroupNorm (x, h1, code embedding): let the epilogue produce the statistics
  const bool gn_target = T > 0 && N == kDim && (C == m.X || C == m.CW || C == m.H1) && Tt <= 64 * T5_BM;
  if (gn_target) g.gn_partial = m.gn_partial;
  const int produced = launch_gemm(L, g);
  if (produced) { m.partial_src = C; m.partial_mtiles = produced; }
  else if (C == m.partial_src) m.partial_src = nullptr;
}

// f16 x f16 -> f32 convolution over nseq sequences of T frames (halo 1)
static void conv(tts_ctx *c, const Launcher &L, const __half *X16, const __half *W, const float *bias, float *C,
                 int nseq, int T, int IC, int OC, int taps, int ldc, int epi) {
  tg(c, L, X16, nullptr, W, nullptr, bias, C, nseq * T, OC, IC, IC, ldc, epi, taps, T, 1);
}

static void gn(tts_ctx *c, const Launcher &L, const float *X, const float *w, const float *b, const float *ss,
               __half *out16, float *out32, int nseq, int T, int silu) {
  DiffModel &m = *c->diff;
  const bool fused = m.partial_src == X;
  if (out32 && out32 == m.partial_src) m.partial_src = nullptr;  // about to be overwritten
  if (!fused) L(gn_stats_kernel, dim3(32, nseq), dim3(256), 0, X, m.stats, T, m.tseq);
  const double *partial = fused ? m.gn_partial : nullptr;
  // 8 rows per block once that still fills the SMs twice, else 2 (S = 191: 194 blocks)
  if ((T + 2 + 7) / 8 * nseq >= 2 * 148)
    L(gn_apply_kernel<8>, dim3((T + 2 + 7) / 8, nseq), dim3(256), 0, X, (const float *)m.stats, w, b, ss, out16, out32, T, 1, kDim,
      silu, (const int *)m.d_step, 16 * 2048, partial, m.partial_mtiles, m.tseq);
  else
    L(gn_apply_kernel<2>, dim3((T + 2 + 1) / 2, nseq), dim3(256), 0, X, (const float *)m.stats, w, b, ss, out16, out32, T, 1, kDim,
      silu, (const int *)m.d_step, 16 * 2048, partial, m.partial_mtiles, m.tseq);
}

// ResBlock (main.cpp:3347-3480): x += conv3(silu((GN(h)w+b)(1+scale)+shift)),
// h = conv1(silu(GN(x)w+b)) + b
static void res_block(tts_ctx *c, const Launcher &L, const DRes &r, float *x, int nseq, int T, const float *ss) {
  DiffModel &m = *c->diff;
  gn(c, L, x, r.gn1_w, r.gn1_b, nullptr, m.A16, nullptr, nseq, T, 1);
  conv(c, L, m.A16, r.w_in2, r.b_in2, m.H1, nseq, T, kDim, kDim, 1, kDim, E_BIAS);
  gn(c, L, m.H1, r.gn2_w, r.gn2_b, ss, m.A16, nullptr, nseq, T, 1);
  conv(c, L, m.A16, r.w_out3, r.b_out3, x, nseq, T, kDim, kDim, 3, kDim, E_BIAS_RESID);
}

// AttentionBlock (main.cpp:3482-3609): x += proj_out(attn(conv1(GN(x)w+b)))
static void attn_block(tts_ctx *c, const Launcher &L, const DAttn &a, float *x, int nseq, int T) {
  DiffModel &m = *c->diff;
  gn(c, L, x, a.n_w, a.n_b, nullptr, m.A16, nullptr, nseq, T, 0);
  // q | k | v leave the GEMM's epilogue as f16 (the attention's tensor-core operands); no f32 copy is kept
  tg(c, L, m.A16, nullptr, a.w_qkv, nullptr, a.b_qkv, nullptr, nseq * T, 3072, kDim, kDim, 3072, E_BIAS, 1, T, 1, m.QKV16);
  // 64 queries per block once that fills the SMs twice over, else 32 (S = 191: 192 blocks instead of 96)
  if (((T + 63) / 64) * kHeads * nseq >= 2 * 148) {
    ensure_smem_attr(diff_attn_tc_kernel<4>, ta_smem_bytes<4>(4096));
    L(diff_attn_tc_kernel<4>, dim3((T + 63) / 64, kHeads, nseq), dim3(128), ta_smem_bytes<4>(T), (const __half *)m.QKV16,
      (const float *)a.relbias, (const int *)m.rpb, m.ATThi, m.ATTlo, T, m.tseq);
  } else {
    ensure_smem_attr(diff_attn_tc_kernel<2>, ta_smem_bytes<2>(4096));
    L(diff_attn_tc_kernel<2>, dim3((T + 31) / 32, kHeads, nseq), dim3(64), ta_smem_bytes<2>(T), (const __half *)m.QKV16,
      (const float *)a.relbias, (const int *)m.rpb, m.ATThi, m.ATTlo, T, m.tseq);
  }
  tg(c, L, m.ATThi, m.ATTlo, a.proj_hi, a.proj_lo, a.b_proj, x, nseq * T, kDim, kDim, kDim, kDim, E_BIAS_RESID,
     1, T, 0);  // per-sequence M tiles so the fused GroupNorm statistics stay per sequence
}

// timestep-invariant conditioning branch -> CE[0] (conditioned, stretched to S), CE[1]
// (unconditioned broadcast).  main.cpp:3157-3328.
// Utterance batching: utterance u of a batch has its own Lf and S (<= Smax, the common row stride) and owns
// sequences 2u / 2u + 1 of CE; rows [S, Smax) of its sequences are zero padding.
static void prepare_conditioning(tts_ctx *c, const Launcher &L, const float *latents_host, int Lf, int S, int u = 0,
                                 int Smax = 0) {
  DiffModel &m = *c->diff;
  if (Smax <= 0) Smax = S;
  m.tseq = nullptr;  // the conditioning branch works on ONE sequence of Lf rows
  std::vector<int> rpb = tts_host::relative_position_table(Smax + 8);
  std::vector<int> up = tts_host::upscale_index(Lf, S);
  TTS_CUDA_TRY(cudaMemcpyAsync(m.rpb, rpb.data(), size_t(Smax + 8) * 4, cudaMemcpyHostToDevice, c->stream));
  TTS_CUDA_TRY(cudaMemcpyAsync(m.up_idx, up.data(), size_t(S) * 4, cudaMemcpyHostToDevice, c->stream));
  TTS_CUDA_TRY(cudaMemcpyAsync(m.lat_dev, latents_host, size_t(Lf) * kDim * 4, cudaMemcpyHostToDevice, c->stream));
  TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));  // host vectors go out of scope
  L(to_f16_halo_kernel, dim3(Lf + 2), dim3(256), 0, (const float *)m.lat_dev, m.A16, Lf, kDim, 1, kDim);
  conv(c, L, m.A16, m.lc_conv_w, m.lc_conv_b, m.CW, 1, Lf, kDim, kDim, 3, kDim, E_BIAS);
  for (int i = 0; i < 4; ++i) attn_block(c, L, m.lc[i], m.CW, 1, Lf);
  // code_norm, then * (1 + cond_latent[:1024]) + cond_latent[1024:]  (main.cpp:3293-3317)
  // (cond_latent is not a per-step table: d_step is 0 whenever the conditioning branch runs)
  gn(c, L, m.CW, m.code_norm_w, m.code_norm_b, m.cond_latent, nullptr, m.H1, 1, Lf, 0);
  L(code_emb_kernel, dim3(Smax, 2), dim3(256), 0, (const float *)m.H1, (const int *)m.up_idx,
    (const float *)m.uncond_emb, m.CE + size_t(2 * u) * Smax * kDim, Smax, S);
  m.cond_L = Lf;
  m.cond_S = S;
}

// time-embedding MLP + all 16 emb_layers for `steps` timesteps (main.cpp:3331-3343, 3418-3440)
static void prepare_time(tts_ctx *c, const Launcher &L, const std::vector<int> &timesteps) {
  DiffModel &m = *c->diff;
  const int n = int(timesteps.size());
  std::vector<float> te(size_t(n) * kDim);
  for (int i = 0; i < n; ++i) tts_host::timestep_embedding(timesteps[i], te.data() + size_t(i) * kDim);
  TTS_CUDA_TRY(cudaMemcpyAsync(m.TE, te.data(), te.size() * 4, cudaMemcpyHostToDevice, c->stream));
  TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
  const size_t ne = size_t(n) * kDim;
  L(silu_split_kernel, dim3(64), dim3(256), 0, (const float *)m.TE, m.P_hi, m.P_lo, ne, 0);
  tg(c, L, m.P_hi, m.P_lo, m.te0_hi, m.te0_lo, m.te0_b, m.T0, n, kDim, kDim, kDim, kDim, E_BIAS);
  L(silu_split_kernel, dim3(64), dim3(256), 0, (const float *)m.T0, m.P_hi, m.P_lo, ne, 1);
  tg(c, L, m.P_hi, m.P_lo, m.te2_hi, m.te2_lo, m.te2_b, m.TEMB, n, kDim, kDim, kDim, kDim, E_BIAS);
  L(silu_split_kernel, dim3(64), dim3(256), 0, (const float *)m.TEMB, m.P_hi, m.P_lo, ne, 1);
  tg(c, L, m.P_hi, m.P_lo, m.emb_hi, m.emb_lo, m.emb_b, m.EMB, n, 16 * 2048, kDim, kDim, 16 * 2048, E_BIAS);
}

// One denoiser evaluation on nseq sequences.  CW must hold the code embedding of each
// sequence, x_dev the current x.  emb: this step's [16][2048] scale|shift table.
// Batched sampling: nseq = 2 U sequences with row stride S and per-sequence lengths d_Tseq (m.tseq), x of the
// U utterances in x_dev at d_xoff; the input block runs once per utterance (d_Sutt).
static void run_denoiser(tts_ctx *c, const Launcher &L, int nseq, int S, const float *emb, int U = 1) {
  DiffModel &m = *c->diff;
  const int *tseq_all = m.tseq;
  const bool batched = tseq_all != nullptr;
  for (int i = 0; i < 3; ++i) {
    res_block(c, L, m.res[i], m.CW, nseq, S, emb + size_t(i) * 2048);
    attn_block(c, L, m.attn[i], m.CW, nseq, S);
  }
  L(xin_kernel, dim3(S + 2, U), dim3(128), 0, (const float *)m.x_dev, m.XIN16, S, (const int *)(batched ? m.d_Sutt : nullptr),
    (const long long *)(batched ? m.d_xoff : nullptr));
  m.tseq = batched ? m.d_Sutt : nullptr;  // the input block runs per UTTERANCE (cond and uncond share x)
  conv(c, L, m.XIN16, m.w_inp, m.b_inp, m.INP, U, S, 128, kDim, 3, kDim, E_BIAS);
  m.tseq = tseq_all;
  L(concat_kernel, dim3(S + 2, nseq), dim3(256), 0, (const float *)m.INP, (const float *)m.CW, m.CAT16, S, m.tseq);
  conv(c, L, m.CAT16, m.w_integ, m.b_integ, m.X, nseq, S, 2048, kDim, 1, kDim, E_BIAS);
  for (int i = 0; i < 10; ++i) {
    res_block(c, L, m.res[3 + i], m.X, nseq, S, emb + size_t(3 + i) * 2048);
    attn_block(c, L, m.attn[3 + i], m.X, nseq, S);
  }
  for (int i = 0; i < 3; ++i) res_block(c, L, m.res[13 + i], m.X, nseq, S, emb + size_t(13 + i) * 2048);
  gn(c, L, m.X, m.out_gn_w, m.out_gn_b, nullptr, m.A16, nullptr, nseq, S, 1);
  conv(c, L, m.A16, m.w_out, m.b_out, m.OUT, nseq, S, kDim, 200, 3, 200, E_BIAS);
}

static float *pin(tts_ctx *c, size_t bytes) {
  DiffModel &m = *c->diff;
  if (m.h_pin_bytes < bytes) {
    if (m.h_pin) ctx_free_host(c, m.h_pin);
    TTS_CUDA_TRY(ctx_malloc_host(c, &m.h_pin, bytes));
    m.h_pin_bytes = bytes;
  }
  return m.h_pin;
}

static void check_sizes(tts_ctx *c, int Lf, int S) {
  if (!c->diff || !c->diff->loaded) throw ArgError("diffusion model not loaded");
  if (Lf < 1 || Lf > 500) throw ArgError("latent length must be in [1,500]", TTS_ELIMIT);
  if (S < Lf || S > 4096) throw ArgError("bad mel length S", TTS_ELIMIT);
}

void diff_eps(tts_ctx *c, const float *latents, int Lf, const float *x, int S, int timestep, int cond_free,
              float *out) {
  check_sizes(c, Lf, S);
  if (timestep < 0 || timestep >= 4000) throw ArgError("timestep out of range");
  DiffModel &m = *c->diff;
  ensure_buffers(c, S, 1);
  Launcher L{c->stream, c->use_pdl, &c->launches};
  TTS_CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
  TTS_CUDA_TRY(cudaMemsetAsync(m.d_step, 0, 4, c->stream));
  prepare_conditioning(c, L, latents, Lf, S);
  prepare_time(c, L, {timestep});
  if (m.x_cap < size_t(100) * S) {
    if (m.x_dev) ctx_free(c, m.x_dev);
    m.x_dev = nullptr;
    TTS_CUDA_TRY(ctx_malloc(c, &m.x_dev, size_t(100) * S * 4));
    m.x_cap = size_t(100) * S;
    if (m.step_graph) { cudaGraphExecDestroy(m.step_graph); m.step_graph = nullptr; m.graph_S = -1; }
  }
  TTS_CUDA_TRY(cudaMemcpyAsync(m.x_dev, x, size_t(100) * S * 4, cudaMemcpyHostToDevice, c->stream));
  TTS_CUDA_TRY(cudaMemcpyAsync(m.CW, m.CE + (cond_free ? size_t(S) * kDim : 0), size_t(S) * kDim * 4,
                               cudaMemcpyDeviceToDevice, c->stream));
  m.partial_src = nullptr;
  run_denoiser(c, L, 1, S, m.EMB);
  // OUT is time-major [S][200]; the C-ABI returns the reference's [200][S]
  float *h = pin(c, size_t(S) * 200 * 4);
  TTS_CUDA_TRY(cudaMemcpyAsync(h, m.OUT, size_t(S) * 200 * 4, cudaMemcpyDeviceToHost, c->stream));
  TTS_CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
  TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
  TTS_CUDA_TRY(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  c->total_ms += c->last_ms;
  for (int t = 0; t < S; ++t)
    for (int ch = 0; ch < 200; ++ch) out[size_t(ch) * S + t] = h[size_t(t) * 200 + ch];
}

// Streaming form of the sampling loop: begin (conditioning, schedule, x0) / step (upload the
// step's noise block, enqueue the captured step graph -- asynchronous) / end (read the mel).
// The host draws noise block i+1 from the reference's RNG stream while the GPU runs step i.
// kernel nodes of a captured graph: what one replay launches (the launch counter is a count, not an estimate)
static int count_kernel_nodes(cudaGraph_t graph) {
  size_t n = 0;
  TTS_CUDA_TRY(cudaGraphGetNodes(graph, nullptr, &n));
  std::vector<cudaGraphNode_t> nodes(n);
  if (n) TTS_CUDA_TRY(cudaGraphGetNodes(graph, nodes.data(), &n));
  int k = 0;
  for (size_t i = 0; i < n; ++i) {
    cudaGraphNodeType t;
    TTS_CUDA_TRY(cudaGraphNodeGetType(nodes[i], &t));
    if (t == cudaGraphNodeTypeKernel) ++k;
  }
  return k;
}

// Measurement only (tortoise_b200_bench.h): the denoiser's 3-tap convolution -- the GEMM shape the diffusion
// stage spends most of its time in -- `iters` back-to-back launches on the model's own weights between two
// CUDA events.  M = 2 S rows (cond + uncond), N = K = 1024.
void diff_bench_conv3(tts_ctx *c, int S, int iters, float *ms, double *flop) {
  if (!c->diff || !c->diff->loaded) throw ArgError("diffusion model not loaded");
  if (S < 8 || S > 4096 || iters < 1) throw ArgError("bad argument");
  DiffModel &m = *c->diff;
  ensure_buffers(c, S, 1);
  int64_t dummy = 0;
  Launcher L{c->stream, c->use_pdl, &dummy};
  TTS_CUDA_TRY(cudaMemsetAsync(m.A16, 0, size_t(2) * (S + 2) * kDim * 2, c->stream));
  m.partial_src = nullptr;
  for (int i = 0; i < 3; ++i) conv(c, L, m.A16, m.res[3 + i].w_out3, m.res[3 + i].b_out3, m.H1, 2, S, kDim, kDim, 3, kDim, E_BIAS);
  TTS_CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
  for (int i = 0; i < iters; ++i) conv(c, L, m.A16, m.res[3 + i % 13].w_out3, m.res[3 + i % 13].b_out3, m.H1, 2, S, kDim, kDim, 3, kDim, E_BIAS);
  TTS_CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
  TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
  float t = 0;
  TTS_CUDA_TRY(cudaEventElapsedTime(&t, c->ev0, c->ev1));
  m.partial_src = nullptr;
  *ms = t / iters;
  *flop = 2.0 * (2.0 * S) * kDim * kDim * 3;
  c->launches += iters + 3;
}

// Streaming form of the sampling loop for a BATCH of U utterances (U = 1: the reference's loop): begin
// (conditioning per utterance, schedule, x0) / step (upload every utterance's noise block, enqueue the captured
// step graph -- asynchronous) / end (read the mels).  The host draws the noise of step i + 1 from each
// utterance's own RNG stream while the GPU runs step i.  Utterance u owns sequences 2u (cond) and 2u + 1
// (uncond) of every activation buffer; all sequences share the row stride Smax = max S_u and carry their own
// length, so one launch set serves utterances of different lengths (BASELINE configs[4]) and the GEMMs see
// M = 2 * sum(S_u) rows instead of 2 S.
void diff_begin_batch(tts_ctx *c, int U, const float *const *latents, const int32_t *Lf, const int32_t *S, int n_steps,
                      const float *const *x0) {
  if (U < 1 || U > 64) throw ArgError("batch of 1..64 utterances", TTS_ELIMIT);
  if (n_steps < 1 || n_steps > 4000) throw ArgError("bad n_steps");
  int Smax = 0;
  for (int u = 0; u < U; ++u) {
    check_sizes(c, Lf[u], S[u]);
    Smax = std::max(Smax, int(S[u]));
  }
  DiffModel &m = *c->diff;
  ensure_buffers(c, Smax, n_steps, U);
  Smax = U > 1 ? Smax : S[0];
  Launcher L{c->stream, c->use_pdl, &c->launches};
  TTS_CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
  TTS_CUDA_TRY(cudaMemsetAsync(m.d_step, 0, 4, c->stream));  // conditioning branch must see step 0
  const std::vector<tts_host::DdpmStep> sched = tts_host::ddpm_schedule(n_steps);
  std::vector<DdpmCoef> coefs(n_steps);
  std::vector<int> timesteps(n_steps);
  for (int i = 0; i < n_steps; ++i) {
    const auto &sd = sched[i];
    coefs[i] = DdpmCoef{sd.cfk, sd.sqrt_recip, sd.sqrt_recipm1, sd.coef1, sd.coef2, sd.min_log, sd.max_log, sd.last};
    timesteps[i] = sd.timestep;
  }
  TTS_CUDA_TRY(cudaMemcpyAsync(m.coefs, coefs.data(), coefs.size() * sizeof(DdpmCoef), cudaMemcpyHostToDevice, c->stream));
  m.run_U = U;
  m.run_Su.assign(S, S + U);
  m.run_xoff.assign(U, 0);
  m.run_noff.assign(U, 0);
  size_t nx_total = 0;
  for (int u = 0; u < U; ++u) {
    m.run_xoff[u] = (long long)nx_total;
    m.run_noff[u] = (long long)(nx_total * size_t(n_steps + 1));
    nx_total += size_t(100) * S[u];
  }
  if (U > 1) TTS_CUDA_TRY(cudaMemsetAsync(m.CE, 0, size_t(2) * U * Smax * kDim * 4, c->stream));
  for (int u = 0; u < U; ++u) prepare_conditioning(c, L, latents[u], Lf[u], S[u], u, Smax);
  prepare_time(c, L, timesteps);
  // noise: per utterance, block 0 = initial x, block i + 1 = the draw of step i (always drawn, used unless last)
  // (the graph embeds these pointers: reallocation drops the captured graph)
  if (m.noise_cap < size_t(n_steps + 1) * nx_total || m.x_cap < nx_total) {
    if (m.noise_dev) ctx_free(c, m.noise_dev);
    if (m.x_dev) ctx_free(c, m.x_dev);
    m.noise_dev = m.x_dev = nullptr;
    TTS_CUDA_TRY(ctx_malloc(c, &m.noise_dev, size_t(n_steps + 1) * nx_total * 4));
    TTS_CUDA_TRY(ctx_malloc(c, &m.x_dev, nx_total * 4));
    m.noise_cap = size_t(n_steps + 1) * nx_total;
    m.x_cap = nx_total;
    if (m.step_graph) { cudaGraphExecDestroy(m.step_graph); m.step_graph = nullptr; m.graph_S = -1; }
  }
  float *hp = pin(c, size_t(n_steps + 1) * nx_total * 4);
  for (int u = 0; u < U; ++u) {
    const size_t nx = size_t(100) * S[u];
    memcpy(hp + m.run_noff[u], x0[u], nx * 4);
    TTS_CUDA_TRY(cudaMemcpyAsync(m.noise_dev + m.run_noff[u], hp + m.run_noff[u], nx * 4, cudaMemcpyHostToDevice, c->stream));
    TTS_CUDA_TRY(cudaMemcpyAsync(m.x_dev + m.run_xoff[u], m.noise_dev + m.run_noff[u], nx * 4, cudaMemcpyDeviceToDevice, c->stream));
  }
  if (U > 1) {  // per-sequence / per-utterance lengths and offsets the batched kernels read
    std::vector<int> tseq(2 * U);
    for (int u = 0; u < U; ++u) tseq[2 * u] = tseq[2 * u + 1] = S[u];
    TTS_CUDA_TRY(cudaMemcpyAsync(m.d_Tseq, tseq.data(), tseq.size() * 4, cudaMemcpyHostToDevice, c->stream));
    TTS_CUDA_TRY(cudaMemcpyAsync(m.d_Sutt, m.run_Su.data(), size_t(U) * 4, cudaMemcpyHostToDevice, c->stream));
    TTS_CUDA_TRY(cudaMemcpyAsync(m.d_xoff, m.run_xoff.data(), size_t(U) * 8, cudaMemcpyHostToDevice, c->stream));
    TTS_CUDA_TRY(cudaMemcpyAsync(m.d_noff, m.run_noff.data(), size_t(U) * 8, cudaMemcpyHostToDevice, c->stream));
    TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));  // host vector goes out of scope
  }
  TTS_CUDA_TRY(cudaMemsetAsync(m.d_step, 0, 4, c->stream));
  m.run_S = Smax;
  m.run_steps = n_steps;
  m.run_i = 0;
}

void diff_step_batch(tts_ctx *c, const float *const *noise_blocks) {
  if (!c->diff || c->diff->run_i < 0 || c->diff->run_i >= c->diff->run_steps) throw ArgError("tts_diffusion_step out of sequence");
  DiffModel &m = *c->diff;
  const int S = m.run_S, i = m.run_i, U = m.run_U;
  Launcher L{c->stream, c->use_pdl, &c->launches};
  for (int u = 0; u < U; ++u) {
    const size_t nx = size_t(100) * m.run_Su[u];
    float *hp = m.h_pin + m.run_noff[u] + size_t(i + 1) * nx;  // pinned slot of this block (stable until the copy ran)
    memcpy(hp, noise_blocks[u], nx * 4);
    TTS_CUDA_TRY(cudaMemcpyAsync(m.noise_dev + m.run_noff[u] + size_t(i + 1) * nx, hp, nx * 4, cudaMemcpyHostToDevice, c->stream));
  }
  // One sampling step = memcpy(code embedding) + ~125 kernels + DDPM update + step counter.
  // Everything that varies per step is read through the device-side counter d_step, so the
  // step is captured once into a CUDA graph and replayed n_steps times.
  auto enqueue_step = [&](const Launcher &LL) {
    TTS_CUDA_TRY(cudaMemcpyAsync(m.CW, m.CE, size_t(2) * U * S * kDim * 4, cudaMemcpyDeviceToDevice, c->stream));
    m.partial_src = nullptr;  // CW was overwritten by a copy: its fused statistics are stale
    m.tseq = U > 1 ? m.d_Tseq : nullptr;
    run_denoiser(c, LL, 2 * U, S, m.EMB, U);
    m.tseq = nullptr;
    const size_t nx_max = size_t(100) * S;
    LL(ddpm_step_kernel, dim3(std::min(148, int((nx_max + 255) / 256)), U), dim3(256), 0, m.x_dev, (const float *)m.OUT,
       (const float *)m.noise_dev, (const DdpmCoef *)m.coefs, (const int *)m.d_step, S, (const int *)(U > 1 ? m.d_Sutt : nullptr),
       (const long long *)(U > 1 ? m.d_xoff : nullptr), (const long long *)(U > 1 ? m.d_noff : nullptr));
    LL(step_inc_kernel, dim3(1), dim3(32), 0, m.d_step);
  };
  if (c->use_graph) {
    if (!m.step_graph || m.graph_S != S || m.graph_U != U) {
      if (m.step_graph) cudaGraphExecDestroy(m.step_graph);
      m.step_graph = nullptr;
      int64_t dummy = 0;
      Launcher LG{c->stream, c->use_pdl, &dummy};
      cudaGraph_t graph;
      // capture must not record work that should run now: drain, then record one step
      TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
      TTS_CUDA_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
      try {
        enqueue_step(LG);
      } catch (...) {
        cudaGraph_t g2;
        cudaStreamEndCapture(c->stream, &g2);
        throw;
      }
      TTS_CUDA_TRY(cudaStreamEndCapture(c->stream, &graph));
      m.graph_kernels = count_kernel_nodes(graph);
      TTS_CUDA_TRY(cudaGraphInstantiate(&m.step_graph, graph, 0));
      cudaGraphDestroy(graph);
      m.graph_S = S;
      m.graph_U = U;
    }
    TTS_CUDA_TRY(cudaGraphLaunch(m.step_graph, c->stream));
    c->launches += m.graph_kernels;
  } else {
    enqueue_step(L);
  }
  m.run_i += 1;
}

void diff_end_batch(tts_ctx *c, float *const *mel) {
  if (!c->diff || c->diff->run_i != c->diff->run_steps || c->diff->run_steps <= 0) throw ArgError("tts_diffusion_end before all steps ran");
  DiffModel &m = *c->diff;
  size_t nx_total = 0;
  for (int u = 0; u < m.run_U; ++u) nx_total += size_t(100) * m.run_Su[u];
  float *h = m.h_pin;  // the x0 slots are long consumed
  TTS_CUDA_TRY(cudaMemcpyAsync(h, m.x_dev, nx_total * 4, cudaMemcpyDeviceToHost, c->stream));
  TTS_CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
  TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
  TTS_CUDA_TRY(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  c->total_ms += c->last_ms;
  for (int u = 0; u < m.run_U; ++u) memcpy(mel[u], h + m.run_xoff[u], size_t(100) * m.run_Su[u] * 4);
  m.run_i = -1;
  m.run_steps = 0;
}

void diff_begin(tts_ctx *c, const float *latents, int Lf, int S, int n_steps, const float *x0) {
  const int32_t l = Lf, sv = S;
  diff_begin_batch(c, 1, &latents, &l, &sv, n_steps, &x0);
}
void diff_step(tts_ctx *c, const float *noise_block) { diff_step_batch(c, &noise_block); }
void diff_end(tts_ctx *c, float *mel) { diff_end_batch(c, &mel); }

void diff_sample(tts_ctx *c, const float *latents, int Lf, int S, int n_steps, const float *noise, float *mel) {
  check_sizes(c, Lf, S);
  const size_t nx = size_t(100) * S;
  diff_begin(c, latents, Lf, S, n_steps, noise);
  for (int i = 0; i < n_steps; ++i) diff_step(c, noise + size_t(i + 1) * nx);
  diff_end(c, mel);
}

}  // namespace tts
