// tts_api.cu -- the C-ABI (include/tortoise_b200.h): argument checks, error plumbing.
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "engine.h"
#include "../../include/tortoise_b200_bench.h"

static std::string g_last_error;
static std::mutex g_err_mu;

static int fail(tts_ctx *c, int code, const std::string &msg) {
  if (c) c->err = msg;
  std::lock_guard<std::mutex> lk(g_err_mu);
  g_last_error = msg;
  return code;
}

#define TTS_API_BODY(ctx, ...)                                          \
  try {                                                                 \
    if (!(ctx)) return fail(nullptr, TTS_EINVAL, "null context");       \
    cudaError_t _se = cudaSetDevice((ctx)->cfg.device);                 \
    if (_se != cudaSuccess) return fail(ctx, TTS_ECUDA, cudaGetErrorString(_se)); \
    __VA_ARGS__;                                                        \
    return TTS_OK;                                                      \
  } catch (const tts::CudaError &e) {                                   \
    cudaGetLastError();                                                 \
    return fail(ctx, TTS_ECUDA, e.msg);                                 \
  } catch (const tts::ArgError &e) {                                    \
    return fail(ctx, e.code, e.msg);                                    \
  } catch (const std::exception &e) {                                   \
    return fail(ctx, TTS_EINVAL, e.what());                             \
  }

extern "C" {

int tts_version(void) { return 100; }

const char *tts_last_error(const tts_ctx *ctx) {
  if (ctx) return ctx->err.c_str();
  return g_last_error.c_str();
}

int tts_init(const tts_config *cfg, tts_ctx **out) {
  if (!cfg || !out) return fail(nullptr, TTS_EINVAL, "null argument");
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(nullptr, TTS_ENODEV, "no CUDA device: libtortoise_b200 has no CPU fallback");
  }
  if (cfg->device < 0 || cfg->device >= n) return fail(nullptr, TTS_ENODEV, "bad device ordinal");
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, cfg->device) != cudaSuccess) return fail(nullptr, TTS_ECUDA, "cudaGetDeviceProperties failed");
  if (p.major != 10) {
    char b[256];
    snprintf(b, sizeof b, "device %d is sm_%d%d; this library ships sm_100a code only", cfg->device, p.major, p.minor);
    return fail(nullptr, TTS_ENODEV, b);
  }
  if (cfg->dtype != TTS_DTYPE_F32 && cfg->dtype != TTS_DTYPE_F16) return fail(nullptr, TTS_EINVAL, "bad dtype");
  tts_ctx *c = new tts_ctx();
  c->cfg = *cfg;
  if (c->cfg.max_batch <= 0) c->cfg.max_batch = 4;
  if (c->cfg.max_positions <= 0) c->cfg.max_positions = 404;
  if (c->cfg.max_batch > 64) { delete c; return fail(nullptr, TTS_ELIMIT, "max_batch > 64"); }
  if (c->cfg.max_positions > 2048) { delete c; return fail(nullptr, TTS_ELIMIT, "max_positions > 2048"); }
  c->num_sms = p.multiProcessorCount;
  const char *eg = getenv("TTS_NO_GRAPH");
  c->use_graph = !(eg && eg[0] == '1');
  const char *ep = getenv("TTS_NO_PDL");
  c->use_pdl = !(ep && ep[0] == '1');
  const char *em = getenv("TTS_NO_MEGA");
  c->use_mega = !(em && em[0] == '1');
  { const char *e2 = getenv("TTS_MEGA_V2"); c->use_mega_v2 = e2 && e2[0] == '1'; }
  try {
    TTS_CUDA_TRY(cudaSetDevice(cfg->device));
    TTS_CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    TTS_CUDA_TRY(cudaEventCreate(&c->ev0));
    TTS_CUDA_TRY(cudaEventCreate(&c->ev1));
  } catch (const tts::CudaError &e) {
    delete c;
    return fail(nullptr, TTS_ECUDA, e.msg);
  }
  *out = c;
  return TTS_OK;
}

void tts_free(tts_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->cfg.device);
  cudaStreamSynchronize(c->stream);
  tts::ar_free(c);
  tts::diff_free(c);
  tts::voc_free(c);
  tts::ctx_free_all(c);  // every device / pinned allocation made through ctx_malloc*
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  cudaStreamDestroy(c->stream);
  delete c;
}

int tts_load_ar(tts_ctx *c, const char *path) { TTS_API_BODY(c, if (!path) throw tts::ArgError("null path"); tts::ar_load(c, path)) }
int tts_load_diffusion(tts_ctx *c, const char *path) { TTS_API_BODY(c, if (!path) throw tts::ArgError("null path"); tts::diff_load(c, path)) }
int tts_load_vocoder(tts_ctx *c, const char *path) { TTS_API_BODY(c, if (!path) throw tts::ArgError("null path"); tts::voc_load(c, path)) }

int tts_ar_prefill(tts_ctx *c, const int32_t *text, int32_t T, const float *voice, int32_t B, float *logits) {
  TTS_API_BODY(c, if (!text || !voice) throw tts::ArgError("null argument"); tts::ar_prefill(c, text, T, voice, B, logits))
}
int tts_ar_prefill_multi(tts_ctx *c, int32_t U, const int32_t *const *text, const int32_t *T, const float *voice, float *logits) {
  TTS_API_BODY(c, if (!text || !T || !voice) throw tts::ArgError("null argument");
               for (int u = 0; u < U; ++u) if (!text[u]) throw tts::ArgError("null argument");
               tts::ar_prefill_multi(c, U, text, T, voice, logits))
}
int tts_ar_step(tts_ctx *c, const int32_t *tokens, int32_t pos_id, float *logits) {
  TTS_API_BODY(c, if (!tokens) throw tts::ArgError("null argument"); tts::ar_step(c, tokens, pos_id, logits, true))
}
int tts_ar_step_dev(tts_ctx *c, const int32_t *tokens, int32_t pos_id, const float **logits_dev) {
  TTS_API_BODY(c, if (!tokens) throw tts::ArgError("null argument"); tts::ar_step(c, tokens, pos_id, nullptr, false);
               if (logits_dev) *logits_dev = c->ars.logits)
}
int tts_ar_step_topk(tts_ctx *c, const int32_t *tokens, int32_t pos_id, float *vals, int32_t *idx, int32_t *flags) {
  TTS_API_BODY(c, if (!tokens || !vals || !idx || !flags) throw tts::ArgError("null argument");
               tts::ar_step_topk(c, tokens, pos_id, vals, idx, flags))
}
int tts_ar_logits(tts_ctx *c, float *logits) {
  TTS_API_BODY(c, if (!logits) throw tts::ArgError("null argument"); tts::ar_logits(c, logits))
}
int tts_ar_latents(tts_ctx *c, const int32_t *text, int32_t T, const float *voice, const int32_t *codes, int32_t B,
                   int32_t n_keep, float *out) {
  TTS_API_BODY(c, if (!text || !voice || !codes || !out) throw tts::ArgError("null argument");
               tts::ar_latents(c, text, T, voice, codes, B, n_keep, out))
}
int tts_diffusion_eps(tts_ctx *c, const float *latents, int32_t L, const float *x, int32_t S, int32_t timestep,
                      int32_t cond_free, float *out) {
  TTS_API_BODY(c, if (!latents || !x || !out) throw tts::ArgError("null argument");
               tts::diff_eps(c, latents, L, x, S, timestep, cond_free, out))
}
int tts_diffusion_sample(tts_ctx *c, const float *latents, int32_t L, int32_t S, int32_t n_steps, const float *noise,
                         float *mel) {
  TTS_API_BODY(c, if (!latents || !noise || !mel) throw tts::ArgError("null argument");
               tts::diff_sample(c, latents, L, S, n_steps, noise, mel))
}
int tts_diffusion_begin(tts_ctx *c, const float *latents, int32_t L, int32_t S, int32_t n_steps, const float *x0) {
  TTS_API_BODY(c, if (!latents || !x0) throw tts::ArgError("null argument"); tts::diff_begin(c, latents, L, S, n_steps, x0))
}
int tts_diffusion_step(tts_ctx *c, const float *noise_block) {
  TTS_API_BODY(c, if (!noise_block) throw tts::ArgError("null argument"); tts::diff_step(c, noise_block))
}
int tts_diffusion_end(tts_ctx *c, float *mel) {
  TTS_API_BODY(c, if (!mel) throw tts::ArgError("null argument"); tts::diff_end(c, mel))
}
int tts_diffusion_begin_batch(tts_ctx *c, int32_t U, const float *const *latents, const int32_t *L, const int32_t *S,
                              int32_t n_steps, const float *const *x0) {
  TTS_API_BODY(c, if (!latents || !L || !S || !x0 || U < 1) throw tts::ArgError("bad argument");
               for (int u = 0; u < U; ++u) if (!latents[u] || !x0[u]) throw tts::ArgError("null argument");
               tts::diff_begin_batch(c, U, latents, L, S, n_steps, x0))
}
int tts_diffusion_step_batch(tts_ctx *c, const float *const *noise_blocks) {
  TTS_API_BODY(c, if (!noise_blocks) throw tts::ArgError("null argument"); tts::diff_step_batch(c, noise_blocks))
}
int tts_diffusion_end_batch(tts_ctx *c, float *const *mel) {
  TTS_API_BODY(c, if (!mel) throw tts::ArgError("null argument"); tts::diff_end_batch(c, mel))
}
int tts_vocoder(tts_ctx *c, const float *mel, int32_t S, const float *noise, float *audio) {
  TTS_API_BODY(c, if (!mel || !noise || !audio) throw tts::ArgError("null argument"); tts::voc_run(c, mel, S, noise, audio))
}
int tts_sync(tts_ctx *c) { TTS_API_BODY(c, TTS_CUDA_TRY(cudaStreamSynchronize(c->stream))) }

int64_t tts_launch_count(const tts_ctx *c) { return c ? c->launches : 0; }
float tts_last_stage_ms(const tts_ctx *c) { return c ? c->last_ms : 0.f; }
double tts_device_ms_total(const tts_ctx *c) { return c ? c->total_ms : 0.0; }
int tts_bench_gemv(tts_ctx *c, int32_t op, int32_t B, int32_t iters, float *ms, double *bytes) {
  TTS_API_BODY(c, if (!ms || !bytes || iters < 1) throw tts::ArgError("bad argument"); tts::ar_bench_gemv(c, op, B, iters, ms, bytes))
}

int tts_bench_conv3(tts_ctx *c, int32_t S, int32_t iters, float *ms, double *flop) {
  TTS_API_BODY(c, if (!ms || !flop) throw tts::ArgError("bad argument"); tts::diff_bench_conv3(c, S, iters, ms, flop))
}

int tts_bench_decode_step(tts_ctx *c, int32_t iters, float *ms, double *bytes) {
  TTS_API_BODY(c, if (!ms || !bytes) throw tts::ArgError("bad argument"); tts::ar_bench_step(c, iters, ms, bytes))
}

}  // extern "C"
