// engine.h -- internal C++ structures behind the C-ABI (include/tortoise_b200.h).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <map>
#include <set>
#include <string>
#include <vector>

#include "../../include/tortoise_b200.h"
#include "common.cuh"

namespace tts {

struct HostTensor {
  std::vector<int> ne;  // ggml order, ne[0] fastest
  size_t offset = 0;    // byte offset of the data in the file
  size_t nelem = 0;
};

// Container parser shared by the three loaders (format: SURVEY.md App. B,
// reference parser main.cpp:811-888).
struct Container {
  std::string path;
  std::map<std::string, HostTensor> tensors;
  std::vector<std::string> order;
  bool open(const std::string &path, std::string &err);
};

struct ArLayer {
  float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
  void *w_qkv, *w_proj, *w_fc, *w_proj2;  // [N][K], f32 or f16
  float *b_qkv, *b_proj, *b_fc, *b_proj2;
  // split-f16 planes for the tensor-core row GEMMs (prefill / latent pass).  Fast mode:
  // hi aliases the f16 GEMV weights, lo is null.  Parity mode: hi/lo pair of the f32 weights.
  __half *qkv_hi, *qkv_lo, *proj_hi, *proj_lo, *fc_hi, *fc_lo, *proj2_hi, *proj2_lo;
};

constexpr size_t kMegaDbgWords = 160 * 128 * 4;  // trace buffer of the persistent decode kernels

struct ArModel {
  bool loaded = false;
  int dtype = 0;
  ArLayer layers[30];
  float *lnf_w, *lnf_b, *lm0_w, *lm0_b, *lm_b;
  void *lm_w;  // [8194][1024]
  float *text_emb, *text_pos, *mel_emb, *mel_pos;
  size_t decode_weight_bytes = 0;
  void *mega_layers = nullptr;   // device MegaLayer[30] (ar_mega2.cuh)
  long long *mega_dbg = nullptr;  // device trace buffer (TTS_MEGA_TRACE=1|2)
  int mega_dbg_mode = 0;
  // (value, tag) exchange buffers of the second-generation persistent step (ar_mega2.cuh)
  uint2 *ll_h = nullptr, *ll_h2 = nullptr, *ll_qkv = nullptr, *ll_m = nullptr, *ll_att = nullptr;
  size_t ll_bytes[5] = {0, 0, 0, 0, 0};
  unsigned int mega_epoch = 0;  // launch counter = tag generation
  // exchange buffers of the 5..16-candidate step (ar_mega4.cuh), laid out for 16 candidates
  uint2 *l4_h = nullptr, *l4_h2 = nullptr, *l4_qkv = nullptr, *l4_m = nullptr, *l4_att = nullptr;
  size_t l4_bytes[5] = {0, 0, 0, 0, 0};
};

struct ArState {
  int Bmax = 0, P = 0;
  int B = 0, T = 0, n_past = 0;
  int n_prefix = 0;  // > 0: the prompt's K/V rows are stored once, in candidate slot 0 (ar_mega4.cuh)
  bool multi = false;  // slots hold DIFFERENT prompts (utterance batching), right-aligned: slot b's rows [0, start[b]) are padding
  int start[16] = {0};
  float *h = nullptr, *q = nullptr, *attn = nullptr, *m = nullptr, *logits = nullptr;
  __half *kc = nullptr, *vc = nullptr;  // [30][Bmax][16][P][64]
  int *d_tokens = nullptr, *d_state = nullptr;
  int *h_tokens = nullptr, *h_state = nullptr;  // pinned
  float *h_logits = nullptr;                    // pinned
  float *d_topv = nullptr, *h_topv = nullptr;   // device top-k (value, index) pairs + overflow flags, and their pinned mirror
  int *d_topi = nullptr, *h_topi = nullptr, *d_topf = nullptr, *h_topf = nullptr;
  // row buffers for prefill / latent pass (grown on demand)
  size_t rows_cap = 0;
  float *H = nullptr, *QKV = nullptr, *Z = nullptr;
  __half *Ahi = nullptr, *Alo = nullptr, *ATThi = nullptr, *ATTlo = nullptr, *Mhi = nullptr;
  int *d_text = nullptr, *d_codes = nullptr, *d_pos = nullptr;
  float *d_voice = nullptr;
  // CUDA graph of one decode step, keyed by B
  cudaGraphExec_t step_graph = nullptr;
  int step_graph_B = 0;
  int step_graph_kernels = 0;
};

struct DiffModel;
struct VocModel;

}  // namespace tts

struct tts_ctx {
  tts_config cfg;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  int64_t launches = 0;
  float last_ms = 0.f;
  double total_ms = 0.0;  // device time of every stage call so far (CUDA events)
  bool use_graph = true;
  bool use_pdl = true;
  bool use_mega = true;   // persistent single-kernel decode step (TTS_NO_MEGA=1 -> per-op graph path)
  bool use_mega_v2 = false;  // TTS_MEGA_V2=1: CUDA-core GEMV phases also for f16 weights (default: tensor-core ar_mega3.cuh)
  tts::ArModel ar;
  tts::ArState ars;
  tts::DiffModel *diff = nullptr;
  tts::VocModel *voc = nullptr;
  float *staging = nullptr;      // pinned host staging for loads
  size_t staging_bytes = 0;
  float *d_scratch = nullptr;    // device scratch for load-time conversion
  size_t d_scratch_bytes = 0;
  // every device / pinned-host allocation of this context (ctx_malloc / ctx_malloc_host): released
  // by tts_free, so creating and destroying engines in one process does not leak
  std::set<void *> dev_allocs, host_allocs;
};

namespace tts {
// ---- context-owned memory ------------------------------------------------------------------
template <typename T>
inline cudaError_t ctx_malloc(tts_ctx *c, T **p, size_t bytes) {
  void *q = nullptr;
  const cudaError_t e = cudaMalloc(&q, bytes ? bytes : 1);
  if (e == cudaSuccess) c->dev_allocs.insert(q);
  *p = static_cast<T *>(q);
  return e;
}
template <typename T>
inline cudaError_t ctx_malloc_host(tts_ctx *c, T **p, size_t bytes) {
  void *q = nullptr;
  const cudaError_t e = cudaMallocHost(&q, bytes ? bytes : 1);
  if (e == cudaSuccess) c->host_allocs.insert(q);
  *p = static_cast<T *>(q);
  return e;
}
inline void ctx_free(tts_ctx *c, const void *p) {
  if (!p) return;
  c->dev_allocs.erase(const_cast<void *>(p));
  cudaFree(const_cast<void *>(p));
}
inline void ctx_free_host(tts_ctx *c, const void *p) {
  if (!p) return;
  c->host_allocs.erase(const_cast<void *>(p));
  cudaFreeHost(const_cast<void *>(p));
}
inline void ctx_free_all(tts_ctx *c) {
  for (void *p : c->dev_allocs) cudaFree(p);
  for (void *p : c->host_allocs) cudaFreeHost(p);
  c->dev_allocs.clear();
  c->host_allocs.clear();
}
// implemented in ar.cu
void ar_load(tts_ctx *c, const char *path);
void ar_prefill(tts_ctx *c, const int32_t *text, int T, const float *voice, int B, float *logits_out);
void ar_prefill_multi(tts_ctx *c, int U, const int32_t *const *text, const int32_t *T, const float *voice, float *logits_out);
void ar_step(tts_ctx *c, const int32_t *tokens, int pos_id, float *logits_out, bool sync_out);
void ar_step_topk(tts_ctx *c, const int32_t *tokens, int pos_id, float *vals_out, int32_t *idx_out, int32_t *flags_out);
void ar_logits(tts_ctx *c, float *logits_out);
void ar_latents(tts_ctx *c, const int32_t *text, int T, const float *voice, const int32_t *codes, int B,
                int n_keep, float *out);
void ar_bench_gemv(tts_ctx *c, int op, int B, int iters, float *ms, double *bytes);
void ar_bench_step(tts_ctx *c, int iters, float *ms, double *bytes);
void ar_free(tts_ctx *c);
// implemented in diffusion.cu / vocoder.cu
void diff_load(tts_ctx *c, const char *path);
void diff_eps(tts_ctx *c, const float *latents, int L, const float *x, int S, int timestep, int cond_free,
              float *out);
void diff_sample(tts_ctx *c, const float *latents, int L, int S, int n_steps, const float *noise, float *mel);
void diff_begin(tts_ctx *c, const float *latents, int L, int S, int n_steps, const float *x0);
void diff_step(tts_ctx *c, const float *noise_block);
void diff_end(tts_ctx *c, float *mel);
void diff_begin_batch(tts_ctx *c, int U, const float *const *latents, const int32_t *Lf, const int32_t *S, int n_steps,
                      const float *const *x0);
void diff_step_batch(tts_ctx *c, const float *const *noise_blocks);
void diff_end_batch(tts_ctx *c, float *const *mel);
void diff_free(tts_ctx *c);
void diff_bench_conv3(tts_ctx *c, int S, int iters, float *ms, double *flop);
void voc_load(tts_ctx *c, const char *path);
void voc_run(tts_ctx *c, const float *mel, int S, const float *noise, float *audio);
void voc_free(tts_ctx *c);
// shared helpers (loader.cpp)
float *upload_f32(tts_ctx *c, const Container &ct, const std::string &name, const std::vector<int> &expect_ne);
void read_tensor_to_staging(tts_ctx *c, const Container &ct, const std::string &name, size_t *nelem);
}  // namespace tts
