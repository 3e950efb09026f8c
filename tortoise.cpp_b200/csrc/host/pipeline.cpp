// pipeline.cpp -- the three stage drivers above the device C-ABI, mirroring the reference's
// autoregressive() (main.cpp:5042-5367), diffusion() (5614-6042) and vocoder() (6044-6127):
// same RNG draw order, same stop rule, same padding / trimming, same S = L*4*24000/22050.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/tortoise_b200.h"
#include "../../../include/tortoise_host.h"
#include "rng.h"

namespace tts_host {
int sample_one(Rng &r, const float *logits_row, const int32_t *prev, int n_prev, float *logprob);
int sample_sparse_one(Rng &r, const float *vals, const int32_t *idx, int n, const int32_t *prev, int n_prev, float *logprob);
void apply_padding(std::vector<int32_t> &vec);
int trim_count(const int32_t *codes500);
}  // namespace tts_host

extern "C" {

int tts_host_autoregressive(struct tts_ctx *ctx, tts_rng *rng, const int32_t *tokens, int T, const float *voice,
                            int B, const tts_ar_options *opt_in, int32_t *codes_out, float *latents_out,
                            int32_t *n_latents, float *score_out, int32_t *steps_out) {
  if (!ctx || !rng || !tokens || !voice || !codes_out || !n_latents || B < 1) return TTS_EINVAL;
  tts_ar_options opt{};
  if (opt_in) opt = *opt_in;
  if (!latents_out && !opt.skip_latents) return TTS_EINVAL;
  const int V = TTS_MEL_VOCAB, STOP = TTS_MEL_STOP;
  std::vector<float> logits(size_t(B) * V);
  int rc = tts_ar_prefill(ctx, tokens, T, voice, B, logits.data());
  if (rc != TTS_OK) return rc;
  // "mel_transformer_inputs_vector": [1]*(T+1) + [8192] per candidate (main.cpp:5095-5105)
  int n_prev = T + 2;
  std::vector<int32_t> prev(size_t(B) * n_prev, 1);
  for (int b = 0; b < B; ++b) prev[size_t(b) * n_prev + n_prev - 1] = TTS_MEL_START;
  std::vector<std::vector<int32_t>> seqs(B);
  std::vector<int32_t> samples(B);
  std::vector<double> lp_sum(B, 0.0);
  std::vector<int> lp_n(B, 0);
  std::vector<char> done(B, 0);
  // device-side pre-selection (steps after the prefill): TTS_AR_TOPK (value, index) pairs per candidate
  // cross PCIe instead of 8194 floats; the full row is fetched only when the sparse sampler asks for it
  std::vector<float> top_v(size_t(B) * TTS_AR_TOPK);
  std::vector<int32_t> top_i(size_t(B) * TTS_AR_TOPK), top_f(B, 0);
  bool sparse = false, have_full = true;
  int i = 0;
  for (;;) {
    for (int b = 0; b < B; ++b) {
      float lp = 0.f;
      int smp = -1;
      const bool suppress_stop = opt.forced_codes > 0 && i < opt.forced_codes;  // bench mode: no early stop
      if (sparse && !top_f[b]) {
        // (a suppressed stop token leaves the set: what remains are the largest entries of the modified row)
        float tv[TTS_AR_TOPK];
        int32_t ti[TTS_AR_TOPK];
        int n = 0;
        for (int k = 0; k < TTS_AR_TOPK; ++k) {
          const int32_t id = top_i[size_t(b) * TTS_AR_TOPK + k];
          if (suppress_stop && id == STOP) continue;
          tv[n] = top_v[size_t(b) * TTS_AR_TOPK + k];
          ti[n++] = id;
        }
        smp = tts_host::sample_sparse_one(rng->r, tv, ti, n, prev.data() + size_t(b) * n_prev, n_prev, &lp);
      }
      if (smp < 0) {
        if (!have_full) {
          rc = tts_ar_logits(ctx, logits.data());
          if (rc != TTS_OK) return rc;
          have_full = true;
        }
        float *row = logits.data() + size_t(b) * V;
        if (suppress_stop) row[STOP] = -1e30f;
        smp = tts_host::sample_one(rng->r, row, prev.data() + size_t(b) * n_prev, n_prev, &lp);
      }
      samples[b] = smp;
      if (opt.forced_codes > 0 && i >= opt.forced_codes) samples[b] = STOP;
      if (!done[b] && std::isfinite(lp)) {
        lp_sum[b] += lp;
        lp_n[b] += 1;
      }
    }
    int stops = 0;
    for (int b = 0; b < B; ++b) {
      if (!(seqs[b].size() > 0 && seqs[b].back() == STOP)) seqs[b].push_back(samples[b]);
      if (samples[b] == STOP) {
        stops += 1;
        done[b] = 1;
      }
    }
    bool finished = stops == B;  // reference rule: all candidates emit 8193 in the SAME step
    if (opt.per_candidate_stop) {
      finished = true;
      for (int b = 0; b < B; ++b) finished = finished && done[b];
    }
    prev.assign(samples.begin(), samples.end());
    n_prev = 1;
    if (finished) break;
    for (int b = 0; b < B; ++b)
      if (seqs[b].size() > 500) return TTS_ELIMIT;  // apply_padding asserts <= 500 (main.cpp:4517)
    if (opt.max_steps > 0 && i + 1 >= opt.max_steps) return TTS_ELIMIT;
    if (opt.full_logits) {
      rc = tts_ar_step(ctx, samples.data(), i + 2, logits.data());  // fixed_position = i + 2 (main.cpp:5227)
    } else {
      rc = tts_ar_step_topk(ctx, samples.data(), i + 2, top_v.data(), top_i.data(), top_f.data());
      sparse = true;
      have_full = false;
    }
    if (rc != TTS_OK) return rc;
    i += 1;
  }
  if (steps_out) *steps_out = i + 1;
  std::vector<int32_t> codes502(size_t(B) * 502);
  int n_keep = 1;
  for (int b = 0; b < B; ++b) {
    tts_host::apply_padding(seqs[b]);
    if (seqs[b].size() != 502) return TTS_ELIMIT;
    memcpy(codes502.data() + size_t(b) * 502, seqs[b].data(), 502 * 4);
    memcpy(codes_out + size_t(b) * 500, seqs[b].data() + 1, 500 * 4);  // trim_latents drops first/last
    n_latents[b] = tts_host::trim_count(codes_out + size_t(b) * 500);
    if (n_latents[b] > n_keep) n_keep = n_latents[b];
    if (score_out) score_out[b] = lp_n[b] ? float(lp_sum[b] / lp_n[b]) : 0.f;
  }
  // candidate selection first, latent pass for the winner only (tts_host_latents): multi-candidate / multi-GPU drivers
  if (opt.skip_latents) return TTS_OK;
  rc = tts_ar_latents(ctx, tokens, T, voice, codes502.data(), B, n_keep, latents_out);
  if (rc != TTS_OK) return rc;
  for (int b = 0; b < B; ++b)  // zero the rows trim_latents would not return
    memset(latents_out + (size_t(b) * 500 + n_latents[b]) * 1024, 0, size_t(500 - n_latents[b]) * 1024 * 4);
  return TTS_OK;
}

// Utterance-batched decode loop (BASELINE configs[4]; no reference counterpart: the reference decodes one prompt per
// run): U different prompts share every batched decode launch, each with its own RNG stream, repetition-penalty
// window and stop state -- utterance u's codes are what tts_host_autoregressive(B = 1) would sample for it from the
// same logits.  The latent pass is left to the caller (tts_host_latents per utterance).
int tts_host_autoregressive_multi(struct tts_ctx *ctx, tts_rng *const *rngs, int U, const int32_t *const *tokens, const int32_t *T,
                                  const float *voice, const int32_t *forced_codes, int max_steps, int32_t *codes_out,
                                  int32_t *n_latents, int32_t *steps_out) {
  if (!ctx || !rngs || !tokens || !T || !voice || !codes_out || !n_latents || U < 1 || U > 16) return TTS_EINVAL;
  const int V = TTS_MEL_VOCAB, STOP = TTS_MEL_STOP;
  std::vector<float> logits(size_t(U) * V);
  int rc = tts_ar_prefill_multi(ctx, U, tokens, T, voice, logits.data());
  if (rc != TTS_OK) return rc;
  std::vector<std::vector<int32_t>> prev(U), seqs(U);
  for (int u = 0; u < U; ++u) {
    if (!rngs[u]) return TTS_EINVAL;
    prev[u].assign(size_t(T[u]) + 2, 1);  // [1] * (T + 1) + [8192] (main.cpp:5095-5105)
    prev[u].back() = TTS_MEL_START;
  }
  std::vector<int32_t> samples(U), steps(U, 0);
  std::vector<char> done(U, 0);
  std::vector<float> top_v(size_t(U) * TTS_AR_TOPK);
  std::vector<int32_t> top_i(size_t(U) * TTS_AR_TOPK), top_f(U, 0);
  bool sparse = false, have_full = true;
  int i = 0;
  for (;;) {
    for (int u = 0; u < U; ++u) {
      if (done[u]) {
        samples[u] = STOP;
        continue;
      }
      const int forced = forced_codes ? forced_codes[u] : 0;
      const bool suppress_stop = forced > 0 && i < forced;
      float lp = 0.f;
      int smp = -1;
      if (sparse && !top_f[u]) {
        float tv[TTS_AR_TOPK];
        int32_t ti[TTS_AR_TOPK];
        int n = 0;
        for (int k = 0; k < TTS_AR_TOPK; ++k) {
          const int32_t id = top_i[size_t(u) * TTS_AR_TOPK + k];
          if (suppress_stop && id == STOP) continue;
          tv[n] = top_v[size_t(u) * TTS_AR_TOPK + k];
          ti[n++] = id;
        }
        smp = tts_host::sample_sparse_one(rngs[u]->r, tv, ti, n, prev[u].data(), int(prev[u].size()), &lp);
      }
      if (smp < 0) {
        if (!have_full) {
          rc = tts_ar_logits(ctx, logits.data());
          if (rc != TTS_OK) return rc;
          have_full = true;
        }
        float *row = logits.data() + size_t(u) * V;
        if (suppress_stop) row[STOP] = -1e30f;
        smp = tts_host::sample_one(rngs[u]->r, row, prev[u].data(), int(prev[u].size()), &lp);
      }
      if (forced > 0 && i >= forced) smp = STOP;
      samples[u] = smp;
      seqs[u].push_back(smp);
      steps[u] = i + 1;
      if (smp == STOP) done[u] = 1;
      prev[u].assign(1, smp);
      if (seqs[u].size() > 500) return TTS_ELIMIT;  // apply_padding asserts <= 500 (main.cpp:4517)
    }
    bool finished = true;
    for (int u = 0; u < U; ++u) finished = finished && done[u];
    if (finished) break;
    if (max_steps > 0 && i + 1 >= max_steps) return TTS_ELIMIT;
    rc = tts_ar_step_topk(ctx, samples.data(), i + 2, top_v.data(), top_i.data(), top_f.data());  // fixed_position = i + 2
    if (rc != TTS_OK) return rc;
    sparse = true;
    have_full = false;
    i += 1;
  }
  for (int u = 0; u < U; ++u) {
    tts_host::apply_padding(seqs[u]);
    if (seqs[u].size() != 502) return TTS_ELIMIT;
    memcpy(codes_out + size_t(u) * 500, seqs[u].data() + 1, 500 * 4);  // trim_latents drops first / last
    n_latents[u] = tts_host::trim_count(codes_out + size_t(u) * 500);
    if (steps_out) steps_out[u] = steps[u];
  }
  return TTS_OK;
}

int tts_host_latents(struct tts_ctx *ctx, const int32_t *tokens, int T, const float *voice, const int32_t *codes500,
                     float *latents_out, int32_t *n_latents) {
  if (!ctx || !tokens || !voice || !codes500 || !latents_out || !n_latents) return TTS_EINVAL;
  std::vector<int32_t> codes502(502);
  codes502[0] = TTS_MEL_START;
  memcpy(codes502.data() + 1, codes500, 500 * 4);
  codes502[501] = TTS_MEL_STOP;
  const int n = tts_host::trim_count(codes500);
  *n_latents = n;
  const int rc = tts_ar_latents(ctx, tokens, T, voice, codes502.data(), 1, n, latents_out);
  if (rc != TTS_OK) return rc;
  memset(latents_out + size_t(n) * 1024, 0, size_t(500 - n) * 1024 * 4);
  return TTS_OK;
}

int tts_host_diffusion(struct tts_ctx *ctx, tts_rng *rng, const float *latents, int L, int n_steps, float *mel_out,
                       int32_t *S_out) {
  if (!ctx || !rng || !latents || !mel_out || L < 1 || n_steps < 1) return TTS_EINVAL;
  const int S = L * 4 * 24000 / 22050;  // main.cpp:5616-5617 (integer arithmetic)
  if (S_out) *S_out = S;
  const size_t nx = size_t(100) * S;
  // RNG order: initial x (main.cpp:5638), then one block per step (main.cpp:6020), drawn every
  // step including the last.  Nothing else consumes the generator in between, so drawing the
  // blocks up front is the same stream.
  // Nothing else consumes the generator in between, so the stream equals the reference's; the
  // blocks are drawn one step ahead of the GPU (tts_diffusion_step is asynchronous).
  std::vector<float> blk(nx);
  for (size_t i = 0; i < nx; ++i) blk[i] = rng->r.normal(rng->r.generator);
  int rc = tts_diffusion_begin(ctx, latents, L, S, n_steps, blk.data());
  if (rc != TTS_OK) return rc;
  for (int s = 0; s < n_steps; ++s) {
    for (size_t i = 0; i < nx; ++i) blk[i] = rng->r.normal(rng->r.generator);
    rc = tts_diffusion_step(ctx, blk.data());
    if (rc != TTS_OK) return rc;
  }
  return tts_diffusion_end(ctx, mel_out);
}

int tts_host_diffusion_batch(struct tts_ctx *ctx, tts_rng *const *rngs, int U, const float *const *latents, const int32_t *L,
                             int n_steps, float *const *mel_out, int32_t *S_out) {
  if (!ctx || !rngs || !latents || !L || !mel_out || U < 1 || n_steps < 1) return TTS_EINVAL;
  std::vector<int32_t> S(U);
  std::vector<std::vector<float>> blk(U);
  std::vector<const float *> ptr(U);
  for (int u = 0; u < U; ++u) {
    if (!rngs[u] || !latents[u] || !mel_out[u] || L[u] < 1) return TTS_EINVAL;
    S[u] = L[u] * 4 * 24000 / 22050;  // main.cpp:5616-5617 (integer arithmetic)
    if (S_out) S_out[u] = S[u];
    blk[u].resize(size_t(100) * S[u]);
    for (float &v : blk[u]) v = rngs[u]->r.normal(rngs[u]->r.generator);  // initial x (main.cpp:5638)
    ptr[u] = blk[u].data();
  }
  int rc = tts_diffusion_begin_batch(ctx, U, latents, L, S.data(), n_steps, ptr.data());
  if (rc != TTS_OK) return rc;
  for (int s = 0; s < n_steps; ++s) {
    for (int u = 0; u < U; ++u)
      for (float &v : blk[u]) v = rngs[u]->r.normal(rngs[u]->r.generator);  // one block per step, last included (main.cpp:6020)
    rc = tts_diffusion_step_batch(ctx, ptr.data());
    if (rc != TTS_OK) return rc;
  }
  return tts_diffusion_end_batch(ctx, mel_out);
}

int tts_host_vocoder(struct tts_ctx *ctx, tts_rng *rng, const float *mel, int S, float *audio_out) {
  if (!ctx || !rng || !mel || !audio_out || S < 1) return TTS_EINVAL;
  std::vector<float> noise(size_t(S + 10) * 64);  // main.cpp:6057
  for (size_t i = 0; i < noise.size(); ++i) noise[i] = rng->r.normal(rng->r.generator);
  return tts_vocoder(ctx, mel, S, noise.data(), audio_out);
}

}  // extern "C"
