/* tortoise_host.h -- C-ABI of the HOST side of the drop-in (pure C++, no CUDA):
 * tokenizer, RNG, logits post-processing + sampling, padding / trimming, WAV writer, and
 * the three stage drivers that sit above the device C-ABI (tortoise_b200.h) exactly where
 * the reference's autoregressive() / diffusion() / vocoder() sit above ggml.
 *
 * The host-only entry points live in libtortoise_host.so (loadable without a GPU; used by
 * the `-m "not gpu"` tests); libtortoise_b200.so exports them too, plus the stage drivers.
 * Bit-exact with the reference for all integer results on the same seed (it uses the same
 * libstdc++ std::mt19937 / uniform_real_distribution<float> / normal_distribution<double>
 * objects in the same draw order, SURVEY A-8/A-10).
 */
#ifndef TORTOISE_HOST_H
#define TORTOISE_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

struct tts_ctx; /* tortoise_b200.h */

/* ---- RNG: the reference's three globals (main.cpp:35-50) as one object ---------------- */
typedef struct tts_rng tts_rng;
tts_rng *tts_rng_create(uint32_t seed);   /* std::mt19937 generator(seed)                   */
void tts_rng_seed(tts_rng *r, uint32_t seed); /* generator.seed(seed)  (main.cpp:6546)      */
void tts_rng_free(tts_rng *r);
float tts_rng_uniform(tts_rng *r);        /* distribution(generator)                        */
void tts_rng_normal(tts_rng *r, float *out, int64_t n); /* sample_normal_noise (main.cpp:4695) */

/* ---- tokenizer: gpt_vocab_init + replaceAll + gpt_tokenize + [255] ... [0]
 *      (main.cpp:6550-6567, common.cpp:166-339).  Returns the token count (or <0), writes
 *      at most cap ids. ---------------------------------------------------------------- */
int tts_host_tokenize(const char *tokenizer_json_path, const char *message, int32_t *out, int cap);
int tts_host_vocab_size(const char *tokenizer_json_path);

/* ---- optional text front-end (not in the reference, which accepts only [a-z .,!?'-],
 *      README.md:32; OFF by default, the CLI turns it on with --normalize).
 *      normalize: lower-case, numbers and & % + = @ $ spelled out, other characters -> space.
 *      Returns the length written (NUL-terminated) or -(bytes needed) when cap is too small.
 *      split: sentence-boundary chunks of at most max_chars characters; writes (start, end)
 *      byte offsets into spans_out[2 * cap_spans], returns the chunk count (or <0). ------- */
int tts_host_normalize_text(const char *in, char *out, int cap);
int tts_host_split_text(const char *in, int max_chars, int32_t *spans_out, int cap_spans);

/* ---- process_logits_and_sample (main.cpp:4753-4806): repetition penalty 2.0 over the
 *      previous inputs prev[B][n_prev], temperature 0.8, top-k 50, top-p (cum <= 0.2 cut),
 *      softmax, multinomial (two uniforms per candidate, second used).  logits [B][8194] is
 *      not modified.  Optional logprob_out[B]: log-probability of the sampled token under
 *      the post-processed distribution (extension used for candidate selection). -------- */
int tts_host_sample(tts_rng *r, const float *logits, const int32_t *prev, int n_prev, int B, int32_t *samples_out,
                    float *logprob_out);
/* The same sampler for ONE candidate fed with the device-side pre-selection of tts_ar_step_topk:
 * n (value, index) pairs = the n largest raw logits of the row.  Returns 0 and the sample when the
 * result is provably the one tts_host_sample gives on the full row (bit-exact, same RNG draws);
 * returns 1 WITHOUT touching the generator when it cannot show that (threshold too close to the
 * cut, ties among survivors): the caller then fetches the row (tts_ar_logits) and uses
 * tts_host_sample.  <0: bad argument. */
int tts_host_sample_sparse(tts_rng *r, const float *vals, const int32_t *idx, int n, const int32_t *prev, int n_prev,
                           int32_t *sample_out, float *logprob_out);
/* literal (slow) restatement of the same function; used to cross-check the fast path */
int tts_host_sample_reference_order(tts_rng *r, const float *logits, const int32_t *prev, int n_prev, int B,
                                    int32_t *samples_out);

/* apply_padding (main.cpp:4510-4532): seq (n <= 500 codes) -> out[502] */
int tts_host_apply_padding(const int32_t *seq, int n, int32_t *out502);
/* trim_latents rule (main.cpp:4894-4911): number of leading frames kept for codes[500] */
int tts_host_trim_count(const int32_t *codes500);

/* writeWav (main.cpp:4821-4868): float32 mono */
int tts_host_write_wav(const char *path, const float *data, int64_t n, int sample_rate);

/* DDPM helpers exposed for tests (host_math.h) */
int tts_host_timestep_map(int n_steps, int32_t *out);
void tts_host_timestep_embedding(int t, float *out1024);
void tts_host_relative_position_buckets(int n, int32_t *out_nxn);
/* out: n_steps rows of {cfk, sqrt_recip, sqrt_recipm1, coef1, coef2, min_log, max_log, last, timestep} */
int tts_host_ddpm_schedule(int n_steps, float *out_nx9);

/* ---- stage drivers (libtortoise_b200.so only) ----------------------------------------- */
typedef struct tts_ar_options {
  int32_t max_steps;        /* 0 = reference behaviour (no cap; KV limit is an error)        */
  int32_t forced_codes;     /* >0: bench mode -- stop token suppressed until this many codes
                               were sampled, then forced (length decoupled from the sampler)  */
  int32_t per_candidate_stop; /* 0 = reference rule: run until ALL candidates emit 8193 in
                               the same step (main.cpp:5206-5222); 1 = each stops on its own   */
  int32_t full_logits;      /* 1 = copy all 8194 logits per candidate to the host every step like the
                               reference (main.cpp:4767); 0 = device-side top-64 pre-selection
                               (tts_ar_step_topk + tts_host_sample_sparse; identical samples)   */
  int32_t skip_latents;     /* 1 = stop after the codes / scores (latents_out may be NULL): the caller
                               selects a candidate -- possibly across GPUs, tts_gather_select -- and
                               runs the latent pass for the winner only (tts_host_latents)       */
  int32_t reserved[3];
} tts_ar_options;

/* autoregressive() (main.cpp:5042-5367).  codes_out [B][500] (after apply_padding, without
 * the leading 8192 / trailing 8193, i.e. what trim_latents sees), latents_out [B][500][1024]
 * (rows >= n_latents[b] zero), n_latents[B], score_out[B] = mean log-prob of sampled codes
 * (may be NULL), steps_out = decode iterations.  Returns 0 or a TTS_E* code. */
int tts_host_autoregressive(struct tts_ctx *ctx, tts_rng *r, const int32_t *tokens, int T, const float *voice_1024,
                            int B, const tts_ar_options *opt, int32_t *codes_out, float *latents_out,
                            int32_t *n_latents, float *score_out, int32_t *steps_out);

/* Utterance-batched decode loop (BASELINE.json configs[4]; the reference decodes one prompt per run,
 * main.cpp:5042): U <= 16 DIFFERENT prompts ride on one batched decode launch per step (tts_ar_prefill_multi +
 * tts_ar_step_topk).  Utterance u samples from its own rngs[u] with its own repetition-penalty window and stop state
 * (forced_codes[u] > 0: exactly that many codes, the stop token suppressed before -- bench mode); it stops being
 * extended once it emits 8193.  codes_out [U][500] / n_latents [U] as tts_host_autoregressive; steps_out [U] =
 * sampled codes per utterance (may be NULL).  The latent pass is the caller's: tts_host_latents per utterance. */
int tts_host_autoregressive_multi(struct tts_ctx *ctx, tts_rng *const *rngs, int32_t U, const int32_t *const *tokens,
                                  const int32_t *T, const float *voice_1024, const int32_t *forced_codes, int32_t max_steps,
                                  int32_t *codes_out, int32_t *n_latents, int32_t *steps_out);

/* The latent pass of autoregressive() (main.cpp:5280-5352) + trim_latents (main.cpp:4873-4915) for ONE
 * candidate's codes500 (as returned in codes_out): latents_out [500][1024], rows >= *n_latents zero. */
int tts_host_latents(struct tts_ctx *ctx, const int32_t *tokens, int T, const float *voice_1024,
                     const int32_t *codes500, float *latents_out, int32_t *n_latents);

/* diffusion() (main.cpp:5614-6042): latents [L][1024] -> mel [100][S], S = L*4*24000/22050
 * (integer arithmetic).  mel_out must hold 100*S floats; *S_out receives S. */
int tts_host_diffusion(struct tts_ctx *ctx, tts_rng *r, const float *latents, int L, int n_steps, float *mel_out,
                       int32_t *S_out);

/* diffusion() for U utterances at once (utterance batching): rngs[u] is utterance u's own generator (each is
 * consumed exactly as tts_host_diffusion would), mel_out[u] holds 100 * S_out[u] floats. */
int tts_host_diffusion_batch(struct tts_ctx *ctx, tts_rng *const *rngs, int U, const float *const *latents, const int32_t *L,
                             int n_steps, float *const *mel_out, int32_t *S_out);

/* vocoder() (main.cpp:6044-6127): mel [100][S] -> audio [(S+10)*256-6] */
int tts_host_vocoder(struct tts_ctx *ctx, tts_rng *r, const float *mel, int S, float *audio_out);

#ifdef __cplusplus
}
#endif
#endif
