/* tortoise_b200.h -- C-ABI of libtortoise_b200.so: the drop-in boundary of the B200-native
 * tortoise-tts hot path (AR mel-token decoder -> diffusion denoiser -> UnivNet vocoder).
 *
 * The reference (balisujohn/tortoise.cpp) has no FFI: its hot path sits behind three C++
 * stage functions called by main() (main.cpp:6570-6577) which build ggml graphs and hand
 * them to ggml_backend_graph_compute (main.cpp:5186, 5247, 5342, 5838, 5955, 6112).  Each
 * entry point below names the reference interface it replaces.  Plain pointers and sizes
 * only; all pointers are caller-owned HOST memory unless the name ends in _dev.  Every
 * function returns 0 on success or a negative TTS_E* code (never aborts); the message is
 * available from tts_last_error().  A context is single-threaded and bound to one GPU.
 *
 * There is NO CPU fallback: tts_init fails with TTS_ENODEV when no sm_100 device exists.
 */
#ifndef TORTOISE_B200_H
#define TORTOISE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TTS_OK 0
#define TTS_EINVAL (-1)   /* bad argument / call order              */
#define TTS_EIO (-2)      /* file missing / malformed container     */
#define TTS_ENODEV (-3)   /* no CUDA device / wrong architecture    */
#define TTS_ECUDA (-4)    /* CUDA runtime error (see last_error)    */
#define TTS_ELIMIT (-5)   /* sequence / batch limit exceeded        */

/* weight storage / arithmetic mode of the three stages */
#define TTS_DTYPE_F32 0   /* parity mode: f32 weights, reference numerics         */
#define TTS_DTYPE_F16 1   /* fast mode:   f16 weight streaming, f32 accumulation  */

#define TTS_MEL_VOCAB 8194
#define TTS_MEL_START 8192
#define TTS_MEL_STOP 8193
#define TTS_DIM 1024

typedef struct tts_ctx tts_ctx;

typedef struct tts_config {
  int32_t device;          /* CUDA ordinal                                             */
  int32_t dtype;           /* TTS_DTYPE_*  (AR weights; convs are always f16xf16->f32
                              like the reference's ggml_conv_1d, ggml.c:6493-6508)       */
  int32_t max_batch;       /* max AR candidates resident on this GPU (reference: 4,
                              main.cpp:794-797; here up to 64)                          */
  int32_t max_positions;   /* KV slots per candidate (reference: 404, main.cpp:608-611) */
  int32_t parity_quirks;   /* 1 = reproduce reference quirks (B=1 mel-position bug A-4) */
  int32_t reserved[3];
} tts_config;

/* ---- lifetime ---------------------------------------------------------------------- */
int tts_init(const tts_config *cfg, tts_ctx **out);
void tts_free(tts_ctx *ctx);
const char *tts_last_error(const tts_ctx *ctx); /* ctx may be NULL: last global error */
int tts_version(void);

/* ---- loaders: replace autoregressive_model_load (main.cpp:482), diffusion_model_load
 *      (main.cpp:931), vocoder_model_load (main.cpp:1665).  Same container format
 *      (magic 0x67676d6c + records), same tensor names; unknown names are an error. --- */
int tts_load_ar(tts_ctx *ctx, const char *path);
int tts_load_diffusion(tts_ctx *ctx, const char *path);
int tts_load_vocoder(tts_ctx *ctx, const char *path);

/* ---- AR stage ---------------------------------------------------------------------- */
/* Replaces the prefill graph run, autoregressive_graph(fake_inputs=true) + compute
 * (main.cpp:5131-5186): embeds [voice | text(T) | start-mel], fills the KV cache of all
 * B candidates, returns logits [B][8194] of the last position. text includes 255 ... 0. */
int tts_ar_prefill(tts_ctx *ctx, const int32_t *text_tokens, int32_t T, const float *voice_1024,
                   int32_t B, float *logits_out);

/* Utterance batching of the decode loop (BASELINE.json configs[4]; no reference counterpart -- the reference
 * prefills ONE prompt, main.cpp:5131-5186): U <= min(16, max_batch) DIFFERENT prompts are prefilled one after the
 * other, prompt u into candidate slot u with its K/V rows right-aligned to the longest prompt, and the following
 * tts_ar_step / tts_ar_step_topk calls decode all U on one weight stream (slot u attends to its own rows only).
 * Needs dtype f16, max_batch >= 5 and max_positions <= 1024.  logits_out [U][8194]: first-step logits per prompt. */
int tts_ar_prefill_multi(tts_ctx *ctx, int32_t U, const int32_t *const *text_tokens, const int32_t *T,
                         const float *voice_1024, float *logits_out);

/* Replaces one decode iteration, autoregressive_graph(fake_inputs=false) + compute
 * (main.cpp:5227-5247): feeds tokens[B] at mel position id pos_id (reference passes i+2),
 * appends to the KV cache, returns logits [B][8194]. */
int tts_ar_step(tts_ctx *ctx, const int32_t *tokens_B, int32_t pos_id, float *logits_out);

/* Same step with outputs left on the device and no host sync (bench / fused samplers):
 * logits stay in the context's device buffer; *logits_dev receives its address. */
int tts_ar_step_dev(tts_ctx *ctx, const int32_t *tokens_B, int32_t pos_id,
                    const float **logits_dev);

/* The same step with the reference's per-step D2H of 8194 floats per candidate (main.cpp:4767)
 * replaced by a device-side pre-selection: vals_out / idx_out [B][TTS_AR_TOPK] receive the
 * TTS_AR_TOPK largest logits of every candidate as unsorted (value, index) pairs (ties at the
 * threshold: lowest indices first), flags_out[B] != 0 marks a candidate whose row must be fetched
 * with tts_ar_logits instead (more than 256 ties at the threshold).  64 entries are enough for the
 * host sampler to reproduce process_logits_and_sample bit-exactly (tts_host_sample_sparse proves
 * it per call or asks for the full row). */
#define TTS_AR_TOPK 64
int tts_ar_step_topk(tts_ctx *ctx, const int32_t *tokens_B, int32_t pos_id, float *vals_out, int32_t *idx_out,
                     int32_t *flags_out);

/* full logits [B][8194] of the last tts_ar_prefill / tts_ar_step* call (blocking) */
int tts_ar_logits(tts_ctx *ctx, float *logits_out);

/* Replaces the latent pass, autoregressive_latent_graph + compute (main.cpp:5280-5352):
 * codes is [B][502] (8192, 500 codes, 8193 as built by apply_padding main.cpp:4510);
 * out is [B][500][1024] (lm_head.0-normalised hidden state of mel positions 0..499).
 * n_keep (1..500): only mel rows < n_keep are computed (causality makes later rows
 * irrelevant to them -- exact, not approximate); rows >= n_keep are returned as zeros.
 * n_keep = 500 is the reference's full pass. */
int tts_ar_latents(tts_ctx *ctx, const int32_t *text_tokens, int32_t T, const float *voice_1024,
                   const int32_t *codes_Bx502, int32_t B, int32_t n_keep, float *latents_out);

/* ---- diffusion stage --------------------------------------------------------------- */
/* Replaces ONE diffusion_graph run (main.cpp:5749-5841 conditioned / 5866-5961
 * unconditioned): latents [L][1024], x [100][S] (channel-major like the reference's
 * noise_tensor), timestep = timestep_map value; out [200][S]. */
int tts_diffusion_eps(tts_ctx *ctx, const float *latents, int32_t L, const float *x, int32_t S,
                      int32_t timestep, int32_t conditioning_free, float *out_200xS);

/* Replaces the whole sampling loop of diffusion() (main.cpp:5723-6033): noise holds
 * (n_steps + 1) blocks of 100*S normals drawn by the host in the reference's RNG order
 * (initial x, then one block per step, main.cpp:5638 and 6020); mel_out [100][S].       */
int tts_diffusion_sample(tts_ctx *ctx, const float *latents, int32_t L, int32_t S,
                         int32_t n_steps, const float *noise, float *mel_out_100xS);

/* Streaming form of the same loop, so the host can draw noise block i+1 (reference RNG order)
 * while the GPU runs step i: begin(x0 = first 100*S draws) ; n_steps x step(next 100*S draws,
 * asynchronous) ; end(mel_out) blocks.  tts_diffusion_sample == begin + steps + end. */
int tts_diffusion_begin(tts_ctx *ctx, const float *latents, int32_t L, int32_t S, int32_t n_steps, const float *x0);
int tts_diffusion_step(tts_ctx *ctx, const float *noise_block_100xS);
int tts_diffusion_end(tts_ctx *ctx, float *mel_out_100xS);

/* The same loop for a BATCH of U utterances of different lengths (BASELINE configs[4], utterance batching;
 * the reference renders one utterance per process): utterance u has latents[u] [L[u]][1024], S[u] mel frames
 * and its own noise stream (x0[u] / noise_blocks[u] = 100 * S[u] normals each, drawn by the host in the
 * reference's order from that utterance's generator).  One launch set serves all of them: the GEMMs see
 * 2 * sum(S[u]) rows.  Results equal the one-at-a-time calls up to f32 summation order. */
int tts_diffusion_begin_batch(tts_ctx *ctx, int32_t U, const float *const *latents, const int32_t *L, const int32_t *S,
                              int32_t n_steps, const float *const *x0);
int tts_diffusion_step_batch(tts_ctx *ctx, const float *const *noise_blocks);
int tts_diffusion_end_batch(tts_ctx *ctx, float *const *mel_out);

/* ---- vocoder stage ----------------------------------------------------------------- */
/* Replaces vocoder_graph + compute (main.cpp:6078-6122): mel [100][S] NORMALISED
 * (denormalisation main.cpp:5575 is done on the device), noise [(S+10)][64] as drawn by
 * the host (main.cpp:6057); audio_out has (S+10)*256-6 samples. */
int tts_vocoder(tts_ctx *ctx, const float *mel_100xS, int32_t S, const float *noise,
                float *audio_out);

/* ---- multi-GPU: the final candidate gather / selection (the path's only exchange; NCCL) ------------
 * Candidates are sharded over GPUs (weights replicated, no data-path collective); every rank contributes
 * (score, n_codes) per candidate to ONE ncclAllGather and the best score wins (first maximum, NaN loses).
 * The reference has neither a scorer nor a multi-GPU path (main() diffuses candidate 0, main.cpp:6575).
 * A group is formed either inside one process that owns several contexts (tts_group_init_local,
 * ncclCommInitAll; what `tortoise --gpus N` does) or with one process per GPU (tts_group_init_rank with a
 * 128-byte id from tts_nccl_unique_id distributed by the launcher; what bench.py does under torchrun).
 * libnccl.so.2 is dlopen'ed on first use: TTS_ENODEV when it is absent.  Free the group before its contexts. */
typedef struct tts_group tts_group;
int tts_nccl_unique_id(char *out128);
int tts_group_init_local(tts_ctx **ctxs, int32_t n, tts_group **out);
int tts_group_init_rank(tts_ctx *ctx, int32_t rank, int32_t world, const char *id128, tts_group **out);
/* scores / lens: [n_local][per] with n_local = number of LOCAL members (all ranks in local mode, 1 in rank
 * mode).  *winner = global candidate index rank * per + i; scores_all / lens_all [world][per] (may be NULL). */
int tts_gather_select(tts_group *g, const float *scores, const int32_t *lens, int32_t per, int32_t *winner,
                      float *scores_all, int32_t *lens_all);
const char *tts_group_last_error(const tts_group *g);
void tts_group_free(tts_group *g);

/* blocks until all work queued by this context has finished (pairs with *_dev calls) */
int tts_sync(tts_ctx *ctx);

/* ---- timing / introspection -------------------------------------------------------- */
/* number of kernels launched by this context since creation (bench "gpu_launches") */
int64_t tts_launch_count(const tts_ctx *ctx);
/* device-side duration in ms of the last tts_* stage call (CUDA events on the stream) */
float tts_last_stage_ms(const tts_ctx *ctx);
/* sum of the device-side durations of all stage calls since tts_init (CUDA events) */
double tts_device_ms_total(const tts_ctx *ctx);
#ifdef __cplusplus
}
#endif
#endif /* TORTOISE_B200_H */
