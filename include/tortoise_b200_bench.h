/* tortoise_b200_bench.h -- measurement-only entry points of libtortoise_b200.so (used by bench.py
 * and tools/; NOT part of the drop-in boundary in tortoise_b200.h). */
#ifndef TORTOISE_B200_BENCH_H
#define TORTOISE_B200_BENCH_H

#include "tortoise_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* micro-benchmark of the streaming GEMV kernel over all 30 layers' weights (what
 * bench.py's roofline figure is computed from): returns average ms per launch and the
 * algorithmic bytes per launch for the chosen op (0 qkv,1 attn-proj,2 fc,3 mlp-proj,4 lm-head) */
int tts_bench_gemv(tts_ctx *ctx, int32_t op, int32_t B, int32_t iters, float *ms_per_launch,
                   double *bytes_per_launch);

/* `iters` consecutive decode steps (after tts_ar_prefill) timed with CUDA events on the stream,
 * no host round trip in between: average ms per step and the algorithmic bytes of one step
 * (streamed weights + KV read/append + embeddings + logits). */
int tts_bench_decode_step(tts_ctx *ctx, int32_t iters, float *ms_per_step, double *bytes_per_step);


/* the denoiser's 3-tap convolution (tcgen05 GEMM, M = 2 S rows, N = K = 1024) on the loaded diffusion
 * weights: `iters` back-to-back launches between two CUDA events; ms per launch and FLOP per launch */
int tts_bench_conv3(tts_ctx *ctx, int32_t S, int32_t iters, float *ms_per_launch, double *flop_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* TORTOISE_B200_BENCH_H */
