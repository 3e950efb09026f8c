// diffusion.cu -- placeholder until the diffusion stage lands.
#include "common.cuh"
#include "engine.h"
namespace tts {
void diff_load(tts_ctx *, const char *) { throw ArgError("diffusion stage not built yet"); }
void diff_eps(tts_ctx *, const float *, int, const float *, int, int, int, float *) { throw ArgError("diffusion stage not built yet"); }
void diff_sample(tts_ctx *, const float *, int, int, int, const float *, float *) { throw ArgError("diffusion stage not built yet"); }
void diff_free(tts_ctx *) {}
}
