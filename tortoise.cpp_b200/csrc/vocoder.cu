// vocoder.cu -- placeholder until the vocoder stage lands.
#include "common.cuh"
#include "engine.h"
namespace tts {
void voc_load(tts_ctx *, const char *) { throw ArgError("vocoder stage not built yet"); }
void voc_run(tts_ctx *, const float *, int, const float *, float *) { throw ArgError("vocoder stage not built yet"); }
void voc_free(tts_ctx *) {}
}
