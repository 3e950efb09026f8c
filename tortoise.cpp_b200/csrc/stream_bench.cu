// stream_bench.cu -- micro-benchmark of the weight-streaming mechanism itself (dev tool exposed
// through tts_bench_stream): how fast can 148 CTAs pull disjoint contiguous slices of HBM
// through a shared-memory ring of TMA bulk copies, as a function of stage size and depth?
#include "common.cuh"
#include "engine.h"

namespace tts {

__global__ void __launch_bounds__(288, 1) stream_ring_kernel(const unsigned char *src, size_t bytes_per_cta,
                                                             int stage_bytes, int stages, unsigned int *sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + size_t(stages) * stage_bytes);
  uint64_t *empty = full + stages;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
    fence_barrier_init();
  }
  __syncthreads();
  const unsigned char *base = src + size_t(blockIdx.x) * bytes_per_cta;
  const int n = int(bytes_per_cta / stage_bytes);
  if (warp == 8) {
    if (lane == 0)
      for (int it = 0; it < n; ++it) {
        const int slot = it % stages;
        mbar_wait(&empty[slot], ((it / stages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[slot], stage_bytes);
        bulk_g2s(smem + size_t(slot) * stage_bytes, base + size_t(it) * stage_bytes, stage_bytes, &full[slot]);
      }
    return;
  }
  unsigned int acc = 0;
  for (int it = 0; it < n; ++it) {
    const int slot = it % stages;
    mbar_wait(&full[slot], (it / stages) & 1);
    acc += *reinterpret_cast<const unsigned int *>(smem + size_t(slot) * stage_bytes + tid * 16);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[slot]);
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

// plain vectorised loads (LDG.128, 4 independent loads in flight per thread) for comparison
__global__ void __launch_bounds__(256) stream_ldg_kernel(const uint4 *src, size_t n16_per_cta, unsigned int *sink) {
  const uint4 *p = src + size_t(blockIdx.x) * n16_per_cta;
  unsigned int acc = 0;
  for (size_t i = threadIdx.x; i + 768 < n16_per_cta; i += 1024) {
    const uint4 a = __ldcs(p + i), b = __ldcs(p + i + 256), c = __ldcs(p + i + 512), d = __ldcs(p + i + 768);
    acc += a.x ^ b.y ^ c.z ^ d.w;
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

void bench_stream(tts_ctx *c, int mode, int stage_bytes, int stages, size_t bytes_per_cta, int iters, float *ms,
                  double *bytes) {
  const int G = c->num_sms;
  const size_t total = size_t(G) * bytes_per_cta;
  unsigned char *buf = nullptr;
  unsigned int *sink = nullptr;
  TTS_CUDA_TRY(cudaMalloc(&buf, total));
  TTS_CUDA_TRY(cudaMalloc(&sink, 4));
  TTS_CUDA_TRY(cudaMemsetAsync(buf, 1, total, c->stream));
  const size_t smem = size_t(stages) * stage_bytes + 2 * stages * 8 + 64;
  if (mode == 0) TTS_CUDA_TRY(cudaFuncSetAttribute(stream_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto run = [&]() {
    if (mode == 0) stream_ring_kernel<<<G, 288, smem, c->stream>>>(buf, bytes_per_cta, stage_bytes, stages, sink);
    else stream_ldg_kernel<<<G, 256, 0, c->stream>>>((const uint4 *)buf, bytes_per_cta / 16, sink);
  };
  run();
  TTS_CUDA_TRY(cudaGetLastError());
  TTS_CUDA_TRY(cudaEventRecord(e0, c->stream));
  for (int i = 0; i < iters; ++i) run();
  TTS_CUDA_TRY(cudaEventRecord(e1, c->stream));
  TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
  float t = 0;
  cudaEventElapsedTime(&t, e0, e1);
  *ms = t / iters;
  *bytes = double(total);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  cudaFree(sink);
}

}  // namespace tts
