// common.cuh -- shared device/host helpers for the sm_100a kernels of libtortoise_b200.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <mutex>
#include <set>
#include <string>
#include <utility>

namespace tts {

constexpr int kDim = 1024;       // model width of all three stages' trunk
constexpr int kHeads = 16;
constexpr int kHeadDim = 64;
constexpr int kLayers = 30;
constexpr int kFF = 4096;
constexpr int kMelVocab = 8194;

// ---- error plumbing (the C-ABI never aborts; see tts_api.cu) -------------------------
struct Status {
  int code = 0;
  std::string msg;
  bool ok() const { return code == 0; }
};

#define TTS_CUDA_TRY(expr)                                                                  \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      char _b[512];                                                                         \
      snprintf(_b, sizeof _b, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                 \
               cudaGetErrorString(_e));                                                     \
      throw ::tts::CudaError(_b);                                                           \
    }                                                                                       \
  } while (0)

struct CudaError {
  std::string msg;
  explicit CudaError(const char *m) : msg(m) {}
};
struct ArgError {
  std::string msg;
  int code;
  explicit ArgError(const std::string &m, int c = -1) : msg(m), code(c) {}
};

// Opt-in to > 48 KB of dynamic shared memory.  The attribute is per DEVICE, a process may own
// contexts on several GPUs and C-ABI calls may come from several threads: remember (device, kernel).
template <typename K>
inline void ensure_smem_attr(K kernel, size_t bytes) {
  static std::mutex mu;
  static std::set<std::pair<int, const void *>> done;
  int dev = 0;
  TTS_CUDA_TRY(cudaGetDevice(&dev));
  const std::pair<int, const void *> key(dev, reinterpret_cast<const void *>(kernel));
  std::lock_guard<std::mutex> lk(mu);
  if (done.count(key)) return;
  TTS_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)));
  done.insert(key);
}

// ---- programmatic dependent launch (PDL) ----------------------------------------------
// Every kernel in a stage's chain is launched with programmatic stream serialization so
// that kernel N+1 is resident (and, for the weight-streaming GEMV, already pulling its
// weights through TMA bulk copies) while kernel N is still running.  Contract used
// throughout: a kernel touches no activation memory before pdl_wait().
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

struct Launcher {
  cudaStream_t stream = nullptr;
  bool pdl = true;
  int64_t *counter = nullptr;

  template <typename... KArgs, typename... Args>
  void operator()(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args) const {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TTS_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
    if (counter) ++*counter;
  }
};

// ---- mbarrier + TMA bulk copy (1-D, no tensor map) ---------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a system-dependent time limit)
__device__ __forceinline__ bool mbar_test_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy completing on an mbarrier (SASS: UBLKCP). bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// same with an L2 eviction-priority hint (createpolicy): streamed-once data must not displace the
// small hot working set (exchange buffers, KV cache) from L2
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
                                              uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// ---- small math helpers matching the reference's CPU numerics ----------------------------
// fp16 round trip (ggml_cpy F32->F16->F32, main.cpp:2789-2790; round-to-nearest-even)
__device__ __forceinline__ float h16(float x) { return __half2float(__float2half_rn(x)); }

// ggml CPU GELU: fp16 lookup table indexed by fp16(x), entry = fp16(gelu_tanh(f32(fp16 x)))
// (ggml.c:2193-2218, table built at ggml.c:3333).  Pass-through branches for |x| >= 10.
__device__ __forceinline__ float gelu16(float x) {
  if (x <= -10.0f) return 0.0f;
  if (x >= 10.0f) return x;
  const float xh = h16(x);
  const float g = 0.5f * xh * (1.0f + tanhf(0.79788456080286535587989211986876f * xh *
                                             (1.0f + 0.044715f * xh * xh)));
  return h16(g);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace tts
