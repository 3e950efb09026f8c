// dev microbenchmark (not part of the product): cycles per tcgen05.mma (kind::f16, M = 128, K = 16, SS operands in
// SWIZZLE_128B K-major shared memory) as a function of N, issued back to back by one elected lane.
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t s32(const void *p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t desc(uint32_t a) {
  return uint64_t((a & 0x3FFFF) >> 4) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
template <int N, int MODE>
__global__ void __launch_bounds__(64) k(long long *out, int nmma) {
  extern __shared__ unsigned char raw[];
  unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < (16384 + N * 128) / 2; i += 64) reinterpret_cast<__half *>(base)[i] = __float2half(float((i * 7) % 13) * 0.01f);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&slot)), "r"(uint32_t(N < 32 ? 32 : N)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (uint32_t(N >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    const uint32_t sa = s32(base), sb = s32(base + 16384);
    long long t0 = clock64();
    int ph = 0;
    for (int j = 0; j < nmma / 4; ++j) {
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t acc = (j | kk) ? 1u : 0u;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm),
                       "l"(desc(sa + kk * 32)), "l"(desc(sb + kk * 32)), "r"(idesc), "r"(acc) : "memory");
        }
        if (MODE == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar[1])) : "memory");
      }
      __syncwarp();
      if (MODE == 2) {  // wait for every group of 4 (serialised issue -> latency of a 4-MMA chain)
        if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar[0])) : "memory");
        __syncwarp();
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&bar[0])), "r"(uint32_t(ph)) : "memory");
        ph ^= 1;
      }
    }
    long long t1 = clock64();
    if (MODE != 2) {
      if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar[0])) : "memory");
      __syncwarp();
      uint32_t ok = 0;
      while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&bar[0])), "r"(0u) : "memory");
    }
    long long t2 = clock64();
    if (threadIdx.x == 32) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(uint32_t(N < 32 ? 32 : N)) : "memory");
}
template <int N, int MODE>
void run(int grid, int nmma) {
  long long *d; CK(cudaMalloc(&d, grid * 16));
  const size_t smem = 16384 + N * 128 + 1024;
  CK(cudaFuncSetAttribute(k<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  for (int r = 0; r < 2; ++r) { k<N, MODE><<<grid, 64, smem>>>(d, nmma); CK(cudaDeviceSynchronize()); }
  long long h[2 * 148]; CK(cudaMemcpy(h, d, grid * 16, cudaMemcpyDeviceToHost));
  long long mx = 0, mi = 0; for (int i = 0; i < grid; ++i) { if (h[2 * i + 1] > mx) mx = h[2 * i + 1]; mi += h[2 * i]; }
  printf("N=%3d mode=%d grid=%3d nmma=%d: issue %.1f cyc/mma, complete %.1f cyc/mma (floor %d)\n", N, MODE, grid, nmma, double(mi) / grid / nmma, double(mx) / nmma, 128 * N / 256);
  CK(cudaFree(d));
}
int main() {
  for (int grid : {1, 148}) {
    run<32, 0>(grid, 512); run<64, 0>(grid, 512); run<128, 0>(grid, 512); run<256, 0>(grid, 512);
    run<128, 1>(grid, 512); run<128, 2>(grid, 512); run<32, 2>(grid, 512);
  }
  return 0;
}
