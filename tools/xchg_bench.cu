// xchg_bench.cu -- dev micro-benchmark: what does ONE all-to-all exchange of a 1024-float vector
// between 148 persistent CTAs cost on B200, per protocol?  (The AR decode step is 121 dependent
// GEMV phases; the exchange between phases, not the weight stream, bounds it at 1 candidate.)
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/xchg_bench tools/xchg_bench.cu
//   run  : tools/xchg_bench            (prints cycles / round for every protocol variant)
// Protocols
//   0  LL pairs (value, tag) in one 8-byte store; every consumer thread polls its own 4 elements
//   1  data + per-producer flag: producers store, bar, thread 0 fence + flag store; warp 0 polls the
//      148 flags, bar, everybody loads the data
//   2  data + counter: like 1 with one atomic counter per replica
//   3  LL pairs polled by warp 0..(PW-1) only into shared memory, bar, everybody reads shared
// Every variant: REP replicas of the vector (CTA c reads replica c % REP), optional background
// TMA weight stream per CTA (to load L2/HBM like the real kernel does).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint4 ld_v4_vol(const void *p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_relaxed_v4(const void *p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ int g_relaxed;
__device__ __forceinline__ uint4 ld_v4(const void *p) { return g_relaxed ? ld_relaxed_v4(p) : ld_v4_vol(p); }
__device__ __forceinline__ void st_v2(void *p, uint32_t a, uint32_t b) {
  if (g_relaxed) asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
  else asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_vol(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(uint32_t *p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release(uint32_t *p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct XArgs {
  uint2 *ll;          // [2][REP][1024] (value, tag)
  float *data;        // [2][REP][1024]
  uint32_t *flags;    // [2][REP][256]
  uint32_t *counter;  // [REP * 32] (one 128-byte line each)
  const unsigned char *bg;  // background stream source (or null)
  size_t bg_bytes_per_cta;
  int rep, rounds, mode, sleep_ns, pollwarps, work, skew, relaxed, nvec;
  long long *cycles;  // [G]
  float *sink;
};

constexpr int NT = 256;

__global__ void __launch_bounds__(NT + 32, 1) xchg_kernel(XArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ float red[8];
  __shared__ float xs[1024];
  __shared__ volatile int done;
  __shared__ uint64_t bars[4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x, cta = blockIdx.x;
  if (tid == 0) {
    done = 0;
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == NT / 32) {
    // background stream: 4 x 16 KB ring, re-armed as soon as a stage lands
    if (lane == 0 && a.bg) {
      const unsigned char *src = a.bg + size_t(cta) * a.bg_bytes_per_cta;
      const size_t n = a.bg_bytes_per_cta / 16384;
      size_t it = 0;
      for (; it < 4; ++it) { mbar_arrive_expect_tx(&bars[it], 16384); bulk_g2s(smem + it * 16384, src + it * 16384, 16384, &bars[it]); }
      while (!done) {
        const int s = int(it & 3);
        while (!mbar_try_wait(&bars[s], uint32_t((it / 4 - 1) & 1))) {
        }
        mbar_arrive_expect_tx(&bars[s], 16384);
        bulk_g2s(smem + s * 16384, src + (it % n) * 16384, 16384, &bars[s]);
        ++it;
      }
      // drain: the last copy of every slot must land before the CTA exits
      for (size_t j = it - 4; j < it; ++j)
        while (!mbar_try_wait(&bars[j & 3], uint32_t((j / 4) & 1))) {
        }
      a.sink[1 + cta] = float(it);
    }
    return;
  }
  const int rep = cta % a.rep;
  const int base = 1024 / G, rem = 1024 % G;
  const int rows = base + (cta < rem ? 1 : 0), row0 = cta * base + min(cta, rem);
  float S = 1.0f;
  long long t0 = clock64();
  for (int r = 0; r < a.rounds; ++r) {
    const uint32_t tag = uint32_t(r + 1);
    const int par = r & 1;
    // ---- "work" between exchanges (dependent FMA chain of a.work steps) ----
    float w = S;
    {
      // per-(CTA, round) pseudo-random extra work in [0, skew) emulates GEMV finish-time skew
      int extra = 0;
      if (a.skew) extra = int((uint32_t(cta * 2654435761u) ^ uint32_t(r * 40503u)) >> 8) % a.skew;
      for (int i = 0; i < a.work + extra; ++i) w = fmaf(w, 1.0000001f, 1e-9f);
    }
    // ---- produce ----
    if (a.mode == 0 || a.mode == 3 || a.mode == 4) {
      for (int i = tid; i < rows; i += NT)
        for (int q = 0; q < a.rep; ++q)
          st_v2(a.ll + (size_t(par) * a.rep + q) * 1024 + row0 + i, __float_as_uint(w * 1e-3f + float(row0 + i) * 1e-6f), tag);
    } else {
      for (int i = tid; i < rows; i += NT)
        for (int q = 0; q < a.rep; ++q) a.data[(size_t(par) * a.rep + q) * 1024 + row0 + i] = w * 1e-3f + float(row0 + i) * 1e-6f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (a.mode == 1) {
        if (tid < a.rep) {
          __threadfence();
          st_release(a.flags + (size_t(par) * a.rep + tid) * 256 + cta, tag);
        }
      } else {
        if (tid < a.rep) {
          __threadfence();
          red_release(a.counter + tid * 32, 1u);
        }
      }
    }
    // ---- consume ----
    float x[4];
    if (a.mode == 0) {
      const uint2 *buf = a.ll + (size_t(par) * a.rep + rep) * 1024;
      uint4 v0 = ld_v4(buf + 2 * tid), v1 = ld_v4(buf + 512 + 2 * tid);
      while (v0.y != tag || v0.w != tag) { if (a.sleep_ns) __nanosleep(a.sleep_ns); v0 = ld_v4(buf + 2 * tid); }
      while (v1.y != tag || v1.w != tag) { if (a.sleep_ns) __nanosleep(a.sleep_ns); v1 = ld_v4(buf + 512 + 2 * tid); }
      x[0] = __uint_as_float(v0.x); x[1] = __uint_as_float(v0.z); x[2] = __uint_as_float(v1.x); x[3] = __uint_as_float(v1.z);
    } else if (a.mode == 4) {
      const uint2 *buf = a.ll + (size_t(par) * a.rep + rep) * 1024;
      if (lane == 0) {
        // sentinel: an element of a producer far away in the grid order
        const int sidx = ((warp * 131 + cta * 7) & 511) * 2;
        uint4 sv = ld_v4(buf + sidx);
        while (sv.y != tag) sv = ld_v4(buf + sidx);
      }
      __syncwarp();
      uint4 v0 = ld_v4(buf + 2 * tid), v1 = ld_v4(buf + 512 + 2 * tid);
      while (v0.y != tag || v0.w != tag) v0 = ld_v4(buf + 2 * tid);
      while (v1.y != tag || v1.w != tag) v1 = ld_v4(buf + 512 + 2 * tid);
      x[0] = __uint_as_float(v0.x); x[1] = __uint_as_float(v0.z); x[2] = __uint_as_float(v1.x); x[3] = __uint_as_float(v1.z);
    } else if (a.mode == 3) {
      const uint2 *buf = a.ll + (size_t(par) * a.rep + rep) * 1024;
      const int pw = a.pollwarps;  // warps 0..pw-1 fetch 1024 / pw elements each
      if (warp < pw) {
        const int per = 1024 / pw;  // elements of this warp
        for (int e = lane * 2; e < per; e += 64) {
          const int idx = warp * per + e;
          uint4 v = ld_v4(buf + idx);
          while (v.y != tag || v.w != tag) { if (a.sleep_ns) __nanosleep(a.sleep_ns); v = ld_v4(buf + idx); }
          xs[idx] = __uint_as_float(v.x);
          xs[idx + 1] = __uint_as_float(v.z);
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      x[0] = xs[2 * tid]; x[1] = xs[2 * tid + 1]; x[2] = xs[512 + 2 * tid]; x[3] = xs[512 + 2 * tid + 1];
    } else {
      if (warp == 0) {
        if (a.mode == 1) {
          const uint32_t *f = a.flags + (size_t(par) * a.rep + rep) * 256;
          for (int c = lane; c < G; c += 32)
            while (ld_acquire(f + c) != tag) { if (a.sleep_ns) __nanosleep(a.sleep_ns); }
        } else if (lane == 0) {
          const uint32_t want = uint32_t(G) * tag;
          while (ld_acquire(a.counter + rep * 32) < want) { if (a.sleep_ns) __nanosleep(a.sleep_ns); }
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float *d = a.data + (size_t(par) * a.rep + rep) * 1024;
      float2 p0, p1;
      asm volatile("ld.relaxed.gpu.global.v2.f32 {%0, %1}, [%2];" : "=f"(p0.x), "=f"(p0.y) : "l"(d + 2 * tid) : "memory");
      asm volatile("ld.relaxed.gpu.global.v2.f32 {%0, %1}, [%2];" : "=f"(p1.x), "=f"(p1.y) : "l"(d + 512 + 2 * tid) : "memory");
      x[0] = p0.x; x[1] = p0.y; x[2] = p1.x; x[3] = p1.y;
    }
    // ---- block reduction (what LayerNorm statistics need) ----
    float s = warp_sum((x[0] + x[1]) + (x[2] + x[3]));
    if (lane == 0) red[warp] = s;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    S = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) S += red[i];
    asm volatile("bar.sync 1, 256;" ::: "memory");
  }
  long long t1 = clock64();
  if (tid == 0) {
    a.cycles[cta] = t1 - t0;
    a.sink[0] = S;
    done = 1;
  }
}

int main(int argc, char **argv) {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  const int G = prop.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  printf("device %s, %d SMs, %d MHz\n", prop.name, G, khz / 1000);
  const int MAXREP = 148;
  XArgs a{};
  CK(cudaMalloc(&a.ll, sizeof(uint2) * 2 * MAXREP * 1024));
  CK(cudaMalloc(&a.data, sizeof(float) * 2 * MAXREP * 1024));
  CK(cudaMalloc(&a.flags, 4 * 2 * MAXREP * 256));
  CK(cudaMalloc(&a.counter, 4 * MAXREP * 32));
  CK(cudaMalloc(&a.cycles, 8 * G));
  CK(cudaMalloc(&a.sink, 4 * (G + 1)));
  unsigned char *bg = nullptr;
  const size_t bg_per = 4u << 20;
  CK(cudaMalloc(&bg, bg_per * G));
  CK(cudaMemset(bg, 1, bg_per * G));
  const size_t smem = 4 * 16384;
  CK(cudaFuncSetAttribute(xchg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int G_run = G;
  auto run = [&](int mode, int rep, int sleep_ns, int pollwarps, int work, bool with_bg) {
    CK(cudaMemset(a.ll, 0, sizeof(uint2) * 2 * MAXREP * 1024));
    CK(cudaMemset(a.flags, 0, 4 * 2 * MAXREP * 256));
    CK(cudaMemset(a.counter, 0, 4 * MAXREP * 32));
    a.mode = mode; a.rep = rep; a.sleep_ns = sleep_ns; a.pollwarps = pollwarps; a.work = work; a.rounds = 2000;
    a.bg = with_bg ? bg : nullptr;
    a.bg_bytes_per_cta = bg_per;
    void *args[] = {&a};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    CK(cudaLaunchCooperativeKernel((void *)xchg_kernel, dim3(G_run), dim3(NT + 32), args, smem, 0));
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> cyc(G_run);
    CK(cudaMemcpy(cyc.data(), a.cycles, 8 * G_run, cudaMemcpyDeviceToHost));
    long long mx = *std::max_element(cyc.begin(), cyc.end());
    printf("mode %d rep %3d sleep %3d pollwarps %d work %5d bg %d : %7.0f cycles/round  %6.3f us/round\n", mode, rep, sleep_ns,
           pollwarps, work, int(with_bg), double(mx) / a.rounds, ms * 1e3 / a.rounds);
    fflush(stdout);
  };
  auto run2 = [&](int mode, int rep, int pw, int work, int skew, bool with_bg, int relaxed, int grid) {
    a.skew = skew; a.relaxed = relaxed;
    CK(cudaMemcpyToSymbol(g_relaxed, &relaxed, sizeof(int)));
    G_run = grid;
    run(mode, rep, 0, pw, work, with_bg);
  };
  (void)run2;
  printf("== grid-size scaling (mode 0 rep 8 / mode 3 pw8 rep1), no bg\n");
  for (int grid : {2, 8, 16, 32, 74, 148}) { printf("grid %d: ", grid); run2(0, 8, 0, 0, 0, false, 0, grid); printf("grid %d: ", grid); run2(3, 1, 8, 0, 0, false, 0, grid); }
  printf("== relaxed.gpu instead of volatile\n");
  for (int rl : {0, 1}) { printf("relaxed %d: ", rl); run2(0, 8, 0, 0, 0, false, rl, G); printf("relaxed %d: ", rl); run2(3, 1, 8, 0, 0, false, rl, G); printf("relaxed %d: ", rl); run2(3, 8, 8, 0, 0, false, rl, G); }
  printf("== skewed arrivals: all-poll (0), smem-poll (3), sentinel (4); work 1000, skew S\n");
  for (int skew : {0, 500, 2000})
    for (int bgf : {0, 1}) {
      printf("skew %d: ", skew); run2(0, 8, 0, 1000, skew, bgf, 0, G);
      printf("skew %d: ", skew); run2(0, 1, 0, 1000, skew, bgf, 0, G);
      printf("skew %d: ", skew); run2(3, 1, 8, 1000, skew, bgf, 0, G);
      printf("skew %d: ", skew); run2(4, 8, 0, 1000, skew, bgf, 0, G);
      printf("skew %d: ", skew); run2(4, 1, 0, 1000, skew, bgf, 0, G);
      printf("skew %d: ", skew); run2(2, 8, 0, 1000, skew, bgf, 0, G);
    }
  return 0;
}
