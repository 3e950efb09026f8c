// tc5v2.cuh -- tcgen05 GEMM / implicit conv, second generation: TMA tensor-map loads,
// 128 x 128 output tiles and split-K across a thread-block cluster.
//
// Same contract as tc5gemm_kernel / tgemm_kernel (gemm.cuh: C = (Ahi+Alo) x (Whi+Wlo)^T over
// `taps` shifted K passes, f32 accumulation in TMEM, fused epilogue).  What changed, and why:
//   * the first-generation kernel (128 x 32 tiles, cp.async loaders) re-read every A tile from
//     32 CTAs at once: ~120 MB of L2->SM traffic per 3-tap convolution, served at ~3 TB/s
//     (profiles/r01d_tc5v2_ncu_full_raw.csv) -- 4x under both the L2 and the per-SM LSU limits, i.e.
//     bound by the same few L2 lines being hammered by every CTA.  128 x 128 tiles cut the
//     traffic 2.5x; the lost parallelism (32 tiles for the [382 x 1024] problems of the
//     denoiser) comes back as split-K: `csz` CTAs of one cluster (1,1,csz) each accumulate a
//     K slice of the SAME output tile in their own TMEM, so concurrent CTAs pull DIFFERENT
//     bytes;
//   * operands arrive through TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B boxes of 64 halves x
//     128 rows) issued by ONE thread -- no per-lane address math, no LDGSTS issue limit;
//     the weight boxes of the first pipeline stages are requested BEFORE griddepcontrol.wait
//     (weights never depend on the previous kernel), activations after;
//   * split-K reduction is deterministic: every CTA parks its f32 partial tile in its own
//     shared memory (row-major, padded), the cluster synchronises, and CTA r sums rows
//     [r * 128 / csz, (r + 1) * 128 / csz) over ranks 0..csz-1 through distributed shared
//     memory in fixed order, then runs the epilogue with fully coalesced 512-byte row stores;
//   * GroupNorm statistics of the output (one 128-column tile = 4 groups) are reduced in the
//     same epilogue in double, per (sequence, group, M tile, cluster rank) in fixed order.
// warps 0-7: TMEM -> smem dump, reduction, epilogue; warp 8 lane 0: tcgen05.mma issuer;
// warp 9 lane 0: TMA producer.
#pragma once
#include <cuda.h>

#include <map>
#include <mutex>
#include <tuple>
#include <type_traits>

#include "tc5gemm.cuh"

namespace tts {

constexpr int T6_BM = 128, T6_BN = 128, T6_BK = 64;
constexpr int T6_MMA_WARP = 8, T6_PROD_WARP = 9, T6_THREADS = 10 * 32;
constexpr int T6_PLANE = 128 * 128;        // bytes of one 128-row x 64-half swizzled box
constexpr int T6_CTRL = 1024;              // barriers + tmem slot + GN scratch
constexpr int T6_MAX_STAGES = 8;
#ifndef T6_MIN_CTAS
#define T6_MIN_CTAS 1  // 2 = register budget for two co-resident CTAs (experiment below: no gain)
#endif

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t local_addr, uint32_t rank) {
  uint32_t ra;
  float4 v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(ra)
               : "memory");
  return v;
}

// rows [row_base, row_base + rows_per) of the tile: split-K sum over the cluster's partial tiles, epilogue, stores.
// EPI is a template parameter: with the epilogue kind read from the kernel argument every element went through an
// indirect branch (LDC + BRX, four per row and lane) and the loop ran at ~450 ns per row -- 7.4 us of a 26 us GEMM
// at csz = 1, 4.3 us of 11.7 at csz = 4 (tc5trace, round 2).
template <int EPI, int RED_LD, bool SPLITK>
__device__ __forceinline__ void t6_epilogue_rows(const TGemmArgs &g, const float *red, int csz, int rows_per, int row_base,
                                                 int rows_valid, int m0, int n, int c, bool vec, const float (&bias)[4], int warp,
                                                 bool do_gn, double &gs1, double &gs2) {
  constexpr bool resid = EPI == E_BIAS_RESID || EPI == E_BIAS_LRELU_RESID;
  const uint32_t red_u32 = smem_u32(red);
  const int nr = min(rows_per, rows_valid - row_base);  // this rank's rows that exist (<= 0: none)
  // split-K sum of row r: every rank's partial in flight before the first use, summed in rank order (deterministic)
  auto row_sum = [&](int r) {
    if (!SPLITK || csz == 1) return *reinterpret_cast<const float4 *>(red + size_t(r) * RED_LD + c);
    const uint32_t off = uint32_t(r * RED_LD + c) * 4;
    float4 p[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < csz) p[k] = ld_dsmem_f4(red_u32 + off, k);
    float4 a4 = p[0];
#pragma unroll
    for (int k = 1; k < 8; ++k)
      if (k < csz) { a4.x += p[k].x; a4.y += p[k].y; a4.z += p[k].z; a4.w += p[k].w; }
    return a4;
  };
  if (vec) {
    // unsplit tiles: four rows per trip, their loads issued together -- with 8 epilogue warps on the SM the loop is
    // bound by the latency of its own dependent instructions (tc5trace: ~200 ns per single-row trip, 10 -> 4.7 us
    // for a 128 x 256 tile)
    float *const cbase = g.C ? g.C + size_t(m0 + row_base) * g.ldc + n : nullptr;
    __half *const hbase = g.Chi ? g.Chi + size_t(m0 + row_base) * g.ldh + n : nullptr;
    __half *const lbase = g.Clo ? g.Clo + size_t(m0 + row_base) * g.ldh + n : nullptr;
    auto trips = [&](auto uc) {
    constexpr int U = decltype(uc)::value;
    for (int rr0 = warp; rr0 < nr; rr0 += 8 * U) {
      float4 a4[U], o4[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int rr = rr0 + 8 * u;
        o4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rr < nr) {
          a4[u] = row_sum(row_base + rr);
          if (resid) o4[u] = *reinterpret_cast<const float4 *>(cbase + size_t(rr) * g.ldc);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int rr = rr0 + 8 * u;
        if (rr >= nr) break;
        const float acc[4] = {a4[u].x, a4[u].y, a4[u].z, a4[u].w}, old[4] = {o4[u].x, o4[u].y, o4[u].z, o4[u].w};
        float out[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) out[e] = apply_epi(EPI, acc[e], bias[e], old[e]);
        if (cbase) *reinterpret_cast<float4 *>(cbase + size_t(rr) * g.ldc) = make_float4(out[0], out[1], out[2], out[3]);
        if (hbase) {
          __half hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            hi[e] = __float2half_rn(out[e]);
            lo[e] = __float2half_rn(out[e] - __half2float(hi[e]));
          }
          *reinterpret_cast<uint2 *>(hbase + size_t(rr) * g.ldh) = *reinterpret_cast<uint2 *>(hi);
          if (lbase) *reinterpret_cast<uint2 *>(lbase + size_t(rr) * g.ldh) = *reinterpret_cast<uint2 *>(lo);
        }
        if (do_gn) {
          // 4-value partials in f32 (relative error 1e-7, the rounding of the values themselves), rows summed in
          // double: two conversions and two additions per row on the FP64 pipe instead of fifteen
          gs1 += double((out[0] + out[1]) + (out[2] + out[3]));
          gs2 += double((out[0] * out[0] + out[1] * out[1]) + (out[2] * out[2] + out[3] * out[3]));
        }
      }
    }
    };
    // (split-K ranks: one row per trip -- four rows x csz partials in flight over distributed shared memory measured
    // SLOWER, 3.9 vs 2.2 us for the 32 rows of a rank at csz = 4: that path is bound by DSMEM bandwidth)
    if (!SPLITK || csz == 1) trips(std::integral_constant<int, 4>{});
    else trips(std::integral_constant<int, 1>{});
  } else {
    for (int rr = warp; rr < nr; rr += 8) {
      const int r = row_base + rr, m = m0 + r;
      const float4 a4 = row_sum(r);
      const float acc[4] = {a4.x, a4.y, a4.z, a4.w};
      for (int e = 0; e < 4 && n + e < g.N; ++e) {
        float old = 0.f;
        if (resid) old = g.C[size_t(m) * g.ldc + n + e];
        const float o = apply_epi(EPI, acc[e], bias[e], old);
        if (g.C) g.C[size_t(m) * g.ldc + n + e] = o;
        if (g.Chi) {
          const __half hi = __float2half_rn(o);
          g.Chi[size_t(m) * g.ldh + n + e] = hi;
          if (g.Clo) g.Clo[size_t(m) * g.ldh + n + e] = __float2half_rn(o - __half2float(hi));
        }
      }
    }
  }
}

// BN = 128, or 256 (two 128-row weight boxes side by side in a stage, one N = 256 MMA per k16 step): per k-block an
// SM then takes in 16 + 32 KB for twice the FLOPs of the 16 + 16 KB of BN = 128 -- the GEMM is bound by the bytes an SM
// can pull from L2 (profiles/r02_diffusion_step_experiments.md), and at M = 2 S >> 1024 rows the 128-wide tiles also
// no longer fit one wave.
template <int BN>
static __global__ void __launch_bounds__(T6_THREADS, T6_MIN_CTAS)
    tc5v2_kernel(TGemmArgs g, const __grid_constant__ CUtensorMap mAhi, const __grid_constant__ CUtensorMap mAlo,
                 const __grid_constant__ CUtensorMap mWhi, const __grid_constant__ CUtensorMap mWlo, int csz,
                 int stages) {
  extern __shared__ unsigned char t6_raw[];
  unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(t6_raw) + 1023) & ~uintptr_t(1023));
  const bool has_alo = g.Alo != nullptr, has_wlo = g.Wlo != nullptr;
  const uint32_t a_planes = has_alo ? 2 : 1, w_planes = has_wlo ? 2 : 1;
  constexpr int NH = BN / 128;         // 128-row weight boxes per plane and stage
  constexpr int RED_LD = BN + 4;       // padded row of the f32 partial tile (conflict-free both ways)
  constexpr uint32_t W_PLANE = NH * T6_PLANE;
  const uint32_t stage_bytes = T6_PLANE * a_planes + W_PLANE * w_planes;
  // the ring doubles as the f32 partial tile after the MMAs: at least 128 x (BN + 4) x 4 bytes
  const size_t ring_bytes = max(size_t(stages) * stage_bytes, size_t(T6_BM) * RED_LD * 4);
  unsigned char *ctrl = base + ring_bytes;
  uint64_t *full = reinterpret_cast<uint64_t *>(ctrl);
  uint64_t *empty = full + T6_MAX_STAGES;
  uint64_t *done = empty + T6_MAX_STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);
  double *gw = reinterpret_cast<double *>(ctrl + 256);  // [8 warps][4 groups][2]
  float *red = reinterpret_cast<float *>(base);         // partial tile, reuses the stage ring after the MMAs

  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int rank = blockIdx.z % csz, seq = blockIdx.z / csz;
  const int t0 = blockIdx.y * T6_BM;
  const int rows_valid = max(0, min(T6_BM, (g.Tseq ? g.Tseq[seq] : g.T) - t0));
  const int m0 = seq * g.T + t0, n0 = blockIdx.x * BN;
  const int kchunks = g.K / T6_BK;
  const int iters = g.taps * kchunks;
  const int it0 = int(int64_t(iters) * rank / csz), it1 = int(int64_t(iters) * (rank + 1) / csz);
  const int nit = it1 - it0;
  const bool tr = g.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  auto trace = [&](int slot) {
    if (tr) {
      long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      g.dbg[slot] = t;
    }
  };
  if (tid == 0) trace(0);

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);   // producer's arrive.expect_tx; the TMA engine completes the bytes
      mbar_init(&empty[s], 1);  // tcgen05.commit
    }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == T6_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(uint32_t(BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc5_fence_before();
  __syncthreads();
  tc5_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  pdl_launch_dependents();
  if (tid == 0) trace(1);

  if (warp == T6_PROD_WARP) {
    {
      // ================= TMA producer (the warp walks the loop, one elected lane issues: see elect_one) =================
      const int a_row0 = seq * (g.T + 2 * g.halo) + t0 + g.halo - g.pad;
      auto load_w = [&](int j) {
        const int it = it0 + j, s = j % stages;
        const int tap = it / kchunks, k0 = (it % kchunks) * T6_BK;
        unsigned char *sp = base + size_t(s) * stage_bytes + T6_PLANE * a_planes;
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
          tma_load_2d(sp + hh * T6_PLANE, &mWhi, k0, tap * g.N + n0 + hh * 128, &full[s]);
          if (has_wlo) tma_load_2d(sp + W_PLANE + hh * T6_PLANE, &mWlo, k0, tap * g.N + n0 + hh * 128, &full[s]);
        }
      };
      auto load_a = [&](int j) {
        const int it = it0 + j, s = j % stages;
        const int tap = it / kchunks, k0 = (it % kchunks) * T6_BK;
        unsigned char *sp = base + size_t(s) * stage_bytes;
        tma_load_2d(sp, &mAhi, k0, a_row0 + tap * g.dil, &full[s]);
        if (has_alo) tma_load_2d(sp + T6_PLANE, &mAlo, k0, a_row0 + tap * g.dil, &full[s]);
      };
      const int pre = min(stages, nit);
      for (int j = 0; j < pre; ++j) {  // weights first: they do not depend on the previous kernel
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[j], stage_bytes);
          load_w(j);
        }
        __syncwarp();
      }
      if (lane == 0) trace(2);
      pdl_wait();
      if (lane == 0) trace(3);
      for (int j = 0; j < pre; ++j) {
        if (elect_one()) load_a(j);
        __syncwarp();
      }
      if (lane == 0) trace(4);
      for (int j = pre; j < nit; ++j) {
        const int s = j % stages;
        mbar_wait(&empty[s], ((j / stages) & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[s], stage_bytes);
          load_w(j);
          load_a(j);
        }
        __syncwarp();
      }
    }
  } else if (warp == T6_MMA_WARP) {
    {
      // ================= MMA issuer (the warp walks the loop, one elected lane issues) =================
      const uint32_t idesc = (1u << 4) | (uint32_t(BN >> 3) << 17) | (uint32_t(T6_BM >> 4) << 24);
      for (int j = 0; j < nit; ++j) {
        const int s = j % stages;
        mbar_wait(&full[s], (j / stages) & 1);
        tc5_fence_after();
        if (j == 0 && lane == 0) trace(5);
        if (elect_one()) {
        const uint32_t sa = smem_u32(base + size_t(s) * stage_bytes);
        const uint32_t sa_lo = sa + T6_PLANE;
        const uint32_t sw = sa + T6_PLANE * a_planes;
        const uint32_t sw_lo = sw + W_PLANE;
        uint32_t acc = j > 0 ? 1u : 0u;  // (derived from the iteration, not carried: whichever lane is elected sees it)
#pragma unroll
        for (int k = 0; k < T6_BK / 16; ++k) {
          const uint32_t koff = k * 32;
          if (has_wlo) { umma_f16(tmem_d, umma_desc_sw128(sa + koff), umma_desc_sw128(sw_lo + koff), idesc, acc); acc = 1; }
          if (has_alo) { umma_f16(tmem_d, umma_desc_sw128(sa_lo + koff), umma_desc_sw128(sw + koff), idesc, acc); acc = 1; }
          umma_f16(tmem_d, umma_desc_sw128(sa + koff), umma_desc_sw128(sw + koff), idesc, acc);
          acc = 1;
        }
        umma_commit(&empty[s]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(done);
      __syncwarp();
      if (lane == 0) trace(6);
    }
  } else {
    // ================= TMEM -> shared partial tile =================
    pdl_wait();  // the epilogue below reads / overwrites C
    mbar_wait(done, 0);
    tc5_fence_after();
    if (tid == 0) trace(7);
    const int q = warp & 3, h = warp >> 2;  // TMEM lane quadrant, column half
    const int r = q * 32 + lane;
#pragma unroll 1
    for (int cb = 0; cb < BN / 2; cb += 32) {
      float v[32];
      tmem_ld32(tmem_d + (uint32_t(q * 32) << 16) + h * (BN / 2) + cb, v);
      float *dst = red + size_t(r) * RED_LD + h * (BN / 2) + cb;
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4 *>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  }
  __syncwarp();
  tc5_fence_before();
  if (tid == 0) trace(8);
  if (csz > 1) cluster_sync_all();
  else __syncthreads();
  if (tid == 0) trace(9);

  if (warp < 8) {
    // ================= split-K reduction + epilogue: rank r owns rows [r, r+1) * 128 / csz =================
    const int rows_per = T6_BM / csz, row_base = rank * rows_per;
    const bool do_gn = g.gn_partial != nullptr;
#pragma unroll 1
    for (int jh = 0; jh < NH; ++jh) {  // 128-column halves of the tile
      const int c = jh * 128 + lane * 4, n = n0 + c;
      const bool vec = (n + 3 < g.N) && (g.ldc % 4 == 0) && (!g.Chi || g.ldh % 4 == 0);
      float bias[4] = {0.f, 0.f, 0.f, 0.f};
      if (g.bias) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (n + e < g.N) bias[e] = g.bias[n + e];
      }
      double gs1 = 0.0, gs2 = 0.0;
      switch (g.epi) {
#define T6_CASE(E) case E: t6_epilogue_rows<E, RED_LD, BN == 128>(g, red, csz, rows_per, row_base, rows_valid, m0, n, c, vec, bias, warp, do_gn, gs1, gs2); break;
        T6_CASE(E_NONE) T6_CASE(E_BIAS) T6_CASE(E_BIAS_H16) T6_CASE(E_BIAS_GELU16) T6_CASE(E_BIAS_RESID) T6_CASE(E_BIAS_LRELU)
        T6_CASE(E_BIAS_LRELU_RESID)
#undef T6_CASE
        default: break;
      }
      if (do_gn) {
        // lanes 8 i .. 8 i + 7 hold the 32 columns of GroupNorm group (n0 + 128 jh) / 32 + i
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          gs1 += __shfl_xor_sync(0xffffffffu, gs1, o);
          gs2 += __shfl_xor_sync(0xffffffffu, gs2, o);
        }
        if ((lane & 7) == 0) {
          gw[(warp * 4 + (lane >> 3)) * 2] = gs1;
          gw[(warp * 4 + (lane >> 3)) * 2 + 1] = gs2;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid < 4) {
          double s1 = 0.0, s2 = 0.0;
          for (int w = 0; w < 8; ++w) {
            s1 += gw[(w * 4 + tid) * 2];
            s2 += gw[(w * 4 + tid) * 2 + 1];
          }
          double *o = g.gn_partial + ((size_t(seq) * 32 + ((n0 + jh * 128) >> 5) + tid) * g.gn_mtiles + blockIdx.y * csz + rank) * 2;
          o[0] = s1;
          o[1] = s2;
        }
        if (NH > 1) asm volatile("bar.sync 1, 256;" ::: "memory");  // gw is reused by the next half
      }
    }
    if (tid == 0) trace(10);
  }
  __syncwarp();
  if (tid == 0) trace(11);
  if (csz > 1) cluster_sync_all();  // peers have finished reading this CTA's partial tile
  else __syncthreads();
  if (tid == 0) trace(12);
  if (warp == T6_MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(uint32_t(BN)) : "memory");
  }
}

// ---- host side: tensor maps (driver entry point fetched through the runtime; no -lcuda) ----
typedef CUresult (*tts_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                        const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

static inline tts_encode_tiled_fn tc5v2_encoder() {
  static tts_encode_tiled_fn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    TTS_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !p) throw CudaError("cuTensorMapEncodeTiled is not available in this driver");
    fn = reinterpret_cast<tts_encode_tiled_fn>(p);
  }
  return fn;
}

// [rows][cols] f16, row pitch ld elements; box = 64 columns x 128 rows, SWIZZLE_128B, OOB -> 0
static inline CUtensorMap tc5v2_map(const __half *ptr, uint64_t rows, uint64_t cols, uint64_t ld) {
  typedef std::tuple<const void *, uint64_t, uint64_t, uint64_t> Key;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  const Key key(ptr, rows, cols, ld);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  CUtensorMap m;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {ld * 2};
  const cuuint32_t box[2] = {64, 128};
  const cuuint32_t es[2] = {1, 1};
  const CUresult r = tc5v2_encoder()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half *>(ptr), dims, strides,
                                     box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char b[160];
    snprintf(b, sizeof b, "cuTensorMapEncodeTiled failed (%d) for [%llu x %llu] ld %llu", int(r),
             (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
    throw CudaError(b);
  }
  if (cache.size() > 4096) cache.clear();
  cache[key] = m;
  return m;
}

static inline bool tc5v2_supported(const TGemmArgs &g) {
  return g.K % T6_BK == 0 && g.lda % 8 == 0 && (reinterpret_cast<uintptr_t>(g.Ahi) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(g.Whi) & 15) == 0 && (!g.Alo || (reinterpret_cast<uintptr_t>(g.Alo) & 15) == 0) &&
         (!g.Wlo || (reinterpret_cast<uintptr_t>(g.Wlo) & 15) == 0) && g.halo >= g.pad;
}

// returns the number of per-(sequence, group) partial entries written when GroupNorm statistics
// were fused (0 = none)
static inline int launch_tc5v2(const Launcher &L, TGemmArgs g) {
  const bool alo = g.Alo != nullptr, wlo = g.Wlo != nullptr;
  const int nseq = g.M / g.T;
  const int mt = (g.T + T6_BM - 1) / T6_BM;
  // 256-wide tiles once the 128-wide ones no longer fit one wave (single-plane operands only: with hi / lo planes a
  // stage would be 96 KB and the ring two stages deep)
  const bool wide = !alo && !wlo && g.N % 256 == 0 && ((g.N + 127) / 128) * mt * nseq > 148;
  const int BN = wide ? 256 : 128;
  const int nt = (g.N + BN - 1) / BN;
  const int iters = g.taps * (g.K / T6_BK);
  // split-K until the grid covers the 148 SMs once (each rank keeps >= 2 K steps)
  int csz = 1;
  while (!wide && csz < 8 && nt * mt * nseq * csz * 2 <= 148 && iters >= csz * 4) csz *= 2;  // (wide tiles: never split)
  const bool want_gn = g.gn_partial != nullptr && g.N % BN == 0 && mt * csz <= 64;
  if (!want_gn) g.gn_partial = nullptr;
  g.gn_mtiles = mt * csz;
  const size_t stage_bytes = size_t(T6_PLANE) * ((alo ? 2 : 1) + (BN / 128) * (wlo ? 2 : 1));
  const size_t budget = 226 * 1024 - 1024 - T6_CTRL;
  // (capping the ring so that two CTAs fit on an SM -- the next kernel of the chain resident under PDL --
  // measured no gain: 30.4 ms uncapped vs 30.7 / 31.5 ms at 100 / 72 KB for 20 sampling steps)
  int stages = int(budget / stage_bytes);
  if (stages > T6_MAX_STAGES) stages = T6_MAX_STAGES;
  const int per_rank = (iters + csz - 1) / csz;
  if (stages > per_rank) stages = per_rank;
  size_t smem = stages * stage_bytes;
  const size_t red_bytes = size_t(T6_BM) * (BN + 4) * 4;  // partial tile
  if (smem < red_bytes) smem = red_bytes;
  smem += 1024 + T6_CTRL;
  auto kern = wide ? tc5v2_kernel<256> : tc5v2_kernel<128>;
  ensure_smem_attr(kern, 226 * 1024);
  const uint64_t a_rows = uint64_t(nseq) * (g.T + 2 * g.halo);
  const CUtensorMap mAhi = tc5v2_map(g.Ahi, a_rows, g.K, g.lda);
  const CUtensorMap mAlo = alo ? tc5v2_map(g.Alo, a_rows, g.K, g.lda) : mAhi;
  const CUtensorMap mWhi = tc5v2_map(g.Whi, uint64_t(g.taps) * g.N, g.K, g.K);
  const CUtensorMap mWlo = wlo ? tc5v2_map(g.Wlo, uint64_t(g.taps) * g.N, g.K, g.K) : mWhi;

  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(nt, mt, nseq * csz);
  cfg.blockDim = dim3(T6_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = L.stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = L.pdl ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 1;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = csz;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  static int trace_mode = -1;
  static long long *trace_buf = nullptr;
  if (trace_mode < 0) {
    const char *e = getenv("TTS_TC5_TRACE");
    trace_mode = (e && e[0] == '1') ? 1 : 0;
    if (trace_mode) TTS_CUDA_TRY(cudaMalloc(&trace_buf, 16 * sizeof(long long)));
  }
  g.dbg = nullptr;
  if (trace_mode) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(L.stream, &cs);
    if (cs == cudaStreamCaptureStatusNone) {
      TTS_CUDA_TRY(cudaMemsetAsync(trace_buf, 0, 16 * sizeof(long long), L.stream));
      g.dbg = trace_buf;
    }
  }
  TTS_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, g, mAhi, mAlo, mWhi, mWlo, csz, stages));
  if (g.dbg) {
    long long t[16];
    TTS_CUDA_TRY(cudaStreamSynchronize(L.stream));
    TTS_CUDA_TRY(cudaMemcpy(t, trace_buf, sizeof t, cudaMemcpyDeviceToHost));
    fprintf(stderr, "tc5trace M=%d N=%d K=%d taps=%d grid=(%d,%d,%d) csz=%d stages=%d planes=%d%d :", g.M, g.N, g.K, g.taps, nt, mt,
            nseq * csz, csz, stages, alo ? 2 : 1, wlo ? 2 : 1);
    for (int i = 1; i < 13; ++i) fprintf(stderr, " %d:%lld", i, t[i] ? t[i] - t[0] : -1);
    fprintf(stderr, "\n");
  }
  if (L.counter) ++*L.counter;
  return want_gn ? mt * csz : 0;
}

}  // namespace tts
