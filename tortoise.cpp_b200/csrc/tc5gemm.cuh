// tc5gemm.cuh -- tcgen05 (5th-gen tensor core) version of the split-f16 GEMM / implicit conv.
//
// Same contract as tgemm_kernel (gemm.cuh): C = (Ahi+Alo) x (Whi+Wlo)^T (+ taps as shifted K
// steps over a zero-haloed time-major activation buffer), f32 accumulation, fused epilogue.
// Blackwell-native structure:
//   * accumulator lives in TMEM (128 lanes x BN f32 columns, tcgen05.alloc by warp 4);
//   * A / W tiles are staged in shared memory in the canonical K-major SWIZZLE_128B layout
//     (rows of 64 halves = 128 B, 16-byte chunks XOR-swizzled by row & 7, 8-row groups 1024 B
//     apart) by up to 8 loader warps with 16-byte cp.async; each warp owns one pipeline stage
//     (BLOCK_K = 64) and signals it the moment its copies land;
//   * ONE thread issues tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = BN, K = 16) from
//     64-bit shared-memory descriptors; tcgen05.commit releases each stage back to the
//     loaders through an mbarrier and finally signals the epilogue;
//   * epilogue: the 8 loader warps read their 32-lane TMEM quadrant with tcgen05.ld
//     (32x32b.x32), apply bias / activation / residual and store rows of C.
// Operand planes: Ahi.Whi (+ Alo.Whi) (+ Ahi.Wlo), exactly like tgemm_kernel.
#pragma once
#include "gemm.cuh"

namespace tts {

constexpr int T5_BM = 128, T5_BK = 64, T5_MMA_WARP = 8, T5_THREADS = (T5_MMA_WARP + 1) * 32;

__host__ __device__ inline size_t tc5_stage_bytes(int BN, bool alo, bool wlo) {
  return size_t(T5_BM) * 128 * (alo ? 2 : 1) + size_t(BN) * 128 * (wlo ? 2 : 1);
}
__host__ __device__ inline size_t tc5_smem_bytes(int BN, int stages, bool alo, bool wlo) {
  return stages * tc5_stage_bytes(BN, alo, wlo) + 1024 /*alignment slack*/ + 256 /*barriers*/;
}

__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  // K-major, SWIZZLE_128B: LBO = 1 (ignored), SBO = 1024 B between 8-row groups, version 1
  return uint64_t((smem_addr & 0x3FFFF) >> 4) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) |
         (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// one lane of a CONVERGED warp.  tcgen05.mma / TMA take their operands from uniform registers; issued
// under `if (lane == 0)` ptxas cannot assume a single active lane and wraps every instruction in an
// ELECT / R2UR / BRA.U.ANY waterfall loop -- measured ~100-150 cycles of issue time per tcgen05.mma
// (profiles/r02f_dstep_mma_issue_trace.txt).  With elect.sync the whole warp walks the loop, one lane issues.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc5_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc5_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

template <int BN, int T5_STAGES>
static __global__ void __launch_bounds__(T5_THREADS) tc5gemm_kernel(TGemmArgs g) {
  extern __shared__ unsigned char t5_raw[];
  // SWIZZLE_128B atoms need 1024-byte alignment
  unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(t5_raw) + 1023) & ~uintptr_t(1023));
  const bool has_alo = g.Alo != nullptr, has_wlo = g.Wlo != nullptr;
  const uint32_t a_bytes = T5_BM * 128, b_bytes = BN * 128;
  const uint32_t stage_bytes = a_bytes * (has_alo ? 2 : 1) + b_bytes * (has_wlo ? 2 : 1);
  uint64_t *full = reinterpret_cast<uint64_t *>(base + T5_STAGES * stage_bytes);
  uint64_t *empty = full + T5_STAGES;
  uint64_t *done = empty + T5_STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(done + 1);

  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  // M tiles never straddle sequences: tile rows are contiguous rows of ONE sequence of the
  // (halo-padded) activation buffer, so loader addresses advance by a constant stride.
  const int seq = blockIdx.z, t0 = blockIdx.y * T5_BM;
  const int rows_valid = min(T5_BM, g.T - t0);  // rows of this tile that exist
  const int m0 = seq * g.T + t0, n0 = blockIdx.x * BN;
  const int kchunks = g.K / T5_BK;
  const int iters = g.taps * kchunks;

  if (tid == 0) {
    for (int s = 0; s < T5_STAGES; ++s) {
      mbar_init(&full[s], 1);   // one arrival: the loader warp that owns the stage
      mbar_init(&empty[s], 1);  // one arrival: tcgen05.commit
    }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == T5_MMA_WARP) {
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
  pdl_wait();

  auto stage_ptr = [&](int s) { return base + size_t(s) * stage_bytes; };

  if (warp < T5_MMA_WARP) {
    // ================= loader warps: warp w owns pipeline stage w =================
    // Each loader warp fills WHOLE stages (it = w, w + STAGES, ...) and signals the stage's
    // full barrier as soon as its own cp.async group has landed, so the STAGES warps keep
    // STAGES independent stage loads in flight and the MMA thread never waits on a loader
    // that is itself waiting for an earlier MMA (true multi-stage overlap for these small,
    // latency-bound GEMMs).  Lane l moves 16-byte chunk (l & 7) of rows (l >> 3) + 4 j.
    if (warp < T5_STAGES) {
      const int s = warp;
      const int c = lane & 7, r0 = lane >> 3;  // lane moves chunk c of rows r0 + 4 j
      unsigned char *sp = stage_ptr(s);
      unsigned char *wbase = sp + a_bytes * (has_alo ? 2 : 1);
      // (r & 7) alternates between r0 and r0 + 4 as j advances: two swizzled chunk offsets
      const uint32_t sw_even = uint32_t((c ^ r0) * 16), sw_odd = uint32_t((c ^ (r0 + 4)) * 16);
      const size_t a_first = (size_t(seq) * (g.T + 2 * g.halo) + t0 + g.halo - g.pad + r0) * g.lda + c * 8;
      const size_t a_step = size_t(4) * g.lda;
      for (int it = warp, round = 0; it < iters; it += T5_STAGES, ++round) {
        mbar_wait(&empty[s], (round & 1) ^ 1);
        const int tap = it / kchunks, k0 = (it % kchunks) * T5_BK;
        const __half *ap = g.Ahi + a_first + size_t(tap) * g.dil * g.lda + k0;
        const __half *alp = has_alo ? g.Alo + a_first + size_t(tap) * g.dil * g.lda + k0 : nullptr;
        unsigned char *d = sp + r0 * 128;
#pragma unroll 8
        for (int j = 0; j < T5_BM / 4; ++j) {
          const int ok = (r0 + 4 * j) < rows_valid ? 16 : 0;
          unsigned char *dd = d + ((j & 1) ? sw_odd : sw_even);
          cp_async16(dd, ok ? ap : g.Ahi, ok);
          if (has_alo) cp_async16(dd + a_bytes, ok ? alp : g.Alo, ok);
          ap += a_step;
          if (has_alo) alp += a_step;
          d += 4 * 128;
        }
        const size_t w_first = (size_t(tap) * g.N + n0 + r0) * g.K + k0 + c * 8;
        const __half *wp = g.Whi + w_first;
        const __half *wlp = has_wlo ? g.Wlo + w_first : nullptr;
        d = wbase + r0 * 128;
#pragma unroll 8
        for (int j = 0; j < BN / 4; ++j) {
          const int ok = (n0 + r0 + 4 * j) < g.N ? 16 : 0;
          unsigned char *dd = d + ((j & 1) ? sw_odd : sw_even);
          cp_async16(dd, ok ? wp : g.Whi, ok);
          if (has_wlo) cp_async16(dd + b_bytes, ok ? wlp : g.Wlo, ok);
          wp += size_t(4) * g.K;
          if (has_wlo) wlp += size_t(4) * g.K;
          d += 4 * 128;
        }
        cp_async_commit();
        cp_async_wait<0>();
        fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[s]);
      }
    }
    // ================= epilogue (8 warps: TMEM lane quadrant = warp % 4) =================
    mbar_wait(done, 0);
    tc5_fence_after();
    const int q = warp & 3;
    double gs1 = 0.0, gs2 = 0.0;  // fused GroupNorm statistics of this thread's output row
    const int row_in_tile = q * 32 + lane;  // TMEM lane == tile row
    const int m = m0 + row_in_tile;
#pragma unroll
    for (int cb = 0; cb < BN; cb += 32) {
      if (((cb / 32) & 1) != (warp >> 2) && BN > 32) continue;  // BN = 64: warps 0-3 cols 0-31, 4-7 cols 32-63
      if (BN == 32 && warp >= 4) continue;
      float v[32];
      tmem_ld32(tmem_d + (uint32_t(q * 32) << 16) + cb, v);
      if (row_in_tile < rows_valid) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int n = n0 + cb + j;
          if (n + 3 < g.N) {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), o4 = b4;
            if (g.bias) b4 = *reinterpret_cast<const float4 *>(g.bias + n);
            if (g.epi == E_BIAS_RESID || g.epi == E_BIAS_LRELU_RESID)
              o4 = *reinterpret_cast<const float4 *>(g.C + size_t(m) * g.ldc + n);
            float4 r4;
            r4.x = apply_epi(g.epi, v[j], b4.x, o4.x);
            r4.y = apply_epi(g.epi, v[j + 1], b4.y, o4.y);
            r4.z = apply_epi(g.epi, v[j + 2], b4.z, o4.z);
            r4.w = apply_epi(g.epi, v[j + 3], b4.w, o4.w);
            if (g.C) *reinterpret_cast<float4 *>(g.C + size_t(m) * g.ldc + n) = r4;
            gs1 += (double(r4.x) + double(r4.y)) + (double(r4.z) + double(r4.w));
            gs2 += (double(r4.x) * r4.x + double(r4.y) * r4.y) + (double(r4.z) * r4.z + double(r4.w) * r4.w);
            if (g.Chi) {
              const float rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __half hi = __float2half_rn(rr[e]);
                g.Chi[size_t(m) * g.ldh + n + e] = hi;
                if (g.Clo) g.Clo[size_t(m) * g.ldh + n + e] = __float2half_rn(rr[e] - __half2float(hi));
              }
            }
          } else {
            for (int e = 0; e < 4 && n + e < g.N; ++e) {
              const float bias = g.bias ? g.bias[n + e] : 0.f;
              float old = 0.f;
              if (g.epi == E_BIAS_RESID || g.epi == E_BIAS_LRELU_RESID) old = g.C[size_t(m) * g.ldc + n + e];
              const float r = apply_epi(g.epi, v[j + e], bias, old);
              if (g.C) g.C[size_t(m) * g.ldc + n + e] = r;
              if (g.Chi) {
                const __half hi = __float2half_rn(r);
                g.Chi[size_t(m) * g.ldh + n + e] = hi;
                if (g.Clo) g.Clo[size_t(m) * g.ldh + n + e] = __float2half_rn(r - __half2float(hi));
              }
            }
          }
        }
      }
    }
    if (BN == 32 && g.gn_partial && warp < 4) {
      // one N tile == one GroupNorm group (32 channels): reduce {sum, sumsq} over the tile rows
      double *gred = reinterpret_cast<double *>(tmem_slot + 2);
      gs1 = warp_sum_d(gs1);
      gs2 = warp_sum_d(gs2);
      if (lane == 0) { gred[warp * 2] = gs1; gred[warp * 2 + 1] = gs2; }
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (tid == 0) {
        double *o = g.gn_partial + ((size_t(seq) * 32 + blockIdx.x) * g.gn_mtiles + blockIdx.y) * 2;
        o[0] = (gred[0] + gred[2]) + (gred[4] + gred[6]);
        o[1] = (gred[1] + gred[3]) + (gred[5] + gred[7]);
      }
    }
  } else if (lane == 0) {
    // ================= MMA issuer (one thread) =================
    // instruction descriptor: D = F32, A = B = F16, both K-major, N = BN, M = 128
    const uint32_t idesc = (1u << 4) | (uint32_t(BN >> 3) << 17) | (uint32_t(T5_BM >> 4) << 24);
    uint32_t acc = 0;
    for (int j = 0; j < iters; ++j) {
      const int s = j % T5_STAGES;
      mbar_wait(&full[s], (j / T5_STAGES) & 1);
      tc5_fence_after();
      const uint32_t sa = smem_u32(stage_ptr(s));
      const uint32_t sa_lo = sa + a_bytes;
      const uint32_t sw = sa + a_bytes * (has_alo ? 2 : 1);
      const uint32_t sw_lo = sw + b_bytes;
#pragma unroll
      for (int k = 0; k < T5_BK / 16; ++k) {
        const uint32_t koff = k * 32;  // 16 halves = 32 bytes inside the 128-byte swizzled row
        if (has_wlo) { umma_f16(tmem_d, umma_desc_sw128(sa + koff), umma_desc_sw128(sw_lo + koff), idesc, acc); acc = 1; }
        if (has_alo) { umma_f16(tmem_d, umma_desc_sw128(sa_lo + koff), umma_desc_sw128(sw + koff), idesc, acc); acc = 1; }
        umma_f16(tmem_d, umma_desc_sw128(sa + koff), umma_desc_sw128(sw + koff), idesc, acc);
        acc = 1;
      }
      umma_commit(&empty[s]);  // frees this stage when the MMAs above have read it
    }
    umma_commit(done);
  }
  tc5_fence_before();
  __syncthreads();
  if (warp == T5_MMA_WARP) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(uint32_t(BN)) : "memory");
  }
}

}  // namespace tts
