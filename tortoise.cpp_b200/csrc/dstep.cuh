// dstep.cuh -- one diffusion SAMPLING STEP as ONE persistent kernel.
//
// Reference: the loop body of diffusion() (main.cpp:5723-6033): two diffusion_graph runs
// (conditioned + unconditioned, main.cpp:3066-4044) and the DDPM update on the host.  Round 1 ran it as
// ~125 dependent kernels inside a CUDA graph: 1.45 ms per step at S = 191 although the arithmetic is
// ~70 us of tensor time and the weights ~55 us of HBM time -- every kernel of the chain paid launch +
// ramp + drain (profiles/r01d_tc5v2_timeline.txt: 8-14 us per GEMM for 1-2 us of MMA).
// Here the step is a PROGRAM of ops (DOp table in global memory, built once per mel length by
// diffusion.cu) interpreted by 128 resident CTAs; consecutive ops are separated by a device-wide
// barrier (one atomic counter), not by a kernel boundary:
//   GEMM   implicit-conv / matmul on tcgen05: 128 x BN output tiles (BN = 32 .. 128, one tile per
//          CTA at S = 191), operands by TMA (SWIZZLE_128B boxes), accumulator in TMEM, NO split-K:
//          the CTAs of a cluster of 4 own neighbouring N tiles of the same M tile and share its A
//          operand by TMA MULTICAST (each CTA loads a quarter of the rows and the hardware copies it
//          into all four shared memories), which is what the split-K + DSMEM reduction of tc5v2.cuh
//          bought (L2 -> SM traffic of a 128 x 128 tile) without the reduction;
//          fused epilogue: bias, residual from a second pointer, GroupNorm statistics of the output;
//   GN     GroupNorm apply + affine + (1 + scale) / shift + SiLU -> f16 conv operand with halo rows;
//   ATTN   self-attention with the T5 bucket bias (SIMT f32, same body as diff_attn_kernel);
//   XIN / CONCAT / DDPM  the small glue kernels of diff_kernels.cuh.
// Numerics are those of the per-op kernels (same operand planes, f32 accumulation, double GroupNorm
// sums); tests/test_diffusion_gpu.py pins both paths against the reference.
#pragma once
#include <cuda.h>

#include "diff_kernels.cuh"
#include "tc5v2.cuh"

namespace tts {

constexpr int DS_THREADS = 256;
constexpr int DS_CLUSTER = 4;           // CTAs sharing an A tile by multicast
constexpr int DS_GRID = 128;            // resident CTAs (132 can be co-resident with clusters of 4)
constexpr int DS_BM = 128, DS_BK = 64;
constexpr int DS_A_BYTES = DS_BM * 128;  // one 128-row x 64-half swizzled box
constexpr int DS_RING_BYTES = 192 * 1024;
constexpr int DS_CTRL_BYTES = 1024;
constexpr int DS_MAX_STAGES = 8;
constexpr size_t DS_SMEM = DS_RING_BYTES + DS_CTRL_BYTES + 1024 /*alignment*/;

enum DOpKind { D_GEMM = 0, D_GN = 1, D_ATTN = 2, D_XIN = 3, D_CONCAT = 4, D_DDPM = 5 };

struct DOp {
  int kind;
  // ---- GEMM: C[nseq*T][N] = epi(A x W^T); A rows = time-major sequences with halo (TGemmArgs contract)
  int mA_hi, mA_lo, mW_hi, mW_lo;  // tensor-map table indices (-1: plane absent)
  int T, nseq, N, K, taps, halo, bn, ldc, epi, stages;
  const float *bias;
  float *C;
  const float *R;       // residual source for E_BIAS_RESID (may equal C)
  double *gn_out;       // fused GroupNorm statistics of the output (or null); needs bn == 32 groups-aligned tiles
  // ---- GN apply
  const float *X;       // [nseq][T][1024]
  const float *gw, *gb; // affine
  const float *ss;      // scale|shift table (or null); + step * ss_step_stride when per_step
  int ss_step_stride, silu;
  const double *gn_in;  // {sum, sumsq} partials [(seq*32+g)*gn_mtiles + i] (or null -> stats)
  int gn_mtiles;
  const float *stats;   // precomputed {mean, rstd} per (seq, group) when gn_in is null
  __half *out16;        // [nseq][T + 2][1024]
  // ---- ATTN
  const float *QKV, *relbias;
  const int *rpb;
  __half *att_hi, *att_lo;
  // ---- XIN / CONCAT / DDPM
  const float *x;       // [100][S]
  __half *xin16;        // [S + 2][128]
  const float *INP, *CW;
  __half *cat16;        // [nseq][S + 2][2048]
  float *xw;            // x (updated in place by DDPM)
  const float *OUT, *noise;
  const DdpmCoef *coefs;
};

struct DStepArgs {
  const DOp *ops;
  int n_ops;
  const CUtensorMap *maps;   // global memory, 64-byte aligned
  unsigned int *bar;         // device-wide barrier counter (monotonic; reset per utterance)
  int *step;                 // sampling-step counter (read at entry, incremented by CTA 0 at exit)
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d_mc(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}

// device-wide barrier between two ops.  Every thread's global writes of the finished op must be visible
// to every CTA -- including to TMA (async proxy) reads of the next op -- before anyone proceeds.
__device__ __forceinline__ void ds_grid_barrier(unsigned int *bar, unsigned int target) {
  __threadfence();
  fence_proxy_async_all();
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (int(v - target) < 0);
    fence_proxy_async_all();
  }
  __syncthreads();
}

struct DsPipe {  // running pipeline state of the GEMM roles (persists across ops)
  // The stage count / size changes from op to op while the mbarrier phases keep counting, so the phase
  // of a slot is tracked per slot: bit s = parity of the number of times slot s has been used so far
  // (producer and MMA thread each keep their own copy; every op starts on slot 0 of a drained ring).
  uint32_t prod_par = 0, mma_par = 0;
  uint32_t items = 0;  // tiles finished by this CTA (accumulator barrier phases)
};

// One GEMM op.  Cluster c of `ncl` clusters takes tile groups c, c + ncl, ...: a tile group = one M tile
// (128 rows of one sequence) x DS_CLUSTER neighbouring N tiles, one per CTA of the cluster.
__device__ __forceinline__ void ds_gemm(const DOp &op, const CUtensorMap *maps, unsigned char *ring, uint64_t *full,
                                        uint64_t *empty, uint64_t *accf, uint64_t *acce, uint32_t tmem_d, double *gw,
                                        DsPipe &ps, int step) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t crank = cluster_ctarank();
  const int cluster = blockIdx.x / DS_CLUSTER, ncl = gridDim.x / DS_CLUSTER;
  const bool has_alo = op.mA_lo >= 0, has_wlo = op.mW_lo >= 0;
  const int a_planes = has_alo ? 2 : 1, w_planes = has_wlo ? 2 : 1;
  const int bn = op.bn;
  const uint32_t w_bytes = uint32_t(bn) * 128u;
  const uint32_t stage_bytes = DS_A_BYTES * a_planes + w_bytes * w_planes;
  const int stages = op.stages;
  const int mt = (op.T + DS_BM - 1) / DS_BM;          // M tiles per sequence
  const int ntg = (op.N + bn * DS_CLUSTER - 1) / (bn * DS_CLUSTER);  // N tile groups
  const int n_groups = op.nseq * mt * ntg;
  const int kchunks = op.K / DS_BK, iters = op.taps * kchunks;
  const int pad = op.taps / 2;
  const CUtensorMap *mAhi = maps + op.mA_hi, *mAlo = has_alo ? maps + op.mA_lo : nullptr;
  const CUtensorMap *mWhi = maps + op.mW_hi, *mWlo = has_wlo ? maps + op.mW_lo : nullptr;
  const uint16_t mc_mask = uint16_t((1u << DS_CLUSTER) - 1);

  if (warp == 0) {
    if (lane == 0) {
      // ================= TMA producer =================
      prefetch_tmap(mAhi);
      prefetch_tmap(mWhi);
      uint32_t slot_it = 0;
      for (int grp = cluster; grp < n_groups; grp += ncl) {
        const int ng = grp % ntg, mi = (grp / ntg) % mt, seq = grp / (ntg * mt);
        const int t0 = mi * DS_BM;
        const int n0 = (ng * DS_CLUSTER + int(crank)) * bn;
        const int a_row0 = seq * (op.T + 2 * op.halo) + t0 + op.halo - pad;
        for (int it = 0; it < iters; ++it, ++slot_it) {
          const uint32_t s = slot_it % stages;
          // the slot is free once ALL CTAs of the cluster have consumed it (peers write into it too)
          mbar_wait(&empty[s], ((ps.prod_par >> s) & 1u) ^ 1u);
          ps.prod_par ^= 1u << s;
          mbar_arrive_expect_tx(&full[s], stage_bytes);
          const int tap = it / kchunks, k0 = (it % kchunks) * DS_BK;
          unsigned char *sp = ring + size_t(s) * stage_bytes;
          // weights: this CTA's own N tile
          tma_load_2d(sp + DS_A_BYTES * a_planes, mWhi, k0, tap * op.N + n0, &full[s]);
          if (has_wlo) tma_load_2d(sp + DS_A_BYTES * a_planes + w_bytes, mWlo, k0, tap * op.N + n0, &full[s]);
          // activations: rows [32 r, 32 r + 32) of the shared A tile, multicast to the whole cluster
          const int qrows = DS_BM / DS_CLUSTER;
          tma_load_2d_mc(sp + crank * qrows * 128, mAhi, k0, a_row0 + tap + int(crank) * qrows, &full[s], mc_mask);
          if (has_alo)
            tma_load_2d_mc(sp + DS_A_BYTES + crank * qrows * 128, mAlo, k0, a_row0 + tap + int(crank) * qrows, &full[s], mc_mask);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ================= MMA issuer =================
      const uint32_t idesc = (1u << 4) | (uint32_t(bn >> 3) << 17) | (uint32_t(DS_BM >> 4) << 24);
      uint32_t items = ps.items, slot_it = 0;
      for (int grp = cluster; grp < n_groups; grp += ncl, ++items) {
        // the epilogue warps must have drained the accumulator of the previous tile
        mbar_wait(acce, (items & 1u) ^ 1u);
        tc5_fence_after();
        uint32_t acc = 0;
        for (int it = 0; it < iters; ++it, ++slot_it) {
          const uint32_t s = slot_it % stages;
          mbar_wait(&full[s], (ps.mma_par >> s) & 1u);
          ps.mma_par ^= 1u << s;
          tc5_fence_after();
          const uint32_t sa = smem_u32(ring + size_t(s) * stage_bytes);
          const uint32_t sa_lo = sa + DS_A_BYTES;
          const uint32_t sw = sa + DS_A_BYTES * a_planes;
          const uint32_t sw_lo = sw + w_bytes;
#pragma unroll
          for (int k = 0; k < DS_BK / 16; ++k) {
            const uint32_t koff = k * 32;
            if (has_wlo) { umma_f16(tmem_d, umma_desc_sw128(sa + koff), umma_desc_sw128(sw_lo + koff), idesc, acc); acc = 1; }
            if (has_alo) { umma_f16(tmem_d, umma_desc_sw128(sa_lo + koff), umma_desc_sw128(sw + koff), idesc, acc); acc = 1; }
            umma_f16(tmem_d, umma_desc_sw128(sa + koff), umma_desc_sw128(sw + koff), idesc, acc);
            acc = 1;
          }
          umma_commit_mc(&empty[s], mc_mask);  // frees the slot in every CTA of the cluster
        }
        umma_commit(accf);
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue: TMEM -> registers -> global, one output row per thread =================
    const int q = warp & 3;  // TMEM lane quadrant of this warp
    const int r = q * 32 + lane;
    uint32_t items = ps.items;
    const bool resid = op.epi == E_BIAS_RESID;
    for (int grp = cluster; grp < n_groups; grp += ncl, ++items) {
      const int ng = grp % ntg, mi = (grp / ntg) % mt, seq = grp / (ntg * mt);
      const int t0 = mi * DS_BM;
      const int n0 = (ng * DS_CLUSTER + int(crank)) * bn;
      const int rows_valid = min(DS_BM, op.T - t0);
      const bool valid = r < rows_valid;
      const size_t m = size_t(seq) * op.T + t0 + r;
      mbar_wait(accf, items & 1u);
      tc5_fence_after();
      for (int cb = 0; cb < bn; cb += 32) {
        float v[32];
        tmem_ld32(tmem_d + (uint32_t(q * 32) << 16) + cb, v);
        const int n = n0 + cb;
        double s1 = 0.0, s2 = 0.0;
        if (valid && n < op.N) {
          float *crow = op.C + m * op.ldc + n;
          const float *rrow = resid ? op.R + m * op.ldc + n : nullptr;
          if (n + 32 <= op.N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = *reinterpret_cast<const float4 *>(op.bias + n + j);
              float4 o = make_float4(v[j] + b4.x, v[j + 1] + b4.y, v[j + 2] + b4.z, v[j + 3] + b4.w);
              if (resid) {
                const float4 old = *reinterpret_cast<const float4 *>(rrow + j);
                o.x = old.x + o.x; o.y = old.y + o.y; o.z = old.z + o.z; o.w = old.w + o.w;
              }
              *reinterpret_cast<float4 *>(crow + j) = o;
              s1 += (double(o.x) + double(o.y)) + (double(o.z) + double(o.w));
              s2 += (double(o.x) * o.x + double(o.y) * o.y) + (double(o.z) * o.z + double(o.w) * o.w);
            }
          } else {
            for (int j = 0; j < 32 && n + j < op.N; ++j) {
              float o = v[j] + op.bias[n + j];
              if (resid) o = rrow[j] + o;
              crow[j] = o;
            }
          }
        }
        if (op.gn_out) {  // one 32-column block = one GroupNorm group: reduce over the tile's rows
          s1 = warp_sum_d(s1);
          s2 = warp_sum_d(s2);
          if (lane == 0) { gw[q * 2] = s1; gw[q * 2 + 1] = s2; }
          asm volatile("bar.sync 2, 128;" ::: "memory");
          if (warp == 4 && lane == 0) {
            double *o = op.gn_out + ((size_t(seq) * 32 + (n >> 5)) * mt + mi) * 2;
            o[0] = (gw[0] + gw[2]) + (gw[4] + gw[6]);
            o[1] = (gw[1] + gw[3]) + (gw[5] + gw[7]);
          }
          asm volatile("bar.sync 2, 128;" ::: "memory");
        }
      }
      tc5_fence_before();
      mbar_arrive(acce);  // 128 arrivals: the accumulator may be overwritten
    }
  }
  __syncwarp();
  // every role walked the same tile list
  int my = 0;
  for (int grp = cluster; grp < n_groups; grp += ncl) ++my;
  ps.items += my;
  (void)step;
}

// GroupNorm apply (body of gn_apply_kernel): rows strided over the CTAs
__device__ __forceinline__ void ds_gn(const DOp &op, int step) {
  const int tid = threadIdx.x, T = op.T, halo = 1;
  const float *ss = op.ss ? op.ss + size_t(step) * op.ss_step_stride : nullptr;
  const int c = tid * 4, g = c >> 5;
  const float4 w4 = *reinterpret_cast<const float4 *>(op.gw + c);
  const float4 b4 = *reinterpret_cast<const float4 *>(op.gb + c);
  float4 sc = make_float4(0.f, 0.f, 0.f, 0.f), sh = sc;
  if (ss) {
    sc = *reinterpret_cast<const float4 *>(ss + c);
    sh = *reinterpret_cast<const float4 *>(ss + kDim + c);
  }
  for (int seq = 0; seq < op.nseq; ++seq) {
    float mean, rstd;
    if (op.gn_in) {
      double s1 = 0.0, s2 = 0.0;
      for (int i = 0; i < op.gn_mtiles; ++i) {
        s1 += op.gn_in[((size_t(seq) * 32 + g) * op.gn_mtiles + i) * 2];
        s2 += op.gn_in[((size_t(seq) * 32 + g) * op.gn_mtiles + i) * 2 + 1];
      }
      const double n = double(T) * 32.0, md = s1 / n;
      mean = float(md);
      double var = s2 / n - 2.0 * md * double(mean) + double(mean) * double(mean);  // E[(x - mean_f)^2]
      if (var < 0.0) var = 0.0;
      rstd = 1.0f / sqrtf(float(var) + 1e-6f);
    } else {
      mean = op.stats[(seq * 32 + g) * 2];
      rstd = op.stats[(seq * 32 + g) * 2 + 1];
    }
    for (int row = blockIdx.x; row < T + 2 * halo; row += gridDim.x) {
      const int t = row - halo;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (t >= 0 && t < T) {
        const float4 x = *reinterpret_cast<const float4 *>(op.X + (size_t(seq) * T + t) * kDim + c);
        v[0] = (x.x - mean) * rstd * w4.x + b4.x;
        v[1] = (x.y - mean) * rstd * w4.y + b4.y;
        v[2] = (x.z - mean) * rstd * w4.z + b4.z;
        v[3] = (x.w - mean) * rstd * w4.w + b4.w;
        if (ss) {
          v[0] = v[0] * (sc.x + 1.0f) + sh.x;
          v[1] = v[1] * (sc.y + 1.0f) + sh.y;
          v[2] = v[2] * (sc.z + 1.0f) + sh.z;
          v[3] = v[3] * (sc.w + 1.0f) + sh.w;
        }
        if (op.silu) {
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = silu_f(v[i]);
        }
      }
      __half h[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) h[i] = __float2half_rn(v[i]);
      *reinterpret_cast<uint2 *>(op.out16 + (size_t(seq) * (T + 2 * halo) + row) * kDim + c) = *reinterpret_cast<uint2 *>(h);
    }
  }
}

// self-attention items (body of diff_attn_kernel): (query tile, head, sequence) strided over the CTAs
__device__ __forceinline__ void ds_attn(const DOp &op, float *da_smem) {
  constexpr int TK = DA_TK, LDK = DA_LDK;
  float (*Ks)[LDK] = reinterpret_cast<float (*)[LDK]>(da_smem);
  float (*Vs)[LDK] = reinterpret_cast<float (*)[LDK]>(da_smem + TK * LDK);
  float (*Qs)[kHeadDim] = reinterpret_cast<float (*)[kHeadDim]>(da_smem + 2 * TK * LDK);
  float (*Ps)[4][TK] = reinterpret_cast<float (*)[4][TK]>(da_smem + 2 * TK * LDK + DA_Q * kHeadDim);
  float *bias_s = da_smem + 2 * TK * LDK + DA_Q * kHeadDim + DA_WARPS * 4 * TK;
  const int t = threadIdx.x, warp = t / 32, lane = t % 32;
  const int T = op.T;
  const int qt = (T + DA_Q - 1) / DA_Q, n_items = qt * kHeads * op.nseq;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int qi0 = item % qt, head = (item / qt) % kHeads, seq = item / (qt * kHeads);
    const int q0 = qi0 * DA_Q;
    const float *base = op.QKV + size_t(seq) * T * 3072 + head * 192;
    __syncthreads();  // the previous item's shared tiles are dead
    if (t < 32) bias_s[t] = 8.0f * op.relbias[t * 16 + head];
    for (int i = t; i < DA_Q * kHeadDim; i += DA_THREADS) {
      const int r = i / kHeadDim, d = i % kHeadDim;
      const int qi = q0 + r;
      Qs[r][d] = qi < T ? base[size_t(qi) * 3072 + d] : 0.f;
    }
    float m[4], l[4], o[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      m[i] = -INFINITY;
      l[i] = 0.f;
      o[i][0] = o[i][1] = 0.f;
    }
    for (int k0 = 0; k0 < T; k0 += TK) {
      __syncthreads();
      for (int i = t; i < TK * (kHeadDim / 4); i += DA_THREADS) {
        const int r = i / (kHeadDim / 4), c = (i % (kHeadDim / 4)) * 4;
        const int kj = k0 + r;
        float4 kv = make_float4(0, 0, 0, 0), vv = kv;
        if (kj < T) {
          kv = *reinterpret_cast<const float4 *>(base + size_t(kj) * 3072 + 64 + c);
          vv = *reinterpret_cast<const float4 *>(base + size_t(kj) * 3072 + 128 + c);
        }
        *reinterpret_cast<float4 *>(&Ks[r][c]) = kv;
        *reinterpret_cast<float4 *>(&Vs[r][c]) = vv;
      }
      __syncthreads();
      float s[4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) s[i][0] = s[i][1] = 0.f;
#pragma unroll 4
      for (int c = 0; c < kHeadDim; c += 4) {
        const float4 ka = *reinterpret_cast<const float4 *>(&Ks[lane][c]);
        const float4 kb = *reinterpret_cast<const float4 *>(&Ks[lane + 32][c]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 qv = *reinterpret_cast<const float4 *>(&Qs[warp * 4 + i][c]);
          s[i][0] += qv.x * ka.x + qv.y * ka.y + qv.z * ka.z + qv.w * ka.w;
          s[i][1] += qv.x * kb.x + qv.y * kb.y + qv.z * kb.z + qv.w * kb.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int qi = q0 + warp * 4 + i;
        const int j0 = k0 + lane, j1 = k0 + lane + 32;
        float s0 = -INFINITY, s1 = -INFINITY;
        if (qi < T && j0 < T) s0 = s[i][0] * 0.125f + bias_s[(j0 > qi ? 16 : 0) + op.rpb[abs(j0 - qi)]];
        if (qi < T && j1 < T) s1 = s[i][1] * 0.125f + bias_s[(j1 > qi ? 16 : 0) + op.rpb[abs(j1 - qi)]];
        const float tmax = warp_max(fmaxf(s0, s1));
        const float mnew = fmaxf(m[i], tmax);
        float p0 = 0.f, p1 = 0.f, corr = 1.f;
        if (mnew != -INFINITY) {
          p0 = expf(s0 - mnew);
          p1 = expf(s1 - mnew);
          corr = expf(m[i] - mnew);
        }
        l[i] = l[i] * corr + warp_sum(p0 + p1);
        o[i][0] *= corr;
        o[i][1] *= corr;
        m[i] = mnew;
        Ps[warp][i][lane] = p0;
        Ps[warp][i][lane + 32] = p1;
      }
      __syncwarp();
      const int kmax = min(TK, T - k0);
      for (int j = 0; j < kmax; ++j) {
        const float v0 = Vs[j][lane], v1 = Vs[j][lane + 32];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float p = Ps[warp][i][j];
          o[i][0] = fmaf(p, v0, o[i][0]);
          o[i][1] = fmaf(p, v1, o[i][1]);
        }
      }
      __syncwarp();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qi = q0 + warp * 4 + i;
      if (qi < T) {
        const float inv = 1.0f / l[i];
        const size_t off = (size_t(seq) * T + qi) * kDim + head * kHeadDim;
        const float v0 = o[i][0] * inv, v1 = o[i][1] * inv;
        const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
        op.att_hi[off + lane] = h0;
        op.att_hi[off + lane + 32] = h1;
        op.att_lo[off + lane] = __float2half_rn(v0 - __half2float(h0));
        op.att_lo[off + lane + 32] = __float2half_rn(v1 - __half2float(h1));
      }
    }
  }
}

__device__ __forceinline__ void ds_small(const DOp &op, int step) {
  const int tid = threadIdx.x, S = op.T;
  if (op.kind == D_XIN) {  // x [100][S] -> f16 [S + 2][128]
    for (int i = blockIdx.x * DS_THREADS + tid; i < (S + 2) * 128; i += gridDim.x * DS_THREADS) {
      const int row = i >> 7, c = i & 127, t = row - 1;
      float v = 0.f;
      if (t >= 0 && t < S && c < 100) v = op.x[size_t(c) * S + t];
      op.xin16[i] = __float2half_rn(v);
    }
  } else if (op.kind == D_CONCAT) {  // [INP | CW] -> f16 with halo rows
    for (int row = blockIdx.x; row < op.nseq * (S + 2); row += gridDim.x) {
      const int seq = row / (S + 2), t = row % (S + 2) - 1;
      __half *o = op.cat16 + size_t(row) * 2048;
      for (int c = tid; c < 2048; c += DS_THREADS) {
        float v = 0.f;
        if (t >= 0 && t < S) v = c < 1024 ? op.INP[size_t(t) * kDim + c] : op.CW[(size_t(seq) * S + t) * kDim + c - 1024];
        o[c] = __float2half_rn(v);
      }
    }
  } else {  // D_DDPM (body of ddpm_step_kernel)
    const DdpmCoef k = op.coefs[step];
    const int n = 100 * S;
    const float *noise = op.noise + size_t(step + 1) * n;
    for (int i = blockIdx.x * DS_THREADS + tid; i < n; i += gridDim.x * DS_THREADS) {
      const int ch = i / S, t = i % S;
      const float eps_c = op.OUT[size_t(t) * 200 + ch];
      const float vraw = op.OUT[size_t(t) * 200 + 100 + ch];
      const float eps_u = op.OUT[(size_t(S) + t) * 200 + ch];
      const float frac = __fdiv_rn(__fadd_rn(vraw, 1.0f), 2.0f);
      const float logvar = __fadd_rn(__fmul_rn(frac, k.min_log), __fmul_rn(__fsub_rn(1.0f, frac), k.max_log));
      const float eps = __fsub_rn(__fmul_rn(__fadd_rn(1.0f, k.cfk), eps_c), __fmul_rn(k.cfk, eps_u));
      const float xv = op.xw[i];
      float x0 = __fsub_rn(__fmul_rn(k.sqrt_recip, xv), __fmul_rn(k.sqrt_recipm1, eps));
      x0 = fminf(1.0f, fmaxf(-1.0f, x0));
      const float mean = __fadd_rn(__fmul_rn(k.coef1, x0), __fmul_rn(k.coef2, xv));
      float r = mean;
      if (!k.last) r = float(__dadd_rn(double(mean), __dmul_rn(exp(__dmul_rn(0.5, double(logvar))), double(noise[i]))));
      op.xw[i] = r;
    }
  }
}

static __global__ void __cluster_dims__(DS_CLUSTER, 1, 1) __launch_bounds__(DS_THREADS, 1) dstep_kernel(DStepArgs a) {
  extern __shared__ unsigned char ds_raw[];
  unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(ds_raw) + 1023) & ~uintptr_t(1023));
  unsigned char *ring = base;
  unsigned char *ctrl = base + DS_RING_BYTES;
  uint64_t *full = reinterpret_cast<uint64_t *>(ctrl);
  uint64_t *empty = full + DS_MAX_STAGES;
  uint64_t *accf = empty + DS_MAX_STAGES;
  uint64_t *acce = accf + 1;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acce + 1);
  double *gw = reinterpret_cast<double *>(ctrl + 256);
  __shared__ DOp s_op;

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < DS_MAX_STAGES; ++s) {
      mbar_init(&full[s], 1);             // this CTA's producer (expect_tx); bytes arrive from 4 CTAs
      mbar_init(&empty[s], DS_CLUSTER);   // one tcgen05.commit per CTA of the cluster
    }
    mbar_init(accf, 1);
    mbar_init(acce, 128);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc5_fence_before();
  cluster_sync_all();  // barriers initialised cluster-wide before any peer multicasts into this CTA
  tc5_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  const int step = *a.step;
  const unsigned int bar_base = (unsigned int)step * (unsigned int)a.n_ops * gridDim.x;
  DsPipe ps;
  for (int i = 0; i < a.n_ops; ++i) {
    if (tid < int(sizeof(DOp) / 4)) reinterpret_cast<uint32_t *>(&s_op)[tid] = reinterpret_cast<const uint32_t *>(a.ops + i)[tid];
    __syncthreads();
    const DOp &op = s_op;
    switch (op.kind) {
      case D_GEMM: ds_gemm(op, a.maps, ring, full, empty, accf, acce, tmem_d, gw, ps, step); break;
      case D_GN: ds_gn(op, step); break;
      case D_ATTN: ds_attn(op, reinterpret_cast<float *>(ring)); break;
      default: ds_small(op, step); break;
    }
    ds_grid_barrier(a.bar, bar_base + (unsigned int)(i + 1) * gridDim.x);
  }
  if (blockIdx.x == 0 && tid == 0) *a.step = step + 1;
  tc5_fence_before();
  cluster_sync_all();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(128u) : "memory");
}

}  // namespace tts
