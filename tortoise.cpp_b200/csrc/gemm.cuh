// gemm.cuh -- dense f32-accumulate GEMMs used off the decode hot loop.
//
//  (1) sgemm_tn_kernel: C[M][N] = A[M][K] (f32) x W[N][K]^T (f32 or f16 storage), f32 FMA.
//      Used where the reference multiplies in F32 (ggml_mul_mat with F32 operands): the AR
//      prefill / latent pass (main.cpp:2053-2519), diffusion proj_out / emb_layers / time
//      MLP (main.cpp:3331-3343, 3592-3596).
//  (2) tgemm_kernel (HMMA tensor cores, split-f16 operands, f32 accumulate): the F32 x F32
//      matmuls of the AR prefill / latent pass and of diffusion (operands carried as hi/lo
//      f16 planes), AND every ggml_conv_1d of the diffusion / vocoder graphs, whose reference
//      numerics are F16 im2col x F16 weights with F32 accumulation (ggml.c:6493-6508,
//      15167-15248) -- run as an implicit GEMM (taps = shifted K steps over a zero-haloed
//      time-major activation buffer; no im2col buffer).
#pragma once
#include <mma.h>

#include "common.cuh"

namespace tts {

enum Epi {
  E_NONE = 0,
  E_BIAS = 1,         // + bias[n]
  E_BIAS_H16 = 2,     // h16(+bias)            (AR qkv, A-2)
  E_BIAS_GELU16 = 3,  // gelu16(+bias)         (AR fc)
  E_BIAS_RESID = 4,   // C += acc + bias       (AR c_proj / mlp c_proj, diffusion proj_out)
  E_BIAS_LRELU = 5,   // leaky_relu(+bias, 0.2)
  E_BIAS_LRELU_RESID = 6,  // C += leaky_relu(acc + bias, 0.2)   (vocoder kernel-predictor residual)
};

struct GemmArgs {
  const float *A;   // [M][lda]
  const void *W;    // [N][K]
  const float *bias;
  float *C;         // [M][ldc]
  int M, N, K, lda, ldc;
  int epi;
};

__device__ __forceinline__ float apply_epi(int epi, float acc, float bias, float old) {
  switch (epi) {
    case E_NONE: return acc;
    case E_BIAS: return acc + bias;
    case E_BIAS_H16: return h16(acc + bias);
    case E_BIAS_GELU16: return gelu16(acc + bias);
    case E_BIAS_RESID: return old + (acc + bias);
    case E_BIAS_LRELU: {
      const float v = acc + bias;
      return v > 0.f ? v : 0.2f * v;
    }
    case E_BIAS_LRELU_RESID: {
      const float v = acc + bias;
      return old + (v > 0.f ? v : 0.2f * v);
    }
  }
  return acc;
}

// ---- (1) SIMT f32 GEMM, 128x128x16 tiles, 256 threads, 8x8 micro-tiles -------------------
template <typename WT>
static __global__ void __launch_bounds__(256) sgemm_tn_kernel(GemmArgs g) {
  constexpr int BM = 128, BN = 128, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = tid % 16, ty = tid / 16;  // 16x16 threads, each 8x8 outputs (strided by 16)
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // loader mapping: 256 threads x 8 elements = 128 rows x 16 k
  const int lr = tid / 2;         // 0..127 row in tile
  const int lk = (tid % 2) * 8;   // 0 or 8
  for (int k0 = 0; k0 < g.K; k0 += BK) {
    {
      float v[8];
      const int m = m0 + lr;
      if (m < g.M) {
        const float4 a0 = *reinterpret_cast<const float4 *>(g.A + size_t(m) * g.lda + k0 + lk);
        const float4 a1 = *reinterpret_cast<const float4 *>(g.A + size_t(m) * g.lda + k0 + lk + 4);
        v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w;
        v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) As[lk + e][lr] = v[e];
    }
    {
      float v[8];
      const int n = n0 + lr;
      if (n < g.N) {
        if constexpr (sizeof(WT) == 4) {
          const float *w = reinterpret_cast<const float *>(g.W) + size_t(n) * g.K + k0 + lk;
          const float4 a0 = *reinterpret_cast<const float4 *>(w);
          const float4 a1 = *reinterpret_cast<const float4 *>(w + 4);
          v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w;
          v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
        } else {
          const __half *w = reinterpret_cast<const __half *>(g.W) + size_t(n) * g.K + k0 + lk;
          const uint4 u = *reinterpret_cast<const uint4 *>(w);
          const __half2 *h2 = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = __half22float2(h2[q]);
            v[2 * q] = f.x;
            v[2 * q + 1] = f.y;
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) Bs[lk + e][lr] = v[e];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty + 16 * i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n >= g.N) continue;
      float *c = g.C + size_t(m) * g.ldc + n;
      const float bias = g.bias ? g.bias[n] : 0.f;
      const float old = (g.epi == E_BIAS_RESID || g.epi == E_BIAS_LRELU_RESID) ? *c : 0.f;
      *c = apply_epi(g.epi, acc[i][j], bias, old);
    }
  }
}

// ---- (1b) split-f16 tensor-core GEMM ---------------------------------------------------------
// C[M][N] = (Ahi + Alo)[M][K] x (Whi + Wlo)[N][K]^T with f32 accumulation on the tensor cores:
// an f32 operand x is carried as two f16 planes hi = f16(x), lo = f16(x - hi) (relative error
// ~2^-22), so F32 x F32 reference matmuls (AR prefill / latent pass, diffusion proj_out) run on
// tensor cores without changing their numerics beyond f32 rounding noise; operands that are
// f16-exact in the reference (gelu16 outputs, f16 weights in fast mode) pass a null lo plane.
// Products issued: Ahi.Whi (+ Alo.Whi) (+ Ahi.Wlo); the lo.lo term (2^-24) is dropped.
// 64x64x32 CTA tile, 4 warps (2x2, 32x32 each), 3-stage cp.async pipeline.
struct TGemmArgs {
  const __half *Ahi, *Alo;  // [M][lda]; Alo may be null
  const __half *Whi, *Wlo;  // [N][K];   Wlo may be null
  const float *bias;        // [N] or null
  float *C;                 // [M][ldc] f32 out (may be null); E_BIAS_RESID accumulates into it
  __half *Chi, *Clo;        // [M][ldh] f16 planes of the result (may be null)
  int M, N, K, lda, ldc, ldh;
  int epi;
  // implicit 1-D convolution (reference: ggml_conv_1d = F16 im2col x F16 kernel, F32 accumulate,
  // ggml.c:6493-6508): taps > 1 turns the K loop into taps x (K/32) steps; A rows are addressed
  // as time-major sequences [seq][T + 2*halo][lda] with zero halo rows, row m = seq*T + t reads
  // A row seq*(T+2*halo) + t + halo + tap*dil - pad; W is [taps][N][K].  Plain GEMM: taps = 1,
  // T = M, halo = pad = 0.
  int taps, dil, pad, halo, T;
  // optional fused GroupNorm statistics of the OUTPUT (tcgen05 path, BN = 32 = one 32-channel
  // group per N tile): per (sequence, group, M tile) {sum, sum of squares} in double.
  double *gn_partial;
  int gn_mtiles;
  long long *dbg;  // optional per-launch clock trace (TTS_TC5_TRACE=1), else null
  // optional per-sequence valid lengths (utterance batching: sequences of different lengths share the row
  // stride T; rows >= Tseq[seq] are padding -- neither stored nor counted in the GroupNorm statistics).
  // tcgen05 path (tc5v2.cuh) only.
  const int *Tseq;
};

constexpr int TG_BM = 64, TG_BN = 64, TG_BK = 32, TG_LD = TG_BK + 8, TG_STAGES = 3;
constexpr int TG_PLANE = TG_BM * TG_LD;  // halves per operand plane per stage
__host__ __device__ inline size_t tgemm_smem_bytes() { return size_t(TG_STAGES) * 4 * TG_PLANE * sizeof(__half); }

__device__ __forceinline__ void cp_async16(void *dst, const void *src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

static __global__ void __launch_bounds__(128) tgemm_kernel(TGemmArgs g) {
  using namespace nvcuda;
  extern __shared__ __align__(128) unsigned char tg_smem[];
  __half *sm = reinterpret_cast<__half *>(tg_smem);
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const int wm = warp / 2, wn = warp % 2;
  const int m0 = blockIdx.y * TG_BM, n0 = blockIdx.x * TG_BN;
  const bool has_alo = g.Alo != nullptr, has_wlo = g.Wlo != nullptr;

  auto plane = [&](int stage, int which) { return sm + (size_t(stage) * 4 + which) * TG_PLANE; };
  // each thread moves 2 x 16 B per plane per stage: rows tid/4 and tid/4+32, chunk (tid%4)*8 halves
  const int kchunks = g.K / TG_BK;
  // A row base (in rows) of this thread's two tile rows, before the per-tap shift
  long arow[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int mc = min(m0 + tid / 4 + h * 32, g.M - 1);
    const int seq = mc / g.T, t = mc % g.T;
    arow[h] = long(seq) * (g.T + 2 * g.halo) + t + g.halo - g.pad;
  }
  auto load_stage = [&](int stage, int it) {
    const int tap = it / kchunks, k0 = (it % kchunks) * TG_BK;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = tid / 4 + h * 32, c = (tid % 4) * 8;
      const int m = m0 + r, n = n0 + r;
      const int nc = min(n, g.N - 1);
      const int mb = m < g.M ? 16 : 0, nb = n < g.N ? 16 : 0;
      const size_t aoff = size_t(arow[h] + tap * g.dil) * g.lda + k0 + c;
      const size_t woff = (size_t(tap) * g.N + nc) * g.K + k0 + c;
      cp_async16(plane(stage, 0) + r * TG_LD + c, g.Ahi + aoff, mb);
      if (has_alo) cp_async16(plane(stage, 1) + r * TG_LD + c, g.Alo + aoff, mb);
      cp_async16(plane(stage, 2) + r * TG_LD + c, g.Whi + woff, nb);
      if (has_wlo) cp_async16(plane(stage, 3) + r * TG_LD + c, g.Wlo + woff, nb);
    }
  };

  wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[2][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) wmma::fill_fragment(acc[i][j], 0.f);

  const int iters = g.taps * kchunks;
  for (int s = 0; s < TG_STAGES - 1; ++s) {
    if (s < iters) load_stage(s, s);
    cp_async_commit();
  }
  for (int it = 0; it < iters; ++it) {
    cp_async_wait<TG_STAGES - 2>();
    __syncthreads();
    {
      const int nx = it + TG_STAGES - 1;
      if (nx < iters) load_stage(nx % TG_STAGES, nx);
      cp_async_commit();
    }
    const int st = it % TG_STAGES;
#pragma unroll
    for (int kk = 0; kk < TG_BK; kk += 16) {
      wmma::fragment<wmma::matrix_a, 16, 16, 16, __half, wmma::row_major> ah[2], al[2];
      wmma::fragment<wmma::matrix_b, 16, 16, 16, __half, wmma::col_major> wh[2], wl[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        wmma::load_matrix_sync(ah[i], plane(st, 0) + (wm * 32 + i * 16) * TG_LD + kk, TG_LD);
        if (has_alo) wmma::load_matrix_sync(al[i], plane(st, 1) + (wm * 32 + i * 16) * TG_LD + kk, TG_LD);
        wmma::load_matrix_sync(wh[i], plane(st, 2) + (wn * 32 + i * 16) * TG_LD + kk, TG_LD);
        if (has_wlo) wmma::load_matrix_sync(wl[i], plane(st, 3) + (wn * 32 + i * 16) * TG_LD + kk, TG_LD);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (has_wlo) wmma::mma_sync(acc[i][j], ah[i], wl[j], acc[i][j]);
          if (has_alo) wmma::mma_sync(acc[i][j], al[i], wh[j], acc[i][j]);
          wmma::mma_sync(acc[i][j], ah[i], wh[j], acc[i][j]);
        }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  // epilogue: each warp stages its 32x32 f32 block in shared memory, then coalesced rows
  float *stg = reinterpret_cast<float *>(tg_smem) + warp * (32 * 32);
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
      wmma::store_matrix_sync(stg + (i * 16) * 32 + j * 16, acc[i][j], 32, wmma::mem_row_major);
  __syncwarp();
  const int n = n0 + wn * 32 + lane;
  const float bias = (g.bias && n < g.N) ? g.bias[n] : 0.f;
  for (int r = 0; r < 32; ++r) {
    const int m = m0 + wm * 32 + r;
    if (m >= g.M || n >= g.N) continue;
    float old = 0.f;
    if (g.epi == E_BIAS_RESID || g.epi == E_BIAS_LRELU_RESID) old = g.C[size_t(m) * g.ldc + n];
    const float v = apply_epi(g.epi, stg[r * 32 + lane], bias, old);
    if (g.C) g.C[size_t(m) * g.ldc + n] = v;
    if (g.Chi) {
      const __half hi = __float2half_rn(v);
      g.Chi[size_t(m) * g.ldh + n] = hi;
      if (g.Clo) g.Clo[size_t(m) * g.ldh + n] = __float2half_rn(v - __half2float(hi));
    }
  }
}

}  // namespace tts
