// gemm.cuh -- dense f32-accumulate GEMMs used off the decode hot loop.
//
//  (1) sgemm_tn_kernel: C[M][N] = A[M][K] (f32) x W[N][K]^T (f32 or f16 storage), f32 FMA.
//      Used where the reference multiplies in F32 (ggml_mul_mat with F32 operands): the AR
//      prefill / latent pass (main.cpp:2053-2519), diffusion proj_out / emb_layers / time
//      MLP (main.cpp:3331-3343, 3592-3596).
//  (2) hgemm_conv_kernel (mma.sync m16n8k16, f16 x f16 -> f32): every ggml_conv_1d of the
//      diffusion and vocoder graphs, whose reference numerics are F16 im2col x F16 weights
//      with F32 accumulation (ggml.c:6493-6508, 15167-15248).  Implicit GEMM: K taps are
//      K shifted GEMMs over a zero-padded, time-major activation buffer (no im2col).
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
  }
  return acc;
}

// ---- (1) SIMT f32 GEMM, 128x128x16 tiles, 256 threads, 8x8 micro-tiles -------------------
template <typename WT>
__global__ void __launch_bounds__(256) sgemm_tn_kernel(GemmArgs g) {
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
      const float old = g.epi == E_BIAS_RESID ? *c : 0.f;
      *c = apply_epi(g.epi, acc[i][j], bias, old);
    }
  }
}

// ---- (2) f16 tensor-core implicit-GEMM 1-D convolution -------------------------------------
// out[seq][t][oc] = epi( sum_{tap, ic} X[seq][t*stride_in? no: t + tap*dil - pad][ic] * W[tap][oc][ic] + bias[oc] )
// Activations are TIME-MAJOR f16: X[seq][Tpad][IC] where each sequence carries `halo` zero
// rows before and after its T valid rows (so taps never need predication); weights are
// pre-converted once at load to f16 [taps][OC][IC] (reference casts them every graph run,
// main.cpp:3163-3166).  Output is f32 [seq][T][ldo] (+ optional f16 copy for the next conv).
struct ConvArgs {
  const __half *X;     // [nseq][T + 2*halo][IC]
  const __half *W;     // [taps][OC][IC]
  const float *bias;   // [OC] or null
  float *Y;            // [nseq][T][ldo]   f32 output (may be null)
  __half *Yh;          // [nseq][T + 2*halo_o][ldoh] f16 output (may be null), written at row t+halo_o
  const float *R;      // residual [nseq][T][ldo] added before store (may be null)
  int nseq, T, IC, OC, taps, dil, pad, halo, ldo, halo_o, ldoh;
  int epi;             // E_BIAS / E_BIAS_LRELU ...
};

// 128 (time) x 128 (oc) x 32 (ic) CTA tile, 8 warps (4 x 2), warp tile 32 x 64.
__global__ void __launch_bounds__(256) hconv_mma_kernel(ConvArgs c) {
  using namespace nvcuda;
  constexpr int BM = 128, BN = 128, BK = 32, LDS = BK + 8;
  __shared__ __align__(32) __half As[2][BM][LDS];
  __shared__ __align__(32) __half Bs[2][BN][LDS];
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x, warp = tid / 32;
  const int wm = warp / 2, wn = warp % 2;
  const int seq = blockIdx.z;
  const int t0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int Tp = c.T + 2 * c.halo;
  const __half *Xs = c.X + size_t(seq) * Tp * c.IC;

  wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) wmma::fill_fragment(acc[i][j], 0.f);

  const int kchunks = c.IC / BK;
  const int iters = c.taps * kchunks;
  // loader: 256 threads, each 16 halves (2 x uint4): row = tid/2, col = (tid%2)*16
  const int lr = tid / 2, lc = (tid % 2) * 16;

  auto load_tile = [&](int it, int buf) {
    const int tap = it / kchunks, k0 = (it % kchunks) * BK;
    {
      const int t = t0 + lr;
      uint4 u0 = make_uint4(0, 0, 0, 0), u1 = u0;
      if (t < c.T) {
        const int row = t + c.halo + tap * c.dil - c.pad;  // inside [0, Tp) by construction
        const __half *p = Xs + size_t(row) * c.IC + k0 + lc;
        u0 = *reinterpret_cast<const uint4 *>(p);
        u1 = *reinterpret_cast<const uint4 *>(p + 8);
      }
      *reinterpret_cast<uint4 *>(&As[buf][lr][lc]) = u0;
      *reinterpret_cast<uint4 *>(&As[buf][lr][lc + 8]) = u1;
    }
    {
      const int n = n0 + lr;
      uint4 u0 = make_uint4(0, 0, 0, 0), u1 = u0;
      if (n < c.OC) {
        const __half *p = c.W + (size_t(tap) * c.OC + n) * c.IC + k0 + lc;
        u0 = *reinterpret_cast<const uint4 *>(p);
        u1 = *reinterpret_cast<const uint4 *>(p + 8);
      }
      *reinterpret_cast<uint4 *>(&Bs[buf][lr][lc]) = u0;
      *reinterpret_cast<uint4 *>(&Bs[buf][lr][lc + 8]) = u1;
    }
  };

  load_tile(0, 0);
  __syncthreads();
  for (int it = 0; it < iters; ++it) {
    const int buf = it & 1;
    if (it + 1 < iters) load_tile(it + 1, buf ^ 1);
#pragma unroll
    for (int kk = 0; kk < BK; kk += 16) {
      wmma::fragment<wmma::matrix_a, 16, 16, 16, __half, wmma::row_major> fa[2];
      wmma::fragment<wmma::matrix_b, 16, 16, 16, __half, wmma::col_major> fb[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) wmma::load_matrix_sync(fa[i], &As[buf][wm * 32 + i * 16][kk], LDS);
#pragma unroll
      for (int j = 0; j < 4; ++j) wmma::load_matrix_sync(fb[j], &Bs[buf][wn * 64 + j * 16][kk], LDS);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) wmma::mma_sync(acc[i][j], fa[i], fb[j], acc[i][j]);
    }
    __syncthreads();
  }

  // epilogue through shared memory (reuse As as f32 staging: 8 warps x 16x16 floats)
  float *stage = reinterpret_cast<float *>(&As[0][0][0]) + warp * 256;
  const int lane = tid % 32;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      wmma::store_matrix_sync(stage, acc[i][j], 16, wmma::mem_row_major);
      __syncwarp();
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int idx = e * 32 + lane;
        const int r = idx / 16, cc = idx % 16;
        const int t = t0 + wm * 32 + i * 16 + r;
        const int n = n0 + wn * 64 + j * 16 + cc;
        if (t < c.T && n < c.OC) {
          float v = stage[idx] + (c.bias ? c.bias[n] : 0.f);
          if (c.epi == E_BIAS_LRELU) v = v > 0.f ? v : 0.2f * v;
          if (c.R) v += c.R[(size_t(seq) * c.T + t) * c.ldo + n];
          if (c.Y) c.Y[(size_t(seq) * c.T + t) * c.ldo + n] = v;
          if (c.Yh)
            c.Yh[(size_t(seq) * (c.T + 2 * c.halo_o) + t + c.halo_o) * c.ldoh + n] = __float2half_rn(v);
        }
      }
      __syncwarp();
    }
}

}  // namespace tts
