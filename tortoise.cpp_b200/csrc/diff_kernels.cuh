// diff_kernels.cuh -- elementwise / normalisation / attention kernels of the diffusion
// denoiser (reference graph: diffusion_graph, main.cpp:3066-4044).  Activations are
// TIME-MAJOR ([seq][t][channel], channel fastest) -- the transpose of the reference's
// [channel][t] -- so that every 1x1 / k3 convolution is a K-contiguous implicit GEMM.
#pragma once
#include "common.cuh"

namespace tts {

// GroupNorm statistics, 32 groups x 32 channels over all T frames (ggml_group_norm,
// ggml.c:12229-12304: eps 1e-6, double sums, mean subtracted in float before squaring).
// X [nseq][T][1024]; stats[(seq*32+g)*2] = {mean, 1/sqrt(var+eps)}.  grid (32, nseq) x 256.
static __global__ void __launch_bounds__(256) gn_stats_kernel(const float *X, float *stats, int T, const int *Tseq) {
  __shared__ double red[8];
  __shared__ float s_mean;
  pdl_launch_dependents();
  pdl_wait();
  const int g = blockIdx.x, seq = blockIdx.y, tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
  const float *base = X + size_t(seq) * T * kDim + g * 32;
  const int n = (Tseq ? Tseq[seq] : T) * 32;  // rows past the sequence's own length are padding
  double s = 0;
  for (int i = tid; i < n; i += 256) s += double(base[size_t(i >> 5) * kDim + (i & 31)]);
  s = warp_sum_d(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (tid == 0) {
    double tot = 0;
    for (int i = 0; i < 8; ++i) tot += red[i];
    s_mean = float(tot / n);
  }
  __syncthreads();
  const float mean = s_mean;
  double s2 = 0;
  for (int i = tid; i < n; i += 256) {
    const float v = base[size_t(i >> 5) * kDim + (i & 31)] - mean;
    s2 += double(v * v);
  }
  s2 = warp_sum_d(s2);
  __syncthreads();
  if (lane == 0) red[warp] = s2;
  __syncthreads();
  if (tid == 0) {
    double tot = 0;
    for (int i = 0; i < 8; ++i) tot += red[i];
    const float var = float(tot / n);
    stats[(seq * 32 + g) * 2 + 0] = mean;
    stats[(seq * 32 + g) * 2 + 1] = 1.0f / sqrtf(var + 1e-6f);
  }
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }  // ggml.c:2256

// GroupNorm apply + per-channel affine (+ (1+scale), shift) (+ SiLU), written either as the
// f16 operand of the next convolution (with `halo` zero rows before/after each sequence --
// the conv's zero padding) or as f32.
//   v = (x - mean) * rstd * w[c] + b[c];  if ss: v = v * (ss[c] + 1) + ss[1024 + c];  if silu: v = silu(v)
// (main.cpp:3356-3372 norm+affine, 3449-3452 scale/shift, 3373/3454 silu; the conv's im2col
// rounds to f16, ggml.c:6493-6508.)  ss_stride: elements between the sequences' ss vectors.
// grid (T + 2*halo, nseq) x 256.
static __global__ void __launch_bounds__(256) gn_apply_kernel(const float *X, const float *stats, const float *w,
                                                       const float *b, const float *ss, __half *out16,
                                                       float *out32, int T, int halo, int ldo, int silu,
                                                       const int *step_ptr, int ss_step_stride,
                                                       const double *partial, int mtiles, const int *Tseq) {
  pdl_launch_dependents();
  pdl_wait();
  // the per-timestep scale|shift table advances with a device-side step counter so that one
  // captured CUDA graph serves every sampling step
  if (ss && step_ptr) ss += size_t(*step_ptr) * ss_step_stride;
  const int row = blockIdx.x, seq = blockIdx.y, tid = threadIdx.x;
  const int t = row - halo;
  const int c = tid * 4;
  const int Ts = Tseq ? Tseq[seq] : T;  // valid frames of this sequence (T is the common row stride)
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (t >= 0 && t < Ts) {
    const float4 x = *reinterpret_cast<const float4 *>(X + (size_t(seq) * T + t) * kDim + c);
    const int g = c >> 5;
    float mean, rstd;
    if (partial) {
      // statistics fused into the producing GEMM's epilogue: {sum, sumsq} per M tile, in double
      double s1 = 0.0, s2 = 0.0;
      for (int i = 0; i < mtiles; ++i) {
        s1 += partial[((size_t(seq) * 32 + g) * mtiles + i) * 2];
        s2 += partial[((size_t(seq) * 32 + g) * mtiles + i) * 2 + 1];
      }
      const double n = double(Ts) * 32.0, md = s1 / n;
      mean = float(md);
      double var = s2 / n - 2.0 * md * double(mean) + double(mean) * double(mean);  // E[(x - mean_f)^2]
      if (var < 0.0) var = 0.0;
      rstd = 1.0f / sqrtf(float(var) + 1e-6f);
    } else {
      mean = stats[(seq * 32 + g) * 2];
      rstd = stats[(seq * 32 + g) * 2 + 1];
    }
    const float4 w4 = *reinterpret_cast<const float4 *>(w + c);
    const float4 b4 = *reinterpret_cast<const float4 *>(b + c);
    v[0] = (x.x - mean) * rstd * w4.x + b4.x;
    v[1] = (x.y - mean) * rstd * w4.y + b4.y;
    v[2] = (x.z - mean) * rstd * w4.z + b4.z;
    v[3] = (x.w - mean) * rstd * w4.w + b4.w;
    if (ss) {
      const float4 sc = *reinterpret_cast<const float4 *>(ss + c);
      const float4 sh = *reinterpret_cast<const float4 *>(ss + kDim + c);
      v[0] = v[0] * (sc.x + 1.0f) + sh.x;
      v[1] = v[1] * (sc.y + 1.0f) + sh.y;
      v[2] = v[2] * (sc.z + 1.0f) + sh.z;
      v[3] = v[3] * (sc.w + 1.0f) + sh.w;
    }
    if (silu) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = silu_f(v[i]);
    }
  }
  if (out16) {
    __half h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __float2half_rn(v[i]);
    *reinterpret_cast<uint2 *>(out16 + (size_t(seq) * (T + 2 * halo) + row) * ldo + c) =
        *reinterpret_cast<uint2 *>(h);
  }
  if (out32 && t >= 0 && t < Ts)
    *reinterpret_cast<float4 *>(out32 + (size_t(seq) * T + t) * kDim + c) = make_float4(v[0], v[1], v[2], v[3]);
}

// f32 [T][C] -> f16 [T + 2*halo][ldo] with zero halo rows and zero channel padding.
// grid (T + 2*halo) x 256
static __global__ void __launch_bounds__(256) to_f16_halo_kernel(const float *X, __half *out, int T, int C, int halo,
                                                          int ldo) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x, t = row - halo;
  for (int c = threadIdx.x; c < ldo; c += 256) {
    float v = 0.f;
    if (t >= 0 && t < T && c < C) v = X[size_t(t) * C + c];
    out[size_t(row) * ldo + c] = __float2half_rn(v);
  }
}

// x [100][S] (channel-major, the reference's noise_tensor layout) -> f16 [S + 2][128]
// Utterance batching: grid.y = utterance u; x of utterance u starts at x + xoff[u] with ITS OWN S_u (the
// noise blocks keep the reference's [100][S_u] layout), out rows use the common stride S.
static __global__ void __launch_bounds__(128) xin_kernel(const float *x, __half *out, int S, const int *Sutt, const long long *xoff) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x, u = blockIdx.y, t = row - 1, c = threadIdx.x;
  const int Su = Sutt ? Sutt[u] : S;
  const float *xu = x + (xoff ? xoff[u] : 0);
  float v = 0.f;
  if (t >= 0 && t < Su && c < 100) v = xu[size_t(c) * Su + t];
  out[(size_t(u) * (S + 2) + row) * 128 + c] = __float2half_rn(v);
}

// CAT16[seq][t+1][0:1024] = f16(INP[t]), [1024:2048] = f16(CW[seq][t])  (channel concat,
// main.cpp:3635-3637), halo rows zero.  grid (S + 2, nseq) x 256
// (sequence seq = 2 u + {0 cond, 1 uncond}: both read the input-block output of utterance u = seq / 2)
static __global__ void __launch_bounds__(256) concat_kernel(const float *INP, const float *CW, __half *out, int S, const int *Tseq) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x, seq = blockIdx.y, t = row - 1, tid = threadIdx.x;
  const int Ts = Tseq ? Tseq[seq] : S;
  __half *o = out + (size_t(seq) * (S + 2) + row) * 2048;
  for (int c = tid; c < 2048; c += 256) {
    float v = 0.f;
    if (t >= 0 && t < Ts) v = c < 1024 ? INP[(size_t(seq >> 1) * S + t) * kDim + c] : CW[(size_t(seq) * S + t) * kDim + c - 1024];
    o[c] = __float2half_rn(v);
  }
}

// nearest-neighbour stretch L -> S (ggml_upscale_ext, ggml.c:15527-15568; index table built
// on the host with the reference's float arithmetic) and the unconditioned broadcast
// (main.cpp:3321-3328).  CE [2][S][1024].  grid (S, 2) x 256
// (CE points at the two sequences of ONE utterance; S is the common row stride, St its own length)
static __global__ void __launch_bounds__(256) code_emb_kernel(const float *CL, const int *src_idx, const float *uncond,
                                                       float *CE, int S, int St) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.x, seq = blockIdx.y, c = threadIdx.x * 4;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t < St) {
    if (seq == 0) v = *reinterpret_cast<const float4 *>(CL + size_t(src_idx[t]) * kDim + c);
    else v = *reinterpret_cast<const float4 *>(uncond + c);
  }
  *reinterpret_cast<float4 *>(CE + (size_t(seq) * S + t) * kDim + c) = v;
}

// in f32 [n] -> (optional SiLU) -> hi/lo f16 planes
static __global__ void silu_split_kernel(const float *in, __half *hi, __half *lo, size_t n, int do_silu) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    float v = in[i];
    if (do_silu) v = silu_f(v);
    const __half h = __float2half_rn(v);
    hi[i] = h;
    lo[i] = __float2half_rn(v - __half2float(h));
  }
}

// Self-attention with T5-style relative position bias (main.cpp:3547-3596):
//   w_ij = softmax_j( q_i.k_j / 8 + 8 * relbias[bucket(i, j)][head] ),  out_i = sum_j w_ij v_j
// QKV [nseq][T][3072] f32 with head h owning channels [192h, 192h+192): q | k | v.
// bucket(i, j) = (j > i ? 16 : 0) + rpb[|j - i|] (main.cpp:4722-4749, table from the host).
// Output as split-f16 planes [nseq*T][1024] (operand of the F32 proj_out matmul).
// grid (ceil(T / DA_Q), 16, nseq) x DA_THREADS: 8 warps x 4 queries, keys staged in tiles of 64.
// (32 queries per block instead of 16: every block re-reads the K and V of its head from L2 --
// 98 KB at T = 191 -- so the block count per head IS the L2 traffic: 37.6 -> 18.8 MB per call.)
constexpr int DA_WARPS = 8, DA_THREADS = DA_WARPS * 32, DA_Q = DA_WARPS * 4, DA_TK = 64, DA_LDK = kHeadDim + 4;
constexpr size_t DA_SMEM = (size_t(2) * DA_TK * DA_LDK + size_t(DA_Q) * kHeadDim + size_t(DA_WARPS) * 4 * DA_TK + 32) * sizeof(float);
static __global__ void __launch_bounds__(DA_THREADS) diff_attn_kernel(const float *QKV, const float *relbias, const int *rpb,
                                                               __half *out_hi, __half *out_lo, int Tstride, const int *Tseq) {
  constexpr int TK = DA_TK, LDK = DA_LDK;
  extern __shared__ __align__(16) float da_smem[];
  float (*Ks)[LDK] = reinterpret_cast<float (*)[LDK]>(da_smem);
  float (*Vs)[LDK] = reinterpret_cast<float (*)[LDK]>(da_smem + TK * LDK);
  float (*Qs)[kHeadDim] = reinterpret_cast<float (*)[kHeadDim]>(da_smem + 2 * TK * LDK);
  float (*Ps)[4][TK] = reinterpret_cast<float (*)[4][TK]>(da_smem + 2 * TK * LDK + DA_Q * kHeadDim);
  float *bias_s = da_smem + 2 * TK * LDK + DA_Q * kHeadDim + DA_WARPS * 4 * TK;
  pdl_launch_dependents();
  pdl_wait();
  const int t = threadIdx.x, warp = t / 32, lane = t % 32;
  const int head = blockIdx.y, seq = blockIdx.z;
  const int q0 = blockIdx.x * DA_Q;
  const int T = Tseq ? Tseq[seq] : Tstride;  // this sequence's own length (keys / queries past it are padding)
  if (q0 >= T) return;
  const float *base = QKV + size_t(seq) * Tstride * 3072 + head * 192;
  if (t < 32) bias_s[t] = 8.0f * relbias[t * 16 + head];
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
      if (qi < T && j0 < T) s0 = s[i][0] * 0.125f + bias_s[(j0 > qi ? 16 : 0) + rpb[abs(j0 - qi)]];
      if (qi < T && j1 < T) s1 = s[i][1] * 0.125f + bias_s[(j1 > qi ? 16 : 0) + rpb[abs(j1 - qi)]];
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
      const size_t off = (size_t(seq) * Tstride + qi) * kDim + head * kHeadDim;
      const float v0 = o[i][0] * inv, v1 = o[i][1] * inv;
      const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
      out_hi[off + lane] = h0;
      out_hi[off + lane + 32] = h1;
      out_lo[off + lane] = __float2half_rn(v0 - __half2float(h0));
      out_lo[off + lane + 32] = __float2half_rn(v1 - __half2float(h1));
    }
  }
}

// One DDPM ancestral step on the device (host math of diffusion(), main.cpp:5970-6030):
//   eps = (1+k) eps_c - k eps_u ; frac = (v+1)/2 ; logvar = frac*min_log + (1-frac)*max_log
//   (argument swap A-9) ; x0 = clamp(a x - b eps, +-1) ; mean = c1 x0 + c2 x ;
//   x' = last ? mean : mean + exp(0.5 logvar) n          (double product, main.cpp:5606)
// x, noise: [100][S] channel-major; OUT: [2][S][200] time-major (seq 0 cond, seq 1 uncond).
struct DdpmCoef {
  float cfk, sqrt_recip, sqrt_recipm1, coef1, coef2, min_log, max_log;
  int last;
};
// Utterance batching (grid.y = utterance u): x of utterance u at x + xoff[u] in ITS OWN [100][S_u] layout, its
// noise at noise_base + noff[u] + (step + 1) * 100 * S_u, its model outputs in sequences 2u / 2u + 1 (row stride S).
static __global__ void __launch_bounds__(256) ddpm_step_kernel(float *x, const float *OUT, const float *noise_base,
                                                        const DdpmCoef *coefs, const int *step_ptr, int S, const int *Sutt,
                                                        const long long *xoff, const long long *noff) {
  pdl_launch_dependents();
  pdl_wait();
  const int step = *step_ptr, u = blockIdx.y;
  const DdpmCoef k = coefs[step];
  const int Su = Sutt ? Sutt[u] : S;
  const int n = 100 * Su;
  x += xoff ? xoff[u] : 0;
  OUT += size_t(2 * u) * S * 200;
  const float *noise = noise_base + (noff ? noff[u] : 0) + size_t(step + 1) * n;  // block 0 was the initial x
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int ch = i / Su, t = i % Su;
    const float eps_c = OUT[size_t(t) * 200 + ch];
    const float vraw = OUT[size_t(t) * 200 + 100 + ch];
    const float eps_u = OUT[(size_t(S) + t) * 200 + ch];
    // un-fused float arithmetic in the reference's operation order (host code built without FMA)
    const float frac = __fdiv_rn(__fadd_rn(vraw, 1.0f), 2.0f);
    const float logvar = __fadd_rn(__fmul_rn(frac, k.min_log), __fmul_rn(__fsub_rn(1.0f, frac), k.max_log));
    const float eps = __fsub_rn(__fmul_rn(__fadd_rn(1.0f, k.cfk), eps_c), __fmul_rn(k.cfk, eps_u));
    const float xv = x[i];
    float x0 = __fsub_rn(__fmul_rn(k.sqrt_recip, xv), __fmul_rn(k.sqrt_recipm1, eps));
    x0 = fminf(1.0f, fmaxf(-1.0f, x0));
    const float mean = __fadd_rn(__fmul_rn(k.coef1, x0), __fmul_rn(k.coef2, xv));
    float r = mean;
    if (!k.last) r = float(__dadd_rn(double(mean), __dmul_rn(exp(__dmul_rn(0.5, double(logvar))), double(noise[i]))));
    x[i] = r;
  }
}

static __global__ void step_inc_kernel(int *step_ptr) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) *step_ptr += 1;
}

// weight re-layout at load: conv weight f32 [OC][IC][K] (K fastest) -> f16 [K][OC][ICpad]
static __global__ void conv_weight_kernel(const float *src, __half *dst, int OC, int IC, int K, int ICpad) {
  const size_t n = size_t(K) * OC * ICpad;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const int ic = int(i % ICpad);
    const int oc = int((i / ICpad) % OC);
    const int k = int(i / (size_t(ICpad) * OC));
    float v = 0.f;
    if (ic < IC) v = src[(size_t(oc) * IC + ic) * K + k];
    dst[i] = __float2half_rn(v);
  }
}

}  // namespace tts
