// diff_kernels.cuh -- elementwise / normalisation / attention kernels of the diffusion
// denoiser (reference graph: diffusion_graph, main.cpp:3066-4044).  Activations are
// TIME-MAJOR ([seq][t][channel], channel fastest) -- the transpose of the reference's
// [channel][t] -- so that every 1x1 / k3 convolution is a K-contiguous implicit GEMM.
#pragma once
#include <type_traits>

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
// grid (ceil((T + 2*halo) / R), nseq) x 256: a block normalises R rows -- the statistics, affine and scale/shift
// vectors are fetched once and the R row loads are in flight together (one row per block was three dependent
// L2 round trips for 4 KB: 13.7 us per call at S = 1306, 1.2 TB/s).
template <int R>
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
  const int row0 = blockIdx.x * R, seq = blockIdx.y, tid = threadIdx.x;
  const int c = tid * 4;
  const int Ts = Tseq ? Tseq[seq] : T;  // valid frames of this sequence (T is the common row stride)
  const int nrows = T + 2 * halo;
  float4 x[R];
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int t = row0 + i - halo;
    x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t >= 0 && t < Ts) x[i] = *reinterpret_cast<const float4 *>(X + (size_t(seq) * T + t) * kDim + c);
  }
  const int g = c >> 5;
  float mean, rstd;
  if (partial) {
    // statistics fused into the producing GEMM's epilogue but not reduced there: {sum, sumsq} per M tile, in double
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
  float4 sc = make_float4(0.f, 0.f, 0.f, 0.f), sh = sc;
  if (ss) {
    sc = *reinterpret_cast<const float4 *>(ss + c);
    sh = *reinterpret_cast<const float4 *>(ss + kDim + c);
  }
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int row = row0 + i, t = row - halo;
    if (row >= nrows) break;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    const bool live = t >= 0 && t < Ts;
    if (live) {
      v[0] = (x[i].x - mean) * rstd * w4.x + b4.x;
      v[1] = (x[i].y - mean) * rstd * w4.y + b4.y;
      v[2] = (x[i].z - mean) * rstd * w4.z + b4.z;
      v[3] = (x[i].w - mean) * rstd * w4.w + b4.w;
      if (ss) {
        v[0] = v[0] * (sc.x + 1.0f) + sh.x;
        v[1] = v[1] * (sc.y + 1.0f) + sh.y;
        v[2] = v[2] * (sc.z + 1.0f) + sh.z;
        v[3] = v[3] * (sc.w + 1.0f) + sh.w;
      }
      if (silu) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = silu_f(v[e]);
      }
    }
    if (out16) {
      __half h[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) h[e] = __float2half_rn(v[e]);
      *reinterpret_cast<uint2 *>(out16 + (size_t(seq) * nrows + row) * ldo + c) = *reinterpret_cast<uint2 *>(h);
    }
    if (out32 && live) *reinterpret_cast<float4 *>(out32 + (size_t(seq) * T + t) * kDim + c) = make_float4(v[0], v[1], v[2], v[3]);
  }
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
// QKV16 [nseq][T][3072] f16 with head h owning channels [192h, 192h+192): q | k | v.
// bucket(i, j) = (j > i ? 16 : 0) + rpb[|j - i|] (main.cpp:4722-4749, table from the host).
// Output as split-f16 planes [nseq*T][1024] (operand of the F32 proj_out matmul).
// Flash-style: a block owns 16 NW queries of one (sequence, head), walks the keys in tiles of 64 (K and V tiles
// double-buffered with cp.async), S = Q K^T and O += P V on mma.sync m16n8k16 (f16 operands, f32 accumulators),
// online softmax in f32 on the accumulator fragments, P goes from the S accumulators straight into the A fragments
// of the second product.  QKV16 is the f16 copy of conv1(GN(x)) the QKV GEMM's epilogue writes ([row][3072], head h
// at columns 192 h: q | k | v); the relative-position bias of the block's (query, key) distances is tabulated once
// per block (T + 16 NW - 1 entries).  exp is exp2 with log2(e) folded into the scale and the table.
// Why mma.sync and not tcgen05 here: per (head, sequence) the products are 64-wide in K (head dim) with a softmax
// between them -- an accumulator round trip TMEM -> registers -> shared per key tile -- and at S = 191 (BASELINE
// configs[1]) the whole call is 0.6 GFLOP.  The f32 SIMT kernel this replaces (round 1) took 12 us per call at
// S = 191 and ~0.55 ms at S = 1306: the sampling step went 1517 -> 1125 us (S = 191) and ~12 -> 4.9 ms (S = 1306).
constexpr int TA_BK = 64, TA_LD = kHeadDim + 8;  // 72 halves per row: ldmatrix rows fall on distinct banks
template <int NW>
constexpr size_t ta_smem_bytes(int T) {
  return size_t(4 * TA_BK) * TA_LD * sizeof(__half) + size_t(T + 16 * NW + TA_BK) * sizeof(float) + 32 * sizeof(float);
}
__device__ __forceinline__ void ta_cp16(void *dst, const void *src, bool valid) {
  const int n = valid ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void ta_ldm4(uint32_t (&r)[4], const void *p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ta_ldm4_t(uint32_t (&r)[4], const void *p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ta_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 2^x for x <= 0 (softmax weights): one MUFU.EX2; exp2f() wraps it in a range fix-up (compare + two multiplies) that
// only matters for |x| > 126, where flushing to zero is what a softmax weight should do anyway
__device__ __forceinline__ float ta_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t ta_pack(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t *>(&h);
}

// grid (ceil(T / (16 NW)), 16 heads, nseq) x 32 NW
template <int NW>
static __global__ void __launch_bounds__(NW * 32, NW == 4 ? 5 : 8) diff_attn_tc_kernel(const __half *QKV16, const float *relbias, const int *rpb,
                                                               __half *out_hi, __half *out_lo, int Tstride, const int *Tseq) {
  constexpr int BQ = 16 * NW, NT = NW * 32, LD = TA_LD;
  constexpr float kLog2e = 1.4426950408889634f;
  extern __shared__ __align__(16) unsigned char ta_raw[];
  __half (*Ks)[TA_BK][LD] = reinterpret_cast<__half (*)[TA_BK][LD]>(ta_raw);
  __half (*Vs)[TA_BK][LD] = reinterpret_cast<__half (*)[TA_BK][LD]>(ta_raw + size_t(2 * TA_BK) * LD * 2);
  __half (*Qs)[LD] = Ks[1];  // the query tile passes through the second K buffer on its way to registers
  float *bias_s = reinterpret_cast<float *>(ta_raw + size_t(4 * TA_BK) * LD * 2);
  static_assert(BQ <= TA_BK, "the query tile is staged in one K buffer");
  float *bt = bias_s + 32;
  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32, g = lane >> 2, t4 = lane & 3;
  const int head = blockIdx.y, seq = blockIdx.z, q0 = blockIdx.x * BQ;
  const int T = Tseq ? Tseq[seq] : Tstride;  // this sequence's own length (rows past it are padding)
  if (q0 >= T) return;
  const __half *base = QKV16 + size_t(seq) * Tstride * 3072 + head * 192;

  auto load_kv = [&](int kt, int buf) {
    const int k0 = kt * TA_BK;
    for (int i = tid; i < TA_BK * 16; i += NT) {
      const int r = (i >> 3) & (TA_BK - 1), c = (i & 7) * 8, isv = i >> 9;
      const int kj = k0 + r;
      const bool ok = kj < T;
      const __half *src = base + size_t(ok ? kj : 0) * 3072 + (isv ? 128 : 64) + c;
      ta_cp16(isv ? &Vs[buf][r][c] : &Ks[buf][r][c], src, ok);
    }
  };
  for (int i = tid; i < BQ * 8; i += NT) {
    const int r = i >> 3, c = (i & 7) * 8;
    const bool ok = q0 + r < T;
    ta_cp16(&Qs[r][c], base + size_t(ok ? q0 + r : 0) * 3072 + c, ok);
  }
  load_kv(0, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  if (tid < 32) bias_s[tid] = 8.0f * kLog2e * relbias[tid * 16 + head];
  __syncthreads();
  // bt[j - qi + q0 + BQ - 1] for the block's queries qi in [q0, q0 + BQ) and every key j in [0, T)
  for (int i = tid; i < T + BQ - 1; i += NT) {
    const int d = i - (q0 + BQ - 1);
    bt[i] = bias_s[(d > 0 ? 16 : 0) + rpb[min(abs(d), T - 1)]];
  }

  uint32_t qf[4][4];
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) ta_ldm4(qf[kk], &Qs[warp * 16 + ((lane & 7) + ((lane >> 3) & 1) * 8)][kk * 16 + (lane >> 4) * 8]);
  __syncthreads();  // every warp holds its query fragments: the buffer is free for key tile 1
  float o[8][4], m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  const int nkt = (T + TA_BK - 1) / TA_BK;
  const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lcol = (lane >> 4) * 8;  // ldmatrix.x4 address pattern (A / V^T)
  const int boff = BQ - 1 - warp * 16 - g;  // bias index of (row g, key j) = j + boff; row g + 8: - 8
  for (int kt = 0; kt < nkt; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nkt) {
      load_kv(kt + 1, buf ^ 1);
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      uint32_t kf[4];
      // K rows nt*8 .. +7: matrices = d chunks (lane / 8) * 8 of this half of the head dim
      ta_ldm4(kf, &Ks[buf][nt * 8 + (lane & 7)][(lane >> 3) * 8]);
      ta_mma(s[nt], qf[0], kf[0], kf[1]);
      ta_mma(s[nt], qf[1], kf[2], kf[3]);
      ta_ldm4(kf, &Ks[buf][nt * 8 + (lane & 7)][32 + (lane >> 3) * 8]);
      ta_mma(s[nt], qf[2], kf[0], kf[1]);
      ta_mma(s[nt], qf[3], kf[2], kf[3]);
    }
    const int k0 = kt * TA_BK;
    float mx_lo = -INFINITY, mx_hi = -INFINITY;
    {
      // bias of (row g, key j) = btp[j - k0 - 2 t4]; row g + 8 sits 8 entries lower: ITS bias at key tile nt is row g's
      // at key tile nt - 1, so each entry is fetched once (18 loads per tile instead of 32).  The table has TA_BK
      // entries of slack behind T + BQ - 1: a partial last tile reads (and then masks) past the live part.
      const float *btp = bt + k0 + 2 * t4 + boff;
      auto scores = [&](auto masked) {
        float p0 = btp[-8], p1 = btp[-7];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const float b0 = btp[nt * 8], b1 = btp[nt * 8 + 1];
          float a0 = fmaf(s[nt][0], 0.125f * kLog2e, b0), a1 = fmaf(s[nt][1], 0.125f * kLog2e, b1);
          float a2 = fmaf(s[nt][2], 0.125f * kLog2e, p0), a3 = fmaf(s[nt][3], 0.125f * kLog2e, p1);
          p0 = b0;
          p1 = b1;
          if (decltype(masked)::value) {
            const int j = k0 + nt * 8 + 2 * t4;
            if (j >= T) a0 = a2 = -INFINITY;
            if (j + 1 >= T) a1 = a3 = -INFINITY;
          }
          s[nt][0] = a0;
          s[nt][1] = a1;
          s[nt][2] = a2;
          s[nt][3] = a3;
          mx_lo = fmaxf(mx_lo, fmaxf(a0, a1));
          mx_hi = fmaxf(mx_hi, fmaxf(a2, a3));
        }
      };
      // every key of the tile exists (all tiles but the last): no masks -- a real branch, not predication
      if (k0 + TA_BK <= T) scores(std::false_type{});
      else scores(std::true_type{});
    }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    const float mn_lo = fmaxf(m_lo, mx_lo), mn_hi = fmaxf(m_hi, mx_hi);  // finite: key k0 is always valid
    const float c_lo = ta_ex2(m_lo - mn_lo), c_hi = ta_ex2(m_hi - mn_hi);
    m_lo = mn_lo;
    m_hi = mn_hi;
    float ps_lo = 0.f, ps_hi = 0.f;
    uint32_t pf[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = ta_ex2(s[nt][0] - mn_lo), p1 = ta_ex2(s[nt][1] - mn_lo);
      const float p2 = ta_ex2(s[nt][2] - mn_hi), p3 = ta_ex2(s[nt][3] - mn_hi);
      ps_lo += p0 + p1;
      ps_hi += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2] = ta_pack(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = ta_pack(p2, p3);
      o[nt][0] *= c_lo;
      o[nt][1] *= c_lo;
      o[nt][2] *= c_hi;
      o[nt][3] *= c_hi;
    }
    l_lo = l_lo * c_lo + ps_lo;
    l_hi = l_hi * c_hi + ps_hi;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dn = 0; dn < 8; dn += 2) {
        uint32_t vf[4];
        ta_ldm4_t(vf, &Vs[buf][kk * 16 + lrow][dn * 8 + lcol]);
        ta_mma(o[dn], pf[kk], vf[0], vf[1]);
        ta_mma(o[dn + 1], pf[kk], vf[2], vf[3]);
      }
    }
    __syncthreads();  // every warp is done with this buffer before the tile after next lands in it
  }
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int qi = q0 + warp * 16 + g + 8 * h;
    if (qi >= T) continue;
    const float inv = 1.0f / (h ? l_hi : l_lo);
    const size_t off = (size_t(seq) * Tstride + qi) * kDim + head * kHeadDim + 2 * t4;
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) {
      const float v0 = o[dn][2 * h] * inv, v1 = o[dn][2 * h + 1] * inv;
      const __half2 hi = __floats2half2_rn(v0, v1);
      const float2 back = __half22float2(hi);
      *reinterpret_cast<__half2 *>(out_hi + off + dn * 8) = hi;
      *reinterpret_cast<__half2 *>(out_lo + off + dn * 8) = __floats2half2_rn(v0 - back.x, v1 - back.y);
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
