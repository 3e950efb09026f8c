// ar_kernels.cuh -- non-GEMV kernels of the AR stage (embedding, attention, LayerNorm rows).
#pragma once
#include "common.cuh"

namespace tts {

// h[b] = mel_emb[tok[b]] + mel_pos[pos_id]      (decode input, main.cpp:2668-2691)
static __global__ void __launch_bounds__(256) ar_embed_decode_kernel(const int *tokens, const int *state,
                                                              const float *mel_emb, const float *mel_pos,
                                                              float *h) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, t = threadIdx.x;
  const int tok = tokens[b], pos = state[1];
  const float4 e = reinterpret_cast<const float4 *>(mel_emb + size_t(tok) * kDim)[t];
  const float4 p = reinterpret_cast<const float4 *>(mel_pos + size_t(pos) * kDim)[t];
  reinterpret_cast<float4 *>(h + size_t(b) * kDim)[t] = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
}

// Rows of the prefill / latent input (main.cpp:2586-2666 and 2076-2165):
//   row 0            = voice conditioning latent
//   rows 1..T        = text_emb[tok_j] + text_pos[j]
//   rows T+1..       = mel_emb[code] + mel_pos[p]
// codes/pos are per candidate ([B][n_mel]); text is shared.
static __global__ void __launch_bounds__(256) ar_embed_rows_kernel(const int *text, int T, const float *voice,
                                                            const int *codes, const int *mel_positions,
                                                            int n_mel, const float *text_emb,
                                                            const float *text_pos, const float *mel_emb,
                                                            const float *mel_pos, float *H) {
  pdl_launch_dependents();
  pdl_wait();
  const int R = 1 + T + n_mel;
  const int row = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  float4 v;
  if (row == 0) {
    v = reinterpret_cast<const float4 *>(voice)[t];
  } else if (row <= T) {
    const int j = row - 1;
    const float4 e = reinterpret_cast<const float4 *>(text_emb + size_t(text[j]) * kDim)[t];
    const float4 p = reinterpret_cast<const float4 *>(text_pos + size_t(j) * kDim)[t];
    v = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
  } else {
    const int j = row - 1 - T;
    const int code = codes[b * n_mel + j], pos = mel_positions[b * n_mel + j];
    const float4 e = reinterpret_cast<const float4 *>(mel_emb + size_t(code) * kDim)[t];
    const float4 p = reinterpret_cast<const float4 *>(mel_pos + size_t(pos) * kDim)[t];
    v = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
  }
  reinterpret_cast<float4 *>(H + (size_t(b) * R + row) * kDim)[t] = v;
}

// y = LN(x) * w + b per row of 1024 (eps 1e-5, double accumulation, ggml.c:11905-11958);
// optional second parameterised LN on top (the "double final norm", SURVEY A-1).
// Output: f32 Y (may be null) and/or split-f16 planes Yhi/Ylo (may be null) for tgemm.
static __global__ void __launch_bounds__(256) ln_rows_kernel(const float *X, float *Y, __half *Yhi, __half *Ylo,
                                                      const float *w1, const float *b1, const float *w2,
                                                      const float *b2, int ldx, int ldy) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double red[8];
  const int row = blockIdx.x, t = threadIdx.x, warp = t / 32, lane = t % 32;
  const float4 v = reinterpret_cast<const float4 *>(X + size_t(row) * ldx)[t];
  float x[4] = {v.x, v.y, v.z, v.w};
  const int passes = w2 ? 2 : 1;
  for (int p = 0; p < passes; ++p) {
    double s = double(x[0]) + double(x[1]) + double(x[2]) + double(x[3]);
    s = warp_sum_d(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    double tot = 0;
    for (int i = 0; i < 8; ++i) tot += red[i];
    __syncthreads();
    const float mean = float(tot / kDim);
    float d[4];
    double s2 = 0;
    for (int i = 0; i < 4; ++i) {
      d[i] = x[i] - mean;
      s2 += double(d[i] * d[i]);
    }
    s2 = warp_sum_d(s2);
    if (lane == 0) red[warp] = s2;
    __syncthreads();
    tot = 0;
    for (int i = 0; i < 8; ++i) tot += red[i];
    __syncthreads();
    const float rstd = 1.0f / sqrtf(float(tot / kDim) + 1e-5f);
    const float4 w4 = reinterpret_cast<const float4 *>(p == 0 ? w1 : w2)[t];
    const float4 b4 = reinterpret_cast<const float4 *>(p == 0 ? b1 : b2)[t];
    x[0] = d[0] * rstd * w4.x + b4.x;
    x[1] = d[1] * rstd * w4.y + b4.y;
    x[2] = d[2] * rstd * w4.z + b4.z;
    x[3] = d[3] * rstd * w4.w + b4.w;
  }
  if (Y) reinterpret_cast<float4 *>(Y + size_t(row) * ldy)[t] = make_float4(x[0], x[1], x[2], x[3]);
  if (Yhi) {
    __half hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      hi[i] = __float2half_rn(x[i]);
      lo[i] = __float2half_rn(x[i] - __half2float(hi[i]));
    }
    *reinterpret_cast<uint2 *>(Yhi + size_t(row) * ldy + t * 4) = *reinterpret_cast<uint2 *>(hi);
    if (Ylo) *reinterpret_cast<uint2 *>(Ylo + size_t(row) * ldy + t * 4) = *reinterpret_cast<uint2 *>(lo);
  }
}

// Single-position attention against the f16 KV cache (main.cpp:2828-2886 with
// test_dimension == 1): scores = q.k/8 over keys 0..n_past, softmax, weighted sum of v.
// grid (16 heads, B), 128 threads.  K/V: [b][head][pos][64] f16 (the cached values ARE
// f16-exact in the reference too: they pass through the F16 round trip, A-2).
static __global__ void __launch_bounds__(128) ar_attn_decode_kernel(const float *q, const __half *kc,
                                                             const __half *vc, float *out,
                                                             const int *state, int P) {
  extern __shared__ float sc[];  // [n] scores, then 2*64 partial outputs
  __shared__ float qs[kHeadDim];
  __shared__ float redf[4];
  pdl_launch_dependents();
  pdl_wait();
  const int head = blockIdx.x, b = blockIdx.y, t = threadIdx.x, warp = t / 32, lane = t % 32;
  const int n = state[0] + 1;
  if (t < kHeadDim) qs[t] = q[size_t(b) * kDim + head * kHeadDim + t];
  __syncthreads();
  const __half *K = kc + (size_t(b) * kHeads + head) * size_t(P) * kHeadDim;
  const __half *V = vc + (size_t(b) * kHeads + head) * size_t(P) * kHeadDim;
  float lmax = -INFINITY;
  for (int j = t; j < n; j += 128) {
    const uint4 *kr = reinterpret_cast<const uint4 *>(K + size_t(j) * kHeadDim);
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint4 u = kr[c];
      const __half2 *h2 = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h2[e]);
        dot = fmaf(qs[c * 8 + 2 * e], f.x, dot);
        dot = fmaf(qs[c * 8 + 2 * e + 1], f.y, dot);
      }
    }
    dot *= 0.125f;
    sc[j] = dot;
    lmax = fmaxf(lmax, dot);
  }
  lmax = warp_max(lmax);
  if (lane == 0) redf[warp] = lmax;
  __syncthreads();
  const float mx = fmaxf(fmaxf(redf[0], redf[1]), fmaxf(redf[2], redf[3]));
  __syncthreads();
  float lsum = 0.f;
  for (int j = t; j < n; j += 128) {
    const float p = expf(sc[j] - mx);
    sc[j] = p;
    lsum += p;
  }
  lsum = warp_sum(lsum);
  if (lane == 0) redf[warp] = lsum;
  __syncthreads();
  const float inv = float(1.0 / (double(redf[0]) + double(redf[1]) + double(redf[2]) + double(redf[3])));
  // PV: thread -> (dim d, key parity half)
  const int d = t % kHeadDim, half = t / kHeadDim;
  float acc = 0.f;
  for (int j = half; j < n; j += 2) acc = fmaf(sc[j] * inv, __half2float(V[size_t(j) * kHeadDim + d]), acc);
  float *po = sc + n;
  po[half * kHeadDim + d] = acc;
  __syncthreads();
  if (t < kHeadDim) out[size_t(b) * kDim + head * kHeadDim + t] = po[t] + po[kHeadDim + t];
}

// Causal self-attention over full rows (prefill and latent pass; main.cpp:2275-2330 /
// 2828-2886 with n_past == 0).  QKV: [b][R][3072] f32 (already f16-rounded values).
// grid (ceil(R/16), 16 heads, B), 128 threads = 4 warps, each warp 4 query rows.
// Keys are staged through shared memory in tiles of 64; lane <-> key; online softmax.
static __global__ void __launch_bounds__(128) ar_attn_causal_kernel(const float *QKV, __half *out_hi, __half *out_lo,
                                                             int R) {
  constexpr int TK = 64, LDK = kHeadDim + 4;
  __shared__ __align__(16) float Ks[TK][LDK];
  __shared__ __align__(16) float Vs[TK][LDK];
  __shared__ __align__(16) float Qs[16][kHeadDim];
  __shared__ float Ps[4][4][TK];
  pdl_launch_dependents();
  pdl_wait();
  const int t = threadIdx.x, warp = t / 32, lane = t % 32;
  const int head = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * 16;
  const float *base = QKV + size_t(b) * R * 3072;
  for (int i = t; i < 16 * kHeadDim; i += 128) {
    const int r = i / kHeadDim, d = i % kHeadDim;
    const int qi = q0 + r;
    Qs[r][d] = qi < R ? base[size_t(qi) * 3072 + head * kHeadDim + d] : 0.f;
  }
  float m[4], l[4], o[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = -INFINITY;
    l[i] = 0.f;
    o[i][0] = o[i][1] = 0.f;
  }
  const int q_last = min(q0 + 15, R - 1);
  for (int k0 = 0; k0 <= q_last; k0 += TK) {
    __syncthreads();
    for (int i = t; i < TK * (kHeadDim / 4); i += 128) {
      const int r = i / (kHeadDim / 4), c = (i % (kHeadDim / 4)) * 4;
      const int kj = k0 + r;
      float4 kv = make_float4(0, 0, 0, 0), vv = kv;
      if (kj < R) {
        kv = *reinterpret_cast<const float4 *>(base + size_t(kj) * 3072 + 1024 + head * kHeadDim + c);
        vv = *reinterpret_cast<const float4 *>(base + size_t(kj) * 3072 + 2048 + head * kHeadDim + c);
      }
      *reinterpret_cast<float4 *>(&Ks[r][c]) = kv;
      *reinterpret_cast<float4 *>(&Vs[r][c]) = vv;
    }
    __syncthreads();
    // scores: lane handles keys lane and lane+32 for the warp's 4 queries
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
      float s0 = s[i][0] * 0.125f, s1 = s[i][1] * 0.125f;
      if (k0 + lane > qi || qi >= R) s0 = -INFINITY;
      if (k0 + lane + 32 > qi || qi >= R) s1 = -INFINITY;
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
    // PV: lane handles output dims lane and lane+32
    const int kmax = min(TK, q_last - k0 + 1);
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
    if (qi < R) {
      const float inv = 1.0f / l[i];
      const size_t off = (size_t(b) * R + qi) * kDim + head * kHeadDim;
      const float v0 = o[i][0] * inv, v1 = o[i][1] * inv;
      const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
      out_hi[off + lane] = h0;
      out_hi[off + lane + 32] = h1;
      out_lo[off + lane] = __float2half_rn(v0 - __half2float(h0));
      out_lo[off + lane + 32] = __float2half_rn(v1 - __half2float(h1));
    }
  }
}

// Copy K,V of the prefill rows into the f16 cache of candidate slots [slot0, slot0 + B):
// QKV [R][3072] (single shared sequence) -> kc/vc[slot][head][pos_off + pos][64].
static __global__ void __launch_bounds__(256) ar_kv_scatter_kernel(const float *QKV, __half *kc, __half *vc, int R,
                                                            int B, int P, int slot0, int pos_off) {
  pdl_launch_dependents();
  pdl_wait();
  const int pos = blockIdx.x;
  for (int i = threadIdx.x; i < 1024; i += 256) {
    const __half k = __float2half_rn(QKV[size_t(pos) * 3072 + 1024 + i]);
    const __half v = __float2half_rn(QKV[size_t(pos) * 3072 + 2048 + i]);
    const int head = i / kHeadDim, d = i % kHeadDim;
    for (int b = 0; b < B; ++b) {
      const size_t idx = ((size_t(slot0 + b) * kHeads + head) * P + pos_off + pos) * kHeadDim + d;
      kc[idx] = k;
      vc[idx] = v;
    }
  }
}

// dst[b][:] = src[:] for b < B (broadcast last prefill row to all candidates)
static __global__ void __launch_bounds__(256) bcast_row_kernel(const float *src, float *dst, int n) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < n; i += 256) dst[size_t(b) * n + i] = src[i];
}

// transpose + convert at load time:  src f32 [K][N]  ->  dst WT [N][K]
template <typename WT>
static __global__ void transpose_convert_kernel(const float *src, WT *dst, int K, int N) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int k = k0 + i, n = n0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && n < N) ? src[size_t(k) * N + n] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int n = n0 + i, k = k0 + threadIdx.x;
    if (n < N && k < K) {
      if constexpr (sizeof(WT) == 4) dst[size_t(n) * K + k] = tile[threadIdx.x][i];
      else dst[size_t(n) * K + k] = __float2half_rn(tile[threadIdx.x][i]);
    }
  }
}
// src f32 [n] -> hi/lo f16 planes (parity-mode weights for the tensor-core row GEMMs)
static __global__ void split_f16_kernel(const float *src, __half *hi, __half *lo, size_t n) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const float v = src[i];
    const __half h = __float2half_rn(v);
    hi[i] = h;
    lo[i] = __float2half_rn(v - __half2float(h));
  }
}
template <typename WT>
static __global__ void convert_kernel(const float *src, WT *dst, size_t n) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    if constexpr (sizeof(WT) == 4) dst[i] = src[i];
    else dst[i] = __float2half_rn(src[i]);
  }
}

// Device-side top-k pre-selection of a logits row (the reference copies all 8194
// floats per candidate to the host every step, main.cpp:4767).  One CTA per candidate: radix select
// (4 x 8 bits over order-preserving keys) of the TOPK-th largest value, then every larger entry plus
// the lowest-index ties are emitted as (value, index) pairs -- unsorted; the host sampler orders
// them itself.  flags[b] = 1 when more than 256 entries tie at the threshold (the host then fetches
// the full row).  The host-side sampler proves that TOPK = 64 entries reproduce the reference's
// repetition penalty + top-k 50 + top-p + multinomial bit-exactly, or asks for the full row.
constexpr int AR_TOPK = 64;
static __global__ void __launch_bounds__(256) ar_topk_kernel(const float *logits, float *vals, int *idx, int *flags) {
  __shared__ uint32_t keys[kMelVocab];
  __shared__ unsigned int hist[256];
  __shared__ unsigned int s_prefix, s_remaining, s_cnt, s_ntie;
  __shared__ int s_ties[256];
  const int b = blockIdx.x, t = threadIdx.x;
  const float *row = logits + size_t(b) * kMelVocab;
  for (int i = t; i < kMelVocab; i += 256) {
    const uint32_t u = __float_as_uint(row[i]);
    keys[i] = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // unsigned order == float order
  }
  if (t == 0) { s_cnt = 0; s_ntie = 0; }
  uint32_t prefix = 0;
  unsigned int remaining = AR_TOPK;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    hist[t] = 0;
    __syncthreads();
    const uint32_t mask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int i = t; i < kMelVocab; i += 256) {
      const uint32_t k = keys[i];
      if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (t == 0) {
      unsigned int rem = remaining;
      int bin = 255;
      for (; bin > 0; --bin) {
        const unsigned int c = hist[bin];
        if (c >= rem) break;
        rem -= c;
      }
      s_prefix = prefix | (uint32_t(bin) << shift);
      s_remaining = rem;
    }
    __syncthreads();
    prefix = s_prefix;
    remaining = s_remaining;
    __syncthreads();
  }
  // prefix = key of the TOPK-th largest entry; `remaining` of the entries equal to it belong to the set
  for (int i = t; i < kMelVocab; i += 256) {
    const uint32_t k = keys[i];
    if (k > prefix) {
      const unsigned int pos = atomicAdd(&s_cnt, 1u);
      vals[b * AR_TOPK + pos] = row[i];
      idx[b * AR_TOPK + pos] = i;
    } else if (k == prefix) {
      const unsigned int pos = atomicAdd(&s_ntie, 1u);
      if (pos < 256u) s_ties[pos] = i;
    }
  }
  __syncthreads();
  if (t == 0) {
    const unsigned int ng = s_cnt, nt = s_ntie;
    int overflow = nt > 256u ? 1 : 0;
    if (!overflow) {
      for (unsigned int i = 1; i < nt; ++i) {  // ascending indices (almost always a single entry)
        const int v = s_ties[i];
        int j = int(i) - 1;
        for (; j >= 0 && s_ties[j] > v; --j) s_ties[j + 1] = s_ties[j];
        s_ties[j + 1] = v;
      }
      for (unsigned int i = 0; i < remaining && ng + i < AR_TOPK; ++i) {
        vals[b * AR_TOPK + ng + i] = row[s_ties[i]];
        idx[b * AR_TOPK + ng + i] = s_ties[i];
      }
    }
    flags[b] = overflow;
  }
}

// tokens of one decode step, passed by value (max_batch <= 64)
struct TokenArgs {
  int tok[64];
};
static __global__ void set_tokens_kernel(int *dst, TokenArgs t, int n) {
  if (threadIdx.x < n) dst[threadIdx.x] = t.tok[threadIdx.x];
}

}  // namespace tts
