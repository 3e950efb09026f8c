// wsgemv.cuh -- weight-streaming GEMV for the AR decode step (the HBM-bound hot loop).
//
// Replaces, for one decode position of B candidates, the reference's
//   ggml_mul_mat(ggml_cont(ggml_transpose(W)), cur) + ggml_add(bias) [+ cpy F16 / gelu / residual]
// chains of autoregressive_graph (main.cpp:2769-2790 c_attn, 2888-2912 c_proj, 2934-2947
// c_fc+gelu, 2955-2981 mlp c_proj, 2985-3007 ln_f/lm_head) -- including the LayerNorm that
// precedes them (main.cpp:2727-2750, 2918-2932).
//
// Design (B200-first):
//  * weights are stored ONCE at load time as [N][K] (out-major) in f32 or f16, so each
//    CTA owns a contiguous byte range = its slice of output rows;
//  * one persistent-size grid (1 CTA per SM); a producer warp streams the CTA's slice
//    through a 6 x 16 KB shared-memory ring with 1-D TMA bulk copies (cp.async.bulk ->
//    UBLKCP) completing on mbarriers; 8 consumer warps each own a fixed 2 KB column slice
//    of every stage, keep the matching activation slice in registers and reduce with warp
//    shuffles;
//  * the kernel is PDL-launched: the producer starts pulling weights immediately
//    (weights never depend on the previous kernel), only the consumers execute
//    griddepcontrol.wait.  The weight stream of op N+1 therefore overlaps the dependency
//    latency of op N: HBM stays busy across the ~150 dependent ops of a decode step;
//  * candidates are processed BT at a time against register-resident activations; the
//    slice is re-streamed (from L2) for further candidate groups.
#pragma once
#include "common.cuh"

namespace tts {

constexpr int GV_WARPS = 8;
constexpr int GV_THREADS = GV_WARPS * 32;  // thread 0 doubles as the TMA producer
constexpr int GV_STAGE_BYTES = 16384;            // 8 consumer warps x 2 KB
constexpr int GV_STAGES = 6;
constexpr int GV_MAX_ROWS_PER_CTA = 64;

enum GemvPrologue { PRO_NONE = 0, PRO_LN = 1, PRO_LN2 = 2 };
enum GemvEpilogue { EPI_STORE = 0, EPI_QKV = 1, EPI_RESID = 2, EPI_GELU16 = 3 };

struct GemvArgs {
  const void *W;       // [N][K] row-major
  const float *bias;   // [N]
  const float *in;     // [B][K] f32
  float *out;          // EPI_STORE/GELU16: [B][N]; EPI_RESID: [B][N] (+=); EPI_QKV: q [B][1024]
  const float *ln_w, *ln_b;    // PRO_LN / first LN of PRO_LN2
  const float *ln2_w, *ln2_b;  // second LN of PRO_LN2
  __half *kcache, *vcache;     // EPI_QKV: this layer's [Bmax][16][P][64]
  const int *state;            // device step state: state[0] = n_past (KV write position)
  int N, K, B;
  int pro, epi;
  int kv_b_stride;  // elements between candidates in the kv cache = 16*P*64
};

__host__ __device__ inline size_t gemv_smem_bytes() {
  return size_t(GV_STAGES) * GV_STAGE_BYTES + 128 /*barriers*/ +
         GV_MAX_ROWS_PER_CTA * 8 * 2 * sizeof(float) /*partials*/ + 8 * 2 * 2 * sizeof(double);
}

template <typename WT>
struct WTraits;
template <>
struct WTraits<float> {
  static constexpr int kElemsPer16B = 4;
};
template <>
struct WTraits<__half> {
  static constexpr int kElemsPer16B = 8;
};

template <typename WT, int BT>
__global__ void __launch_bounds__(GV_THREADS, 2) wsgemv_kernel(GemvArgs a) {
  constexpr int E = WTraits<WT>::kElemsPer16B;  // weights per 16-byte load
  constexpr int KS = 32 * 4 * E;                // K elements owned by one warp (2 KB)
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *ring = smem;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + GV_STAGES * GV_STAGE_BYTES);
  uint64_t *empty = full + GV_STAGES;
  float *partial = reinterpret_cast<float *>(smem + GV_STAGES * GV_STAGE_BYTES + 128);
  double *red = reinterpret_cast<double *>(partial + GV_MAX_ROWS_PER_CTA * 8 * 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = a.N, K = a.K;
  const int wpr = K / KS;          // warps cooperating on one row
  const int rps = GV_WARPS / wpr;  // rows per stage
  const int G = gridDim.x;
  const int base = N / G, rem = N % G;
  const int cta = blockIdx.x;
  const int rows_cta = base + (cta < rem ? 1 : 0);
  const int row0 = cta * base + min(cta, rem);
  const int n_stages = (rows_cta + rps - 1) / rps;
  const size_t row_bytes = size_t(K) * sizeof(WT);
  const int n_groups = (a.B + BT - 1) / BT;

  if (tid == 0) {
    for (int s = 0; s < GV_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], GV_WARPS);
    }
    fence_barrier_init();
  }
  __syncthreads();
  pdl_launch_dependents();

  // ---------------- producer role (thread 0): weights never depend on the previous kernel,
  // so the ring is filled BEFORE griddepcontrol.wait; refills are issued from the main loop.
  const unsigned char *wsrc = reinterpret_cast<const unsigned char *>(a.W) + size_t(row0) * row_bytes;
  const size_t wtotal = size_t(rows_cta) * row_bytes;
  const int total_iters = n_groups * n_stages;
  auto produce = [&](int pit) {  // fill the slot of global iteration pit
    const int slot = pit % GV_STAGES;
    const uint32_t ph = (pit / GV_STAGES) & 1;
    mbar_wait(&empty[slot], ph ^ 1);
    const size_t off = size_t(pit % n_stages) * GV_STAGE_BYTES;
    const uint32_t bytes = uint32_t(min(size_t(GV_STAGE_BYTES), wtotal - off));
    mbar_arrive_expect_tx(&full[slot], bytes);
    bulk_g2s(ring + size_t(slot) * GV_STAGE_BYTES, wsrc + off, bytes, &full[slot]);
  };
  if (tid == 0)
    for (int pit = 0; pit < min(GV_STAGES, total_iters); ++pit) produce(pit);

  // ---------------- consumers ---------------------------------------------------------------
  pdl_wait();  // activations of the previous op are complete and visible from here on
  const int ks = warp % wpr;    // which K slice of the row this warp owns
  const int rsub = warp / wpr;  // which row of the stage
  const int n_past = a.state ? a.state[0] : 0;

  int it = 0;
  for (int g = 0; g < n_groups; ++g) {
    const int b0 = g * BT;
    // ---- prologue: (optional) LayerNorm statistics, double accumulation like
    //      ggml_compute_forward_norm_f32 (ggml.c:11905-11958); K == 1024 here.
    float mean1[BT], rstd1[BT], mean2[BT], rstd2[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) { mean1[b] = 0.f; rstd1[b] = 1.f; mean2[b] = 0.f; rstd2[b] = 1.f; }
    if (a.pro != PRO_NONE) {
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        if (b0 + b >= a.B) break;
        const float *x = a.in + size_t(b0 + b) * K;
        const float4 v = reinterpret_cast<const float4 *>(x)[tid];  // 256 thr x 4 = 1024
        double s = double(v.x) + double(v.y) + double(v.z) + double(v.w);
        s = warp_sum_d(s);
        if (lane == 0) red[warp * 2 + 0] = s;
        __syncthreads();
        double tot = 0;
        for (int w = 0; w < GV_WARPS; ++w) tot += red[w * 2];
        const float mean = float(tot / K);
        __syncthreads();
        const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
        double s2 = double(d0 * d0) + double(d1 * d1) + double(d2 * d2) + double(d3 * d3);
        s2 = warp_sum_d(s2);
        if (lane == 0) red[warp * 2 + 0] = s2;
        __syncthreads();
        tot = 0;
        for (int w = 0; w < GV_WARPS; ++w) tot += red[w * 2];
        const float var = float(tot / K);
        const float rstd = 1.0f / sqrtf(var + 1e-5f);
        __syncthreads();
        mean1[b] = mean;
        rstd1[b] = rstd;
        if (a.pro == PRO_LN2) {
          const float4 w4 = reinterpret_cast<const float4 *>(a.ln_w)[tid];
          const float4 b4 = reinterpret_cast<const float4 *>(a.ln_b)[tid];
          const float y0 = d0 * rstd * w4.x + b4.x, y1 = d1 * rstd * w4.y + b4.y;
          const float y2 = d2 * rstd * w4.z + b4.z, y3 = d3 * rstd * w4.w + b4.w;
          double t = double(y0) + double(y1) + double(y2) + double(y3);
          t = warp_sum_d(t);
          if (lane == 0) red[warp * 2 + 0] = t;
          __syncthreads();
          tot = 0;
          for (int w = 0; w < GV_WARPS; ++w) tot += red[w * 2];
          const float m2 = float(tot / K);
          __syncthreads();
          const float e0 = y0 - m2, e1 = y1 - m2, e2 = y2 - m2, e3 = y3 - m2;
          double t2 = double(e0 * e0) + double(e1 * e1) + double(e2 * e2) + double(e3 * e3);
          t2 = warp_sum_d(t2);
          if (lane == 0) red[warp * 2 + 0] = t2;
          __syncthreads();
          tot = 0;
          for (int w = 0; w < GV_WARPS; ++w) tot += red[w * 2];
          mean2[b] = m2;
          rstd2[b] = 1.0f / sqrtf(float(tot / K) + 1e-5f);
          __syncthreads();
        }
      }
    }
    // ---- register-resident activation slice of this warp: xr[b][j][e]
    float xr[BT][4][E];
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      const bool live = (b0 + b) < a.B;
      const float *x = a.in + size_t(live ? b0 + b : 0) * K;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = ks * KS + j * (32 * E) + lane * E;
#pragma unroll
        for (int e4 = 0; e4 < E; e4 += 4) {
          float4 v = *reinterpret_cast<const float4 *>(x + k + e4);
          if (a.pro != PRO_NONE) {
            const float4 w4 = *reinterpret_cast<const float4 *>(a.ln_w + k + e4);
            const float4 b4 = *reinterpret_cast<const float4 *>(a.ln_b + k + e4);
            v.x = (v.x - mean1[b]) * rstd1[b] * w4.x + b4.x;
            v.y = (v.y - mean1[b]) * rstd1[b] * w4.y + b4.y;
            v.z = (v.z - mean1[b]) * rstd1[b] * w4.z + b4.z;
            v.w = (v.w - mean1[b]) * rstd1[b] * w4.w + b4.w;
            if (a.pro == PRO_LN2) {
              const float4 w5 = *reinterpret_cast<const float4 *>(a.ln2_w + k + e4);
              const float4 b5 = *reinterpret_cast<const float4 *>(a.ln2_b + k + e4);
              v.x = (v.x - mean2[b]) * rstd2[b] * w5.x + b5.x;
              v.y = (v.y - mean2[b]) * rstd2[b] * w5.y + b5.y;
              v.z = (v.z - mean2[b]) * rstd2[b] * w5.z + b5.z;
              v.w = (v.w - mean2[b]) * rstd2[b] * w5.w + b5.w;
            }
          }
          if (!live) v = make_float4(0.f, 0.f, 0.f, 0.f);
          xr[b][j][e4 + 0] = v.x;
          xr[b][j][e4 + 1] = v.y;
          xr[b][j][e4 + 2] = v.z;
          xr[b][j][e4 + 3] = v.w;
        }
      }
    }

    // ---- main loop over the ring
    for (int s = 0; s < n_stages; ++s, ++it) {
      const int slot = it % GV_STAGES;
      const uint32_t ph = (it / GV_STAGES) & 1;
      // refill the slot every warp released one iteration ago
      if (tid == 0 && it >= 1 && it - 1 + GV_STAGES < total_iters) produce(it - 1 + GV_STAGES);
      __syncwarp();
      mbar_wait(&full[slot], ph);
      const int r = s * rps + rsub;  // row inside this CTA's slice
      float acc[BT];
#pragma unroll
      for (int b = 0; b < BT; ++b) acc[b] = 0.f;
      if (r < rows_cta) {
        const unsigned char *wp = ring + size_t(slot) * GV_STAGE_BYTES + warp * 2048 + lane * 16;
        uint4 wv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) wv[j] = *reinterpret_cast<const uint4 *>(wp + j * 512);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float wf[E];
          if constexpr (sizeof(WT) == 4) {
            wf[0] = __uint_as_float(wv[j].x);
            wf[1] = __uint_as_float(wv[j].y);
            wf[2] = __uint_as_float(wv[j].z);
            wf[3] = __uint_as_float(wv[j].w);
          } else {
            const __half2 *h2 = reinterpret_cast<const __half2 *>(&wv[j]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 f = __half22float2(h2[q]);
              wf[2 * q] = f.x;
              wf[2 * q + 1] = f.y;
            }
          }
#pragma unroll
          for (int b = 0; b < BT; ++b)
#pragma unroll
            for (int e = 0; e < E; ++e) acc[b] = fmaf(wf[e], xr[b][j][e], acc[b]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);  // smem slot can be refilled
      if (r < rows_cta) {
#pragma unroll
        for (int b = 0; b < BT; ++b) {
          const float v = warp_sum(acc[b]);
          if (lane == 0) partial[(r * 8 + ks) * 2 + b] = v;
        }
      }
    }
    __syncthreads();

    // ---- epilogue: one thread per (row, candidate)
    for (int i = tid; i < rows_cta * BT; i += GV_WARPS * 32) {
      const int r = i / BT, b = i % BT;
      if (b0 + b >= a.B) continue;
      float v = 0.f;
      for (int w = 0; w < wpr; ++w) v += partial[(r * 8 + w) * 2 + b];
      const int n = row0 + r;
      v += a.bias[n];
      const int bb = b0 + b;
      if (a.epi == EPI_STORE) {
        a.out[size_t(bb) * N + n] = v;
      } else if (a.epi == EPI_RESID) {
        a.out[size_t(bb) * N + n] += v;
      } else if (a.epi == EPI_GELU16) {
        a.out[size_t(bb) * N + n] = gelu16(v);
      } else {  // EPI_QKV: q|k|v = rows 0-1023|1024-2047|2048-3071, f16 round trip (A-2)
        const __half hv = __float2half_rn(v);
        const int which = n >> 10, c = n & 1023;
        if (which == 0) {
          a.out[size_t(bb) * kDim + c] = __half2float(hv);
        } else {
          __half *cache = which == 1 ? a.kcache : a.vcache;
          const int head = c >> 6, d = c & 63;
          cache[size_t(bb) * a.kv_b_stride + (size_t(head) * (a.kv_b_stride / (kHeads * kHeadDim)) + n_past) * kHeadDim + d] = hv;
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace tts
