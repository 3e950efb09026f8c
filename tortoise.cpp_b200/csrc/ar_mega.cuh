// ar_mega.cuh -- the whole AR decode step as ONE persistent cooperative kernel.
//
// Same math as the per-op path (wsgemv.cuh + ar_attn_decode_kernel; reference graph
// autoregressive_graph(fake_inputs=false), main.cpp:2668-3029), restructured for B200:
//   * grid = one CTA per SM, resident for the whole step; the 152 dependent launches of the
//     per-op path become phases separated by a device-wide barrier (release/acquire atomics
//     in global memory), so a phase boundary costs ~1 us instead of a kernel boundary;
//   * every CTA owns a fixed row slice of every weight matrix; thread 0 streams those slices,
//     in the order the step consumes them, through ONE shared-memory ring (6 x 16 KB) with
//     1-D TMA bulk copies.  The ring is independent of the phase structure: while the grid
//     waits at a barrier, the next phases' weights are already landing in shared memory, so
//     HBM keeps streaming across the dependency stalls;
//   * activations cross CTAs only through L2 (ld.global.cg), 4 KB per candidate per phase.
// Phases per layer: LN1+QKV(+f16 round trip, KV append) | attention | c_proj+residual |
// LN2+FC+GELU16 | mlp c_proj+residual; then double-LN + lm_head.
#pragma once
#include "ar_kernels.cuh"
#include "wsgemv.cuh"

namespace tts {

constexpr int MG_STAGES = 6;
constexpr int MG_CONSUMERS = 256;            // 8 consumer warps
constexpr int MG_THREADS = MG_CONSUMERS + 32;  // + 1 dedicated weight-stream producer warp

struct MegaLayer {
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
  const void *w_qkv, *w_proj, *w_fc, *w_proj2;
  const float *b_qkv, *b_proj, *b_fc, *b_proj2;
};

struct MegaArgs {
  const MegaLayer *layers;  // [30], device
  const float *lnf_w, *lnf_b, *lm0_w, *lm0_b, *lm_b;
  const void *lm_w;
  const float *mel_emb, *mel_pos;
  const int *tokens;
  float *h, *q, *attn, *m, *logits;
  __half *kc, *vc;  // [30][Bmax][16][P][64]
  unsigned int *bar;  // [2]: arrival count, generation
  int B, Bmax, P, n_past, pos_id;
  long long *dbg;  // optional [CTA0 thread0] clock64 trace (TTS_MEGA_TRACE=1), else null
};

__host__ __device__ inline size_t mega_smem_bytes() {
  return size_t(MG_STAGES) * GV_STAGE_BYTES + 256 /*barriers*/ + GV_MAX_ROWS_PER_CTA * 8 * 2 * sizeof(float) +
         64 * sizeof(double) + (1024 + 384) * sizeof(float) /*attention scratch*/;
}

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// device-wide barrier: all CTAs are co-resident (cooperative launch).  Generation counting,
// release on arrival / acquire on departure, so global writes before the barrier are visible
// to every thread after it.
__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int nblocks, unsigned int &gen) {
  // (consumer warps only: named barrier 1; the producer warp never joins a barrier)
  asm volatile("bar.sync 1, 256;" ::: "memory");  // the phase's global writes happen-before thread 0's release
  if (threadIdx.x == 0) {
    gen += 1;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    const unsigned int target = gen * nblocks;
    while (ld_acquire_u32(bar) < target) {
    }
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");
}

template <typename WT, int BT>
static __global__ void __launch_bounds__(MG_THREADS, 1) ar_decode_mega_kernel(MegaArgs a) {
  constexpr int E = WTraits<WT>::kElemsPer16B;
  constexpr int KS = 32 * 4 * E;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *ring = smem;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + MG_STAGES * GV_STAGE_BYTES);
  uint64_t *empty = full + MG_STAGES;
  float *partial = reinterpret_cast<float *>(smem + MG_STAGES * GV_STAGE_BYTES + 256);
  double *red = reinterpret_cast<double *>(partial + GV_MAX_ROWS_PER_CTA * 8 * 2);
  float *att = reinterpret_cast<float *>(red + 64);  // [P] scores + 2*64... sized 1024 + 256

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x, cta = blockIdx.x;
  const int B = a.B;
  const int n_groups = (B + BT - 1) / BT;
  unsigned int gen = 0;

  if (tid == 0) {
    for (int s = 0; s < MG_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], GV_WARPS);
    }
    fence_barrier_init();
  }
  __syncthreads();

  // ---------------- weight stream (thread 0): segments in consumption order -----------------
  // segment id = (layer * 4 + op) for layer < 30, then 120 = lm_head; each repeated n_groups x
  auto seg_shape = [&](int sid, int &N, int &K, const void *&W) {
    if (sid >= 120) { N = kMelVocab; K = kDim; W = a.lm_w; return; }
    const MegaLayer &l = a.layers[sid >> 2];
    switch (sid & 3) {
      case 0: N = 3072; K = kDim; W = l.w_qkv; break;
      case 1: N = kDim; K = kDim; W = l.w_proj; break;
      case 2: N = kFF; K = kDim; W = l.w_fc; break;
      default: N = kDim; K = kFF; W = l.w_proj2; break;
    }
  };
  auto slice = [&](int N, int &row0, int &rows) {
    const int base = N / G, rem = N % G;
    rows = base + (cta < rem ? 1 : 0);
    row0 = cta * base + min(cta, rem);
  };
  // producer cursor
  int p_seg = 0, p_grp = 0, p_stage = 0;
  long p_it = 0;
  const unsigned char *p_src = nullptr;
  size_t p_total = 0;
  int p_nstages = -1;
  auto p_load_seg = [&]() {
    int N, K, row0, rows;
    const void *W;
    seg_shape(p_seg, N, K, W);
    slice(N, row0, rows);
    const size_t row_bytes = size_t(K) * sizeof(WT);
    p_src = reinterpret_cast<const unsigned char *>(W) + size_t(row0) * row_bytes;
    p_total = size_t(rows) * row_bytes;
    p_nstages = int((p_total + GV_STAGE_BYTES - 1) / GV_STAGE_BYTES);
  };
  auto p_done = [&]() { return p_seg > 120; };
  // issue one stage; blocking => wait for the slot, else give up if it is still in use
  auto produce_one = [&](bool blocking) -> bool {
    if (p_done()) return false;
    if (p_nstages < 0) p_load_seg();
    while (p_stage >= p_nstages) {  // (segments with zero rows are skipped)
      p_stage = 0;
      if (++p_grp >= n_groups) { p_grp = 0; ++p_seg; if (p_done()) return false; p_load_seg(); }
    }
    const int slot = int(p_it % MG_STAGES);
    const uint32_t ph = uint32_t((p_it / MG_STAGES) & 1) ^ 1u;
    if (blocking) mbar_wait(&empty[slot], ph);
    else if (!mbar_test_wait(&empty[slot], ph)) return false;
    const size_t off = size_t(p_stage) * GV_STAGE_BYTES;
    const uint32_t bytes = uint32_t(min(size_t(GV_STAGE_BYTES), p_total - off));
    mbar_arrive_expect_tx(&full[slot], bytes);
    bulk_g2s(ring + size_t(slot) * GV_STAGE_BYTES, p_src + off, bytes, &full[slot]);
    ++p_stage;
    ++p_it;
    return true;
  };
  int dbg_n = 0;
  auto trace = [&](int tag) {
    if (a.dbg && cta == 0 && tid == 0 && dbg_n < 1000) { a.dbg[2 * dbg_n] = tag; a.dbg[2 * dbg_n + 1] = clock64(); ++dbg_n; }
  };
  long c_it = 0;  // consumer stage counter (same order as the producer)
  if (warp == MG_CONSUMERS / 32) {
    // dedicated producer warp: streams every weight slice of the step, in consumption order,
    // bounded only by ring capacity -- never on the consumers' critical path
    if (lane == 0)
      while (produce_one(true)) {
      }
    return;
  }

  // ---------------- phase 0: h[b] = mel_emb[tok[b]] + mel_pos[pos] ----------------------------
  for (int b = cta; b < B; b += G) {
    const int tok = a.tokens[b];
    const float4 e = reinterpret_cast<const float4 *>(a.mel_emb + size_t(tok) * kDim)[tid];
    const float4 p = reinterpret_cast<const float4 *>(a.mel_pos + size_t(a.pos_id) * kDim)[tid];
    reinterpret_cast<float4 *>(a.h + size_t(b) * kDim)[tid] = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
  }
  grid_barrier(a.bar, G, gen);

  // ---------------- generic GEMV phase (consumer side of wsgemv_kernel) -----------------------
  auto gemv_phase = [&](int N, int K, const float *bias, const float *in, float *out, int pro, int epi,
                        const float *ln_w, const float *ln_b, const float *ln2_w, const float *ln2_b,
                        __half *kcache, __half *vcache) {
    const int wpr = K / KS, rps = GV_WARPS / wpr;
    int row0, rows_cta;
    slice(N, row0, rows_cta);
    const size_t row_bytes = size_t(K) * sizeof(WT);
    const int n_stages = int((size_t(rows_cta) * row_bytes + GV_STAGE_BYTES - 1) / GV_STAGE_BYTES);
    const int ks = warp % wpr, rsub = warp / wpr;
    const int kvb = kHeads * a.P * kHeadDim;
    for (int g = 0; g < n_groups; ++g) {
      const int b0 = g * BT;
      // ---- everything this phase needs from L2 is requested up front (one round trip):
      //      the LayerNorm row (statistics), this warp's activation slice, bias and residual.
      float mean1[BT], rstd1[BT], mean2[BT], rstd2[BT];
      float4 lnrow[BT];
      float xr[BT][4][E];
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        mean1[b] = 0.f; rstd1[b] = 1.f; mean2[b] = 0.f; rstd2[b] = 1.f;
        const bool live = (b0 + b) < B;
        const float *x = in + size_t(live ? b0 + b : 0) * K;
        lnrow[b] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pro != PRO_NONE && live) lnrow[b] = __ldcg(reinterpret_cast<const float4 *>(x) + tid);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = ks * KS + j * (32 * E) + lane * E;
#pragma unroll
          for (int e4 = 0; e4 < E; e4 += 4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) v = __ldcg(reinterpret_cast<const float4 *>(x + k + e4));
            xr[b][j][e4 + 0] = v.x; xr[b][j][e4 + 1] = v.y; xr[b][j][e4 + 2] = v.z; xr[b][j][e4 + 3] = v.w;
          }
        }
      }
      trace(1);  // loads issued
      float ep_bias = 0.f, ep_old = 0.f;  // epilogue operands of output element `tid`
      if (tid < rows_cta * BT) {
        const int r = tid / BT, b = tid % BT;
        ep_bias = bias[row0 + r];
        if (epi == EPI_RESID && b0 + b < B) ep_old = __ldcg(out + size_t(b0 + b) * N + row0 + r);
      }
      if (pro != PRO_NONE) {
        // LayerNorm statistics, single pass with double accumulation (sum, sum of squares):
        // mean is narrowed to float like ggml's (ggml.c:11935-11955); the variance differs from
        // the two-pass float/double form by O(1e-7) relative -- far below the fp16 chaos floor.
        double *rb = red + ((c_it + g) & 1) * 32;
#pragma unroll
        for (int b = 0; b < BT; ++b) {
          const float4 v = lnrow[b];
          double s1 = double(v.x) + double(v.y) + double(v.z) + double(v.w);
          double s2 = double(v.x) * v.x + double(v.y) * v.y + double(v.z) * v.z + double(v.w) * v.w;
          s1 = warp_sum_d(s1);
          s2 = warp_sum_d(s2);
          if (lane == 0) { rb[(warp * BT + b) * 2] = s1; rb[(warp * BT + b) * 2 + 1] = s2; }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
        for (int b = 0; b < BT; ++b) {
          double t1 = 0, t2 = 0;
          for (int w = 0; w < GV_WARPS; ++w) { t1 += rb[(w * BT + b) * 2]; t2 += rb[(w * BT + b) * 2 + 1]; }
          const double md = t1 / K;
          const float mean = float(md);
          double var = t2 / K - 2.0 * md * double(mean) + double(mean) * double(mean);  // E[(x - mean_f)^2]
          if (var < 0) var = 0;
          mean1[b] = mean;
          rstd1[b] = 1.0f / sqrtf(float(var) + 1e-5f);
        }
        if (pro == PRO_LN2) {
          // second LayerNorm (lm_head.0) on y = LN(x) ln_f: once per step, plain two-pass form
#pragma unroll
          for (int b = 0; b < BT; ++b) {
            const float4 v = lnrow[b];
            const float4 w4 = reinterpret_cast<const float4 *>(ln_w)[tid];
            const float4 b4 = reinterpret_cast<const float4 *>(ln_b)[tid];
            const float y0 = (v.x - mean1[b]) * rstd1[b] * w4.x + b4.x, y1 = (v.y - mean1[b]) * rstd1[b] * w4.y + b4.y;
            const float y2 = (v.z - mean1[b]) * rstd1[b] * w4.z + b4.z, y3 = (v.w - mean1[b]) * rstd1[b] * w4.w + b4.w;
            double t = double(y0) + double(y1) + double(y2) + double(y3);
            t = warp_sum_d(t);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (lane == 0) rb[warp] = t;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            double tot = 0;
            for (int w = 0; w < GV_WARPS; ++w) tot += rb[w];
            const float m2 = float(tot / K);
            const float e0 = y0 - m2, e1 = y1 - m2, e2 = y2 - m2, e3 = y3 - m2;
            double t2 = double(e0 * e0) + double(e1 * e1) + double(e2 * e2) + double(e3 * e3);
            t2 = warp_sum_d(t2);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (lane == 0) rb[warp] = t2;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            tot = 0;
            for (int w = 0; w < GV_WARPS; ++w) tot += rb[w];
            mean2[b] = m2;
            rstd2[b] = 1.0f / sqrtf(float(tot / K) + 1e-5f);
          }
        }
#pragma unroll
        for (int b = 0; b < BT; ++b) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = ks * KS + j * (32 * E) + lane * E;
#pragma unroll
            for (int e4 = 0; e4 < E; e4 += 4) {
              const float4 w4 = *reinterpret_cast<const float4 *>(ln_w + k + e4);
              const float4 b4 = *reinterpret_cast<const float4 *>(ln_b + k + e4);
              float vv[4] = {xr[b][j][e4], xr[b][j][e4 + 1], xr[b][j][e4 + 2], xr[b][j][e4 + 3]};
              const float ww[4] = {w4.x, w4.y, w4.z, w4.w}, bb4[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) vv[q] = (vv[q] - mean1[b]) * rstd1[b] * ww[q] + bb4[q];
              if (pro == PRO_LN2) {
                const float4 w5 = *reinterpret_cast<const float4 *>(ln2_w + k + e4);
                const float4 b5 = *reinterpret_cast<const float4 *>(ln2_b + k + e4);
                const float w2[4] = {w5.x, w5.y, w5.z, w5.w}, b2[4] = {b5.x, b5.y, b5.z, b5.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) vv[q] = (vv[q] - mean2[b]) * rstd2[b] * w2[q] + b2[q];
              }
              const bool live = (b0 + b) < B;
#pragma unroll
              for (int q = 0; q < 4; ++q) xr[b][j][e4 + q] = live ? vv[q] : 0.f;
            }
          }
        }
      }
      trace(2);  // prologue done (LN applied)
      for (int s = 0; s < n_stages; ++s, ++c_it) {
        const int slot = int(c_it % MG_STAGES);
        mbar_wait(&full[slot], uint32_t((c_it / MG_STAGES) & 1));
        const int r = s * rps + rsub;
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
              for (int qd = 0; qd < 4; ++qd) {
                const float2 f = __half22float2(h2[qd]);
                wf[2 * qd] = f.x;
                wf[2 * qd + 1] = f.y;
              }
            }
#pragma unroll
            for (int b = 0; b < BT; ++b)
#pragma unroll
              for (int e = 0; e < E; ++e) acc[b] = fmaf(wf[e], xr[b][j][e], acc[b]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
        if (r < rows_cta) {
#pragma unroll
          for (int b = 0; b < BT; ++b) {
            const float v = warp_sum(acc[b]);
            if (lane == 0) partial[(r * 8 + ks) * 2 + b] = v;
          }
        }
      }
      trace(3);  // stage loop done
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (tid < rows_cta * BT) {  // rows_cta * BT <= 128 < MG_THREADS: one output element per thread
        const int r = tid / BT, b = tid % BT;
        if (b0 + b < B) {
          float v = 0.f;
          for (int w = 0; w < wpr; ++w) v += partial[(r * 8 + w) * 2 + b];
          const int n = row0 + r;
          v += ep_bias;
          const int bb = b0 + b;
          if (epi == EPI_STORE) {
            out[size_t(bb) * N + n] = v;
          } else if (epi == EPI_RESID) {
            out[size_t(bb) * N + n] = ep_old + v;
          } else if (epi == EPI_GELU16) {
            out[size_t(bb) * N + n] = gelu16(v);
          } else {
            const __half hv = __float2half_rn(v);
            const int which = n >> 10, c = n & 1023;
            if (which == 0) {
              out[size_t(bb) * kDim + c] = __half2float(hv);
            } else {
              __half *cache = which == 1 ? kcache : vcache;
              const int head = c >> 6, d = c & 63;
              cache[size_t(bb) * kvb + (size_t(head) * a.P + a.n_past) * kHeadDim + d] = hv;
            }
          }
        }
      }
      trace(4);  // epilogue stored
      if (g + 1 < n_groups) asm volatile("bar.sync 1, 256;" ::: "memory");  // partial[] is reused by the next candidate group
    }
  };

  // ---------------- attention phase: items (b, head) round-robin over CTAs --------------------
  auto attn_phase = [&](const __half *kc, const __half *vc) {
    const int n = a.n_past + 1;
    float *sc = att;            // [n <= 1024] scores
    float *qs = att + 1024;     // [64] query
    float *pp = att + 1088;     // [4][64] partial outputs
    float *redf = att + 1344;   // [8] block reductions
    for (int item = cta; item < B * kHeads; item += G) {
      const int b = item / kHeads, head = item % kHeads;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (tid < kHeadDim) qs[tid] = __ldcg(a.q + size_t(b) * kDim + head * kHeadDim + tid);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const __half *K = kc + (size_t(b) * kHeads + head) * size_t(a.P) * kHeadDim;
      const __half *V = vc + (size_t(b) * kHeads + head) * size_t(a.P) * kHeadDim;
      float lmax = -INFINITY;
      for (int j = tid; j < n; j += MG_CONSUMERS) {
        const uint4 *kr = reinterpret_cast<const uint4 *>(K + size_t(j) * kHeadDim);
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 u = __ldcg(kr + c);
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
      asm volatile("bar.sync 1, 256;" ::: "memory");
      float mx = redf[0];
      for (int w = 1; w < GV_WARPS; ++w) mx = fmaxf(mx, redf[w]);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      float lsum = 0.f;
      for (int j = tid; j < n; j += MG_CONSUMERS) {
        const float p = expf(sc[j] - mx);
        sc[j] = p;
        lsum += p;
      }
      lsum = warp_sum(lsum);
      if (lane == 0) redf[warp] = lsum;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      double tot = 0;
      for (int w = 0; w < GV_WARPS; ++w) tot += double(redf[w]);
      const float inv = float(1.0 / tot);
      const int d = tid % kHeadDim, part = tid / kHeadDim;  // 4 key partitions x 64 dims
      float acc = 0.f;
      const unsigned short *Vu = reinterpret_cast<const unsigned short *>(V);
      for (int j = part; j < n; j += 4)
        acc = fmaf(sc[j] * inv, __half2float(__ushort_as_half(__ldcg(Vu + size_t(j) * kHeadDim + d))), acc);
      pp[part * kHeadDim + d] = acc;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (tid < kHeadDim)
        a.attn[size_t(b) * kDim + head * kHeadDim + tid] = (pp[tid] + pp[64 + tid]) + (pp[128 + tid] + pp[192 + tid]);
    }
  };

  // ---------------- the 30 layers ---------------------------------------------------------------
  const size_t layer_kv = size_t(a.Bmax) * kHeads * a.P * kHeadDim;
  for (int li = 0; li < kLayers; ++li) {
    const MegaLayer &l = a.layers[li];
    __half *kc = a.kc + size_t(li) * layer_kv, *vc = a.vc + size_t(li) * layer_kv;
    trace(10);
    gemv_phase(3072, kDim, l.b_qkv, a.h, a.q, PRO_LN, EPI_QKV, l.ln1_w, l.ln1_b, nullptr, nullptr, kc, vc);
    grid_barrier(a.bar, G, gen);
    trace(11);
    attn_phase(kc, vc);
    trace(12);
    grid_barrier(a.bar, G, gen);
    trace(13);
    gemv_phase(kDim, kDim, l.b_proj, a.attn, a.h, PRO_NONE, EPI_RESID, nullptr, nullptr, nullptr, nullptr, nullptr,
               nullptr);
    grid_barrier(a.bar, G, gen);
    gemv_phase(kFF, kDim, l.b_fc, a.h, a.m, PRO_LN, EPI_GELU16, l.ln2_w, l.ln2_b, nullptr, nullptr, nullptr, nullptr);
    grid_barrier(a.bar, G, gen);
    gemv_phase(kDim, kFF, l.b_proj2, a.m, a.h, PRO_NONE, EPI_RESID, nullptr, nullptr, nullptr, nullptr, nullptr,
               nullptr);
    grid_barrier(a.bar, G, gen);
  }
  gemv_phase(kMelVocab, kDim, a.lm_b, a.h, a.logits, PRO_LN2, EPI_STORE, a.lnf_w, a.lnf_b, a.lm0_w, a.lm0_b, nullptr,
             nullptr);
}

}  // namespace tts
