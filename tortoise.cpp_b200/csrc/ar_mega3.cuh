// ar_mega3.cuh -- AR decode step as one persistent kernel, third generation (f16 weights):
// the K = 1024 GEMV phases run on the tensor cores (the K = 4096 phase too for 4 candidates).
//
// Same math and the same cross-CTA protocol as ar_mega2.cuh (reference graph
// autoregressive_graph(fake_inputs=false), main.cpp:2668-3029).  What changed, and why
// (per-CTA globaltimer stamps, profiles/r01_decode_step.md): with the exchange at its floor
// (~1.5 us per all-to-all of a 1024-vector through L2, tools/xchg_bench.cu) the CUDA-core GEMV
// was the largest remaining term of the critical path -- 0.8 us fixed + 0.3 us per 16 KB stage
// although every stage of a phase is resident in shared memory when the phase starts: 64
// HADD2.F32/FFMA per thread and stage in two dependent chains, a try_wait/arrive per stage, a
// 5-level shuffle tree per row.  Here
//   * the producer warp lands every weight row with its own bulk copy at a pitch of K*2 + 16
//     bytes, so ldmatrix reads 16 x 16 tiles of the row-major [N][K] matrix without bank conflicts;
//   * the activations are split once per phase into two f16 planes (x = hi + lo, |error| <=
//     2^-22 |x|) and sit in shared memory as the B operand: columns 0..3 = hi of up to 4
//     candidates, 4..7 = lo, so ONE mma.sync.m16n8k16 (f32 accumulate; f16 x f16 products are
//     exact) covers both planes of every candidate.  The MLP hidden vector is f16-exact after
//     the reference's fp16 GELU table and needs no lo plane;
//   * each warp owns K/8 of the reduction for ALL rows of the phase (8 or 32 k-steps), waits once
//     for the phase's stages, and the eight partial tiles are summed through shared memory in a
//     fixed order (deterministic).
// 1, 2 or 4 candidates cost the same in these phases.  Measured afterwards (clock trace,
// profiles/r01_decode_step.md): HMMA.16816 issues only every ~30-48 cycles per scheduler on B200,
// so the K = 4096 phase (32 MMAs per warp for 6-8 useful rows) stays on the CUDA cores for 1-2
// candidates; the other changes of this generation (deferred ring-slot release, one attention
// item per head over 128-key tiles, same-round re-polling, 32-bit ring counters) are described
// where they are implemented.
#pragma once
#include "ar_mega2.cuh"

namespace tts {

constexpr int M3_STAGES = 8;                       // ring depth
constexpr int M3_ROWS_K1 = 8, M3_ROWS_K4 = 2;      // weight rows per stage at K = 1024 / 4096 (16 KB of weights)
constexpr int M3_PITCH_K1 = kDim * 2 + 16;         // bytes between rows in shared memory (conflict-free ldmatrix)
constexpr int M3_PITCH_K4 = kFF * 2 + 16;
constexpr int M3_STAGE_SMEM = M3_ROWS_K1 * M3_PITCH_K1;  // 16512 (>= 2 x 8208)
constexpr int M3_GROUP = 4;                        // stages consumed together (<= 32 rows at K = 1024, 8 at K = 4096)
constexpr int M3_XP_K1 = kDim + 8, M3_XP_K4 = kFF + 8;   // halves between rows of the activation planes

template <int BT>
__host__ __device__ inline size_t mega3_smem_bytes() {
  return size_t(M3_STAGES) * M3_STAGE_SMEM + 256 /*mbarriers*/ + size_t(8) * 32 * BT * sizeof(float) /*partial tiles*/ +
         64 * sizeof(double) /*LN sums*/ + size_t(BT) * M3_XP_K4 * sizeof(__half) /*activation planes*/ +
         size_t(BT) * kDim * sizeof(float) /*residual*/ + size_t(2) * M2_KV_TILE * M2_KV_LD * sizeof(__half) /*K, V tile*/ +
         (128 + 64 + 8 * 64 + 32) * sizeof(float);
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// x = hi + lo in f16 (lo = f16(x - hi))
__device__ __forceinline__ void split16(float x0, float x1, __half2 &hi, __half2 &lo) {
  hi = __floats2half2_rn(x0, x1);
  const float2 h = __half22float2(hi);
  lo = __floats2half2_rn(x0 - h.x, x1 - h.y);
}

template <int BT>
static __global__ void __launch_bounds__(M2_THREADS, 1) ar_decode_mega3_kernel(Mega2Args a) {
  static_assert(BT == 1 || BT == 2 || BT == 4, "hi and lo planes of every candidate share one n = 8 MMA tile");
  constexpr int STAGES = M3_STAGES;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *ring = smem;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + STAGES * M3_STAGE_SMEM);
  uint64_t *empty = full + STAGES;
  float *partial = reinterpret_cast<float *>(smem + STAGES * M3_STAGE_SMEM + 256);  // [8 warps][32 rows][BT]
  double *red = reinterpret_cast<double *>(partial + 8 * 32 * BT);
  // activation planes (B operand): K = 1024 phases: rows 0..BT-1 = hi, BT..2BT-1 = lo, pitch M3_XP_K1;
  // K = 4096 phase: rows 0..BT-1 = hi only, pitch M3_XP_K4
  __half *xs = reinterpret_cast<__half *>(red + 64);
  float *hres = reinterpret_cast<float *>(xs + BT * M3_XP_K4);  // [BT][kDim] residual stream
  __half *kt = reinterpret_cast<__half *>(hres + BT * kDim);  // [128][72] K tile
  __half *vt = kt + M2_KV_TILE * M2_KV_LD;                    // [128][72] V tile
  float *sc = reinterpret_cast<float *>(vt + M2_KV_TILE * M2_KV_LD);  // [128] scores
  float *qs = sc + 128;                                               // [64] query
  float *pp = qs + 64;                                                // [8][64] partial outputs
  float *redf = pp + 8 * 64;                                          // [32] block reductions
  // the layer table (pointers) lives in shared memory: a pointer fetched from global memory in
  // front of every dependent load stalls the in-order issue for an L2 round trip
  __shared__ MegaLayer s_layers[kLayers];
  __shared__ int s_slice[4][2];  // {row0, rows} of this CTA for N = 3072, 1024, 4096, 8194

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x, cta = blockIdx.x;
  const int B = a.B;
  for (int i = tid; i < int(kLayers * sizeof(MegaLayer) / 8); i += M2_THREADS)
    reinterpret_cast<unsigned long long *>(s_layers)[i] = reinterpret_cast<const unsigned long long *>(a.layers)[i];
  if (tid < 4) {
    // the c_fc slices start and end on even rows (its outputs are exchanged as half2 pairs)
    const int N = tid == 0 ? 3072 : (tid == 1 ? kDim : (tid == 2 ? kFF : kMelVocab));
    const int gran = N == kFF ? 2 : 1;
    const int U = N / gran, base = U / G, rem = U % G;
    s_slice[tid][1] = gran * (base + (cta < rem ? 1 : 0));
    s_slice[tid][0] = gran * (cta * base + min(cta, rem));
  }

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], GV_WARPS);
    }
    fence_barrier_init();
  }
  __syncthreads();

  auto seg_shape = [&](int sid, int &N, int &K, const void *&W) {
    if (sid >= 120) { N = kMelVocab; K = kDim; W = a.lm_w; return; }
    const MegaLayer &l = s_layers[sid >> 2];
    switch (sid & 3) {
      case 0: N = 3072; K = kDim; W = l.w_qkv; break;
      case 1: N = kDim; K = kDim; W = l.w_proj; break;
      case 2: N = kFF; K = kDim; W = l.w_fc; break;
      default: N = kDim; K = kFF; W = l.w_proj2; break;
    }
  };
  // rows of a matrix owned by this CTA (table filled before the role split)
  auto slice = [&](int N, int &row0, int &rows) {
    const int i = N == 3072 ? 0 : (N == kDim ? 1 : (N == kFF ? 2 : 3));
    row0 = s_slice[i][0];
    rows = s_slice[i][1];
  };

  if (warp >= M2_CONSUMERS / 32) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(M2_PRODUCER_REGS));
    // ---------------- weight stream: every slice of the step, in consumption order ------------
    if (warp == M2_CONSUMERS / 32 && lane == 0) {
      // 0.77 GB of weights stream through a 126 MB L2 every step: evict-first keeps them from
      // displacing the (value, tag) exchange lines, whose misses show up as sporadic 2-3 us phases
      const uint64_t pol = l2_policy_evict_first();
      long it = 0;
      for (int sid = 0; sid <= 120; ++sid) {
        int N, K, row0, rows;
        const void *W;
        seg_shape(sid, N, K, W);
        slice(N, row0, rows);
        const uint32_t row_bytes = uint32_t(K) * 2u;
        const int rps = K == kDim ? M3_ROWS_K1 : M3_ROWS_K4, pitch = K == kDim ? M3_PITCH_K1 : M3_PITCH_K4;
        const unsigned char *src = reinterpret_cast<const unsigned char *>(W) + size_t(row0) * row_bytes;
        for (int r0 = 0; r0 < rows; r0 += rps, ++it) {
          const int slot = int(it % STAGES), nr = min(rps, rows - r0);
          mbar_wait(&empty[slot], uint32_t((it / STAGES) & 1) ^ 1u);
          mbar_arrive_expect_tx(&full[slot], uint32_t(nr) * row_bytes);
          for (int r = 0; r < nr; ++r) {
            void *dst = ring + size_t(slot) * M3_STAGE_SMEM + r * pitch;
            if (a.evict_first) bulk_g2s_hint(dst, src + size_t(r0 + r) * row_bytes, row_bytes, &full[slot], pol);
            else bulk_g2s(dst, src + size_t(r0 + r) * row_bytes, row_bytes, &full[slot]);
          }
        }
      }
    }
    return;
  }

  // =========================== consumers (256 threads) ========================================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(M2_CONSUMER_REGS));
  int dbg_n = 0;
  auto trace = [&](int tag) {
    if (a.dbg && a.dbg_mode == 1 && cta == 0 && tid == 0 && dbg_n < 4000) { a.dbg[2 * dbg_n] = tag; a.dbg[2 * dbg_n + 1] = clock64(); ++dbg_n; }
  };
  // per-CTA phase stamps (mode 2): [cta][phase][4] = phase start, GEMV start, GEMV done, epilogue done
  auto stamp = [&](int ph, int k) {
    if (a.dbg && a.dbg_mode == 2 && tid == 0) a.dbg[(size_t(cta) * 128 + ph) * 4 + k] = global_ns();
  };
  const uint32_t tag_base = a.epoch << 8;
  auto tag_of = [&](int layer, int phase) { return tag_base + uint32_t(layer * 8 + phase); };
  static_assert((M3_STAGES & (M3_STAGES - 1)) == 0, "slot = counter & (STAGES - 1)");
  uint32_t c_it = 0;    // ring stage counter (same order as the producer)
  uint32_t rel_it = 0;  // first stage of the group whose slots are still held
  int rel_n = 0;
  const int nb = min(a.Bmax, 4);  // candidates the exchange buffers are sized for
  const int nrep = a.nrep;       // replicas in use (<= M2_REP)
  const int rep = cta % nrep;   // the replica this CTA reads
  const size_t h_rep = size_t(nb) * kDim, m_rep = size_t(nb) * (kFF / 2), att_rep = size_t(nb) * kHeads * M2_SMAX * M2_REC;

  // poll one 16-byte unit (two pairs) until both tags match
  auto poll_unit = [&](const uint2 *p, uint32_t tag) -> float2 {
    uint4 v = ld_ll(p);
    while (v.y != tag || v.w != tag) {
      poll_backoff(a.poll_spin);
      v = ld_ll(p);
    }
    return make_float2(__uint_as_float(v.x), __uint_as_float(v.z));
  };

  // thread t owns elements {2t, 2t+1, 512+2t, 512+2t+1} of a 1024-vector
  float hv[BT][4];
  // Every pending unit is re-requested in the SAME round: polling unit after unit costs one extra
  // L2 round trip (~0.35 us) per unit after the previous one has arrived, because the first load
  // of every unit is issued before any producer has stored.
  auto poll_h = [&](const uint2 *buf, uint32_t tag) {
    uint4 v[BT][2];
#pragma unroll
    for (int b = 0; b < BT; ++b)
#pragma unroll
      for (int u = 0; u < 2; ++u) v[b][u] = make_uint4(0, ~tag, 0, ~tag);
    for (;;) {
      bool pending = false;
#pragma unroll
      for (int b = 0; b < BT; ++b)
#pragma unroll
        for (int u = 0; u < 2; ++u)
          if (b < B && (v[b][u].y != tag || v[b][u].w != tag)) {
            v[b][u] = ld_ll(buf + size_t(b) * kDim + u * 512 + 2 * tid);
            pending = true;
          }
      if (!pending) break;
    }
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      if (b < B) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          hv[b][2 * u] = __uint_as_float(v[b][u].x);
          hv[b][2 * u + 1] = __uint_as_float(v[b][u].z);
        }
      } else {
        hv[b][0] = hv[b][1] = hv[b][2] = hv[b][3] = 0.f;
      }
    }
  };

  // LayerNorm of hv over the CTA (every CTA normalises the full row itself): writes the raw row to
  // hres (if keep) and leaves LN(x) w + b in hv.  Statistics: single pass, double accumulation; the
  // mean is narrowed to float like ggml's (ggml.c:11935-11955).  lw/lb: this thread's 4 weights.
  // `slot` selects one of two scratch halves.
  auto layer_norm = [&](const float (&lw)[4], const float (&lb)[4], int slot, bool keep) {
    double *rb = red + slot * 32;
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      double s1 = (double(hv[b][0]) + double(hv[b][1])) + (double(hv[b][2]) + double(hv[b][3]));
      double s2 = (double(hv[b][0]) * hv[b][0] + double(hv[b][1]) * hv[b][1]) +
                  (double(hv[b][2]) * hv[b][2] + double(hv[b][3]) * hv[b][3]);
      s1 = warp_sum_d(s1);
      s2 = warp_sum_d(s2);
      if (lane == 0) { rb[(warp * BT + b) * 2] = s1; rb[(warp * BT + b) * 2 + 1] = s2; }
    }
    if (keep) {
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        *reinterpret_cast<float2 *>(hres + b * kDim + 2 * tid) = make_float2(hv[b][0], hv[b][1]);
        *reinterpret_cast<float2 *>(hres + b * kDim + 512 + 2 * tid) = make_float2(hv[b][2], hv[b][3]);
      }
    }
    bar_consumers();
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      double t1 = 0, t2 = 0;
#pragma unroll
      for (int w = 0; w < GV_WARPS; ++w) { t1 += rb[(w * BT + b) * 2]; t2 += rb[(w * BT + b) * 2 + 1]; }
      const double md = t1 * (1.0 / kDim);
      const float mean = float(md);
      double var = t2 * (1.0 / kDim) - 2.0 * md * double(mean) + double(mean) * double(mean);  // E[(x - mean_f)^2]
      if (var < 0) var = 0;
      const float rstd = 1.0f / sqrtf(float(var) + 1e-5f);
#pragma unroll
      for (int q = 0; q < 4; ++q) hv[b][q] = (hv[b][q] - mean) * rstd * lw[q] + lb[q];
    }
  };
  auto load_ln = [&](const float *w, const float *bb, float (&lw)[4], float (&lb)[4]) {
    const float2 w0 = *reinterpret_cast<const float2 *>(w + 2 * tid), w1 = *reinterpret_cast<const float2 *>(w + 512 + 2 * tid);
    const float2 b0 = *reinterpret_cast<const float2 *>(bb + 2 * tid), b1 = *reinterpret_cast<const float2 *>(bb + 512 + 2 * tid);
    lw[0] = w0.x; lw[1] = w0.y; lw[2] = w1.x; lw[3] = w1.y;
    lb[0] = b0.x; lb[1] = b0.y; lb[2] = b1.x; lb[3] = b1.y;
  };

  // ---------------- attention items -------------------------------------------------------------
  // One item per (candidate, head), all keys, M2_KV_TILE at a time with running (max, sum, acc).
  // Splitting the keys of a head over several CTAs costs ~4 us per layer and split level on
  // B200 (same-box A/B, profiles/r01_decode_keys_per_split_ab.txt: 520 -> 640 us per step when
  // 64 keys go from one item to two) -- far more than walking the tiles in one CTA.
  const int n_keys = a.n_past + 1;
  const int S = 1;  // records per (candidate, head) in the exchange buffer (the merge below is general)
  const int n_items = B * kHeads;
  const int n_tiles = (n_keys + M2_KV_TILE - 1) / M2_KV_TILE;
  const size_t layer_kv = size_t(a.Bmax) * kHeads * a.P * kHeadDim;
  // cached rows [t * TILE, min((t + 1) * TILE, n_past)) of item (b, head) -> shared tiles (cp.async, 16 B per op)
  auto prefetch_kv = [&](int li, int item, int t) {
    const int b = item / kHeads, head = item % kHeads;
    const int j0 = t * M2_KV_TILE;
    const int rows = min(j0 + M2_KV_TILE, a.n_past) - j0;
    const __half *K = a.kc + size_t(li) * layer_kv + ((size_t(a.b0 + b) * kHeads + head) * a.P + j0) * kHeadDim;
    const __half *V = a.vc + size_t(li) * layer_kv + ((size_t(a.b0 + b) * kHeads + head) * a.P + j0) * kHeadDim;
    for (int u = tid; u < rows * 8; u += M2_CONSUMERS) {
      const int r = u >> 3, c = u & 7;
      cp_async_cg16(kt + r * M2_KV_LD + c * 8, K + size_t(r) * kHeadDim + c * 8);
      cp_async_cg16(vt + r * M2_KV_LD + c * 8, V + size_t(r) * kHeadDim + c * 8);
    }
  };
  auto attention_item = [&](int li, int item, bool prefetched) {
    const int b = item / kHeads, head = item % kHeads;
    const uint32_t tq = tag_of(li, 1);
    const uint2 *qkv = a.ll_qkv + size_t(b) * 3072;
    float Mr = -INFINITY, Lr = 0.f, orun = 0.f;  // running max / sum (every thread), output dim tid (tid < 64)
    for (int t = 0; t < n_tiles; ++t) {
      const int j0 = t * M2_KV_TILE, j1 = min(n_keys, j0 + M2_KV_TILE), c = j1 - j0;
      const bool has_new = j1 == n_keys;
      if (t > 0 || !prefetched) prefetch_kv(li, item, t);
      if (tid < 32) {
        if (t == 0) {
          const float2 v = poll_unit(qkv + head * kHeadDim + 2 * tid, tq);
          qs[2 * tid] = v.x;
          qs[2 * tid + 1] = v.y;
        }
      } else if (tid < 64 && has_new) {
        const int u = tid - 32;
        const float2 v = poll_unit(qkv + 1024 + head * kHeadDim + 2 * u, tq);
        *reinterpret_cast<__half2 *>(kt + (c - 1) * M2_KV_LD + 2 * u) = __floats2half2_rn(v.x, v.y);
      } else if (tid < 96 && has_new) {
        const int u = tid - 64;
        const float2 v = poll_unit(qkv + 2048 + head * kHeadDim + 2 * u, tq);
        *reinterpret_cast<__half2 *>(vt + (c - 1) * M2_KV_LD + 2 * u) = __floats2half2_rn(v.x, v.y);
      }
      cp_async_wait_all();
      bar_consumers();
      // scores: two threads per key, 32 dims each
      float lmax = -INFINITY;
      {
        const int j = tid >> 1, half = tid & 1;
        float dot = 0.f;
        if (j < c) {
          const uint4 *kr = reinterpret_cast<const uint4 *>(kt + j * M2_KV_LD + half * 32);
          const float *qh = qs + half * 32;
          float d0 = 0.f, d1 = 0.f;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const uint4 u = kr[cc];
            const __half2 *h2 = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __half22float2(h2[e]);
              d0 = fmaf(qh[cc * 8 + 2 * e], f.x, d0);
              d1 = fmaf(qh[cc * 8 + 2 * e + 1], f.y, d1);
            }
          }
          dot = d0 + d1;
        }
        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
        dot *= 0.125f;
        if (j < c) {
          if (half == 0) sc[j] = dot;
          lmax = dot;
        }
      }
      lmax = warp_max(lmax);
      if (lane == 0) redf[warp] = lmax;
      bar_consumers();
      float mx = redf[0];
#pragma unroll
      for (int w = 1; w < GV_WARPS; ++w) mx = fmaxf(mx, redf[w]);
      float lsum = 0.f;
      if (tid < c) {
        const float p = expf(sc[tid] - mx);
        sc[tid] = p;
        lsum = p;
      }
      lsum = warp_sum(lsum);
      if (lane == 0) redf[8 + warp] = lsum;
      bar_consumers();
      // unnormalised output of the tile: lane = dim pair, warp = key partition
      {
        float o0 = 0.f, o1 = 0.f;
        for (int j = warp; j < c; j += GV_WARPS) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(vt + j * M2_KV_LD + 2 * lane));
          const float p = sc[j];
          o0 = fmaf(p, f.x, o0);
          o1 = fmaf(p, f.y, o1);
        }
        pp[warp * 64 + 2 * lane] = o0;
        pp[warp * 64 + 2 * lane + 1] = o1;
      }
      bar_consumers();
      // fold the tile into the running statistics (exp(-inf) = 0 on the first tile)
      {
        float ts = 0.f;
#pragma unroll
        for (int w = 0; w < GV_WARPS; ++w) ts += redf[8 + w];
        const float nM = fmaxf(Mr, mx);
        const float so = expf(Mr - nM), sn = expf(mx - nM);
        if (tid < 64) {
          float o = 0.f;
#pragma unroll
          for (int w = 0; w < GV_WARPS; ++w) o += pp[w * 64 + tid];
          orun = orun * so + o * sn;
        }
        Lr = Lr * so + ts * sn;
        Mr = nM;
      }
      // (the next tile / item starts with loads into kt/vt and writes sc/pp/redf: separate them from the reads above)
      bar_consumers();
    }
    uint2 *rec = a.ll_att + size_t(b * kHeads + head) * M2_SMAX * M2_REC;
    const uint32_t to = tag_of(li, 2);
    if (tid < 66) {
      const float o = tid < 64 ? orun : (tid == 64 ? Mr : Lr);
      for (int r = 0; r < nrep; ++r) st_ll(rec + r * att_rep + tid, o, to);
    }
  };
  // normalise the attention output of every (candidate, head): thread t -> head t/16, dims 4 (t%16) .. +3
  auto attention_merge = [&](int li) {
    const uint32_t tg = tag_of(li, 2);
    const int head = tid >> 4, d0 = (tid & 15) * 4;
    // one record per (candidate, head): {acc[64], max, sum}; this thread needs 4 acc values and the
    // sum (softmax-normalised output = acc / sum; the max only matters when records are merged).
    // All pending units of all candidates are re-requested in the same round (see poll_h).
    uint4 v[BT][3];
#pragma unroll
    for (int b = 0; b < BT; ++b)
#pragma unroll
      for (int u = 0; u < 3; ++u) v[b][u] = make_uint4(0, ~tg, 0, ~tg);
    for (;;) {
      bool pending = false;
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        const uint2 *rec = a.ll_att + rep * att_rep + size_t(b * kHeads + head) * M2_SMAX * M2_REC;
        const int off[3] = {d0, d0 + 2, 64};
#pragma unroll
        for (int u = 0; u < 3; ++u)
          if (b < B && (v[b][u].y != tg || v[b][u].w != tg)) {
            v[b][u] = ld_ll(rec + off[u]);
            pending = true;
          }
      }
      if (!pending) break;
    }
#pragma unroll
    for (int b = 0; b < BT; ++b) {
      if (b >= B) break;
      const float o[4] = {__uint_as_float(v[b][0].x), __uint_as_float(v[b][0].z), __uint_as_float(v[b][1].x),
                          __uint_as_float(v[b][1].z)};
      const float L = __uint_as_float(v[b][2].z);
      const float inv = 1.0f / L;
      __half2 h0, l0, h1, l1;
      split16(o[0] * inv, o[1] * inv, h0, l0);
      split16(o[2] * inv, o[3] * inv, h1, l1);
      const int k = head * kHeadDim + d0;
      *reinterpret_cast<uint2 *>(xs + b * M3_XP_K1 + k) = make_uint2(*reinterpret_cast<uint32_t *>(&h0), *reinterpret_cast<uint32_t *>(&h1));
      *reinterpret_cast<uint2 *>(xs + (BT + b) * M3_XP_K1 + k) = make_uint2(*reinterpret_cast<uint32_t *>(&l0), *reinterpret_cast<uint32_t *>(&l1));
    }
  };

  // =========================== the step: 30 x (QKV | c_proj | c_fc | mlp c_proj) + lm_head ==========
  // One loop, one copy of the GEMV body.  Phase p of layer li: prologue (what xin holds), then
  // the activation planes x this CTA's rows of the phase's matrix, then the epilogue that feeds the next phase.
  float lw[4], lb[4];
  constexpr int kPhases = kLayers * 4 + 1;
  for (int ph = 0; ph < kPhases; ++ph) {
    const bool tail = ph == kLayers * 4;
    const int li = tail ? kLayers - 1 : (ph >> 2), p = tail ? 0 : (ph & 3);
    const MegaLayer &l = s_layers[li];
    stamp(ph, 0);
    // ---------------- prologue ----------------
    if (p == 0 || p == 2) {
      if (p == 0 && !tail && cta < n_items) prefetch_kv(li, cta, 0);  // lands while the QKV phase runs
      load_ln(tail ? a.lnf_w : (p == 0 ? l.ln1_w : l.ln2_w), tail ? a.lnf_b : (p == 0 ? l.ln1_b : l.ln2_b), lw, lb);
      if (ph == 0) {  // h = mel_emb[tok] + mel_pos[pos]   (main.cpp:2676-2691)
#pragma unroll
        for (int b = 0; b < BT; ++b) {
          if (b < B) {
            const int tok = a.tokens[a.b0 + b];
            const float *e = a.mel_emb + size_t(tok) * kDim, *pe = a.mel_pos + size_t(a.pos_id) * kDim;
            const float2 e0 = *reinterpret_cast<const float2 *>(e + 2 * tid), e1 = *reinterpret_cast<const float2 *>(e + 512 + 2 * tid);
            const float2 p0 = *reinterpret_cast<const float2 *>(pe + 2 * tid), p1 = *reinterpret_cast<const float2 *>(pe + 512 + 2 * tid);
            hv[b][0] = e0.x + p0.x; hv[b][1] = e0.y + p0.y; hv[b][2] = e1.x + p1.x; hv[b][3] = e1.y + p1.y;
          } else {
            hv[b][0] = hv[b][1] = hv[b][2] = hv[b][3] = 0.f;
          }
        }
      } else if (p == 0) {
        poll_h(a.ll_h + rep * h_rep, tag_of(tail ? kLayers - 1 : li - 1, 5));
      } else {
        poll_h(a.ll_h2 + rep * h_rep, tag_of(li, 3));
      }
      trace(1 + p);
      const int passes = tail ? 2 : 1;  // tail: z = LN(LN(h) ln_f) lm_head.0   (main.cpp:2985-3003)
      for (int pass = 0; pass < passes; ++pass) {
        if (pass == 1) load_ln(a.lm0_w, a.lm0_b, lw, lb);
        layer_norm(lw, lb, pass ? 1 : (p >> 1), !tail);
      }
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        __half2 h0, l0, h1, l1;
        split16(hv[b][0], hv[b][1], h0, l0);
        split16(hv[b][2], hv[b][3], h1, l1);
        *reinterpret_cast<__half2 *>(xs + b * M3_XP_K1 + 2 * tid) = h0;
        *reinterpret_cast<__half2 *>(xs + b * M3_XP_K1 + 512 + 2 * tid) = h1;
        *reinterpret_cast<__half2 *>(xs + (BT + b) * M3_XP_K1 + 2 * tid) = l0;
        *reinterpret_cast<__half2 *>(xs + (BT + b) * M3_XP_K1 + 512 + 2 * tid) = l1;
      }
    } else if (p == 1) {
      trace(11);
      for (int item = cta; item < n_items; item += G) attention_item(li, item, item == cta);
      trace(12);
      attention_merge(li);
    } else {
      trace(15);
      // input of the second MLP matrix: 4096 f16-exact values per candidate as {half2, tag} pairs:
      // 1024 units of 4 values, 4 units per thread
      const uint32_t tg = tag_of(li, 4);
#pragma unroll
      for (int b = 0; b < BT; ++b) {
        if (b >= B) break;
        const uint2 *src = a.ll_m + rep * m_rep + size_t(b) * (kFF / 2);
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = make_uint4(0, ~tg, 0, ~tg);
        for (;;) {  // all pending units per round (see poll_h)
          bool pending = false;
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (v[u].y != tg || v[u].w != tg) {
              v[u] = ld_ll(src + 2 * (tid + u * M2_CONSUMERS));
              pending = true;
            }
          if (!pending) break;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          *reinterpret_cast<uint2 *>(xs + b * M3_XP_K4 + 4 * (tid + u * M2_CONSUMERS)) = make_uint2(v[u].x, v[u].z);
        }
      }
    }
    bar_consumers();
    trace(20 + p);
    stamp(ph, 1);

    // ---------------- GEMV on the tensor cores: D[rows x 8] = W[rows x K] . X[K x 8] ----------------
    // kind: 0 = QKV (f16 round trip, KV append), 1 = residual, 2 = GELU16, 3 = logits
    const int N = tail ? kMelVocab : (p == 0 ? 3072 : (p == 2 ? kFF : kDim));
    const bool k4 = p == 3;  // K = 4096
    const int kind = tail ? 3 : (p == 0 ? 0 : (p == 2 ? 2 : 1));
    const float *bias = tail ? a.lm_b : (p == 0 ? l.b_qkv : (p == 1 ? l.b_proj : (p == 2 ? l.b_fc : l.b_proj2)));
    const uint32_t out_tag = tag_of(li, p == 0 ? 1 : (p == 1 ? 3 : (p == 2 ? 4 : 5)));
    {
      const int rps = k4 ? M3_ROWS_K4 : M3_ROWS_K1, pitch = k4 ? M3_PITCH_K4 : M3_PITCH_K1, xp = k4 ? M3_XP_K4 : M3_XP_K1;
      const int ksteps = k4 ? 32 : 8;                    // 16-wide k steps of this warp's K / 8 slice
      const int kbase = warp * ksteps * 16;
      int row0, rows_cta;
      slice(N, row0, rows_cta);
      const int rlog = k4 ? 1 : 3;  // log2(rps): no runtime division in a phase
      const int n_stages = (rows_cta + rps - 1) >> rlog;
      // B fragment source: column n = lane / 4 of the n = 8 tile -> plane row (hi: n < 4, lo: n >= 4)
      const int ncol = lane >> 2;
      const int cand = ncol & 3;
      const bool col_on = cand < BT && !(k4 && ncol >= 4);
      const int xrow = (ncol < 4 ? 0 : BT) + cand;
      const uint32_t xaddr = smem_u32(xs) + uint32_t((col_on ? xrow : 0) * xp + kbase + (lane & 3) * 2) * 2u;
      // A fragment source: ldmatrix lane -> (row within the 16-row tile, k half)
      const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_kofs = (lane >> 4) * 8;
      for (int s0 = 0; s0 < n_stages; s0 += M3_GROUP) {
        const int gs = min(M3_GROUP, n_stages - s0);          // stages of this group
        const int grow0 = s0 * rps, grows = min(gs * rps, rows_cta - grow0);  // rows of this group
        float ep_bias = 0.f;
        if (tid < grows * BT) ep_bias = bias[row0 + grow0 + tid / BT];
        // Deferred release: the slots of the previous group are handed back to the producer only
        // now, i.e. AFTER the exchange that followed it.  Released at the end of a GEMV, the
        // refill burst (148 CTAs x 3-4 stages = 8 MB of TMA traffic) coincides exactly with the
        // latency-critical (value, tag) exchange and roughly doubles it (tools/xchg_bench.cu:
        // 3757 -> 5873 cycles per exchange under a saturating stream); released here it overlaps
        // this group's arithmetic, which reads shared memory only.  (<= 4 pending + <= 4 current
        // stages never exceed the ring of 8.)
        __syncwarp();
        if (lane == 0)
          for (int j = 0; j < rel_n; ++j) mbar_arrive(&empty[(rel_it + j) & (STAGES - 1)]);
        rel_n = 0;
        // the (up to 4) stage barriers are probed together: four blocking try_wait in a row cost
        // ~90 cycles each even when the stages landed long ago
        for (;;) {
          bool ok = true;
#pragma unroll
          for (int j = 0; j < M3_GROUP; ++j)
            if (j < gs) ok = mbar_test_wait(&full[(c_it + j) & (STAGES - 1)], ((c_it + j) / STAGES) & 1u) && ok;
          if (ok) break;
        }
        trace(50);
        if (BT <= 2 && k4) {
          // K = 4096 phase (6-8 rows per CTA) on the CUDA cores: with so few rows an m16n8k16 tile
          // is 5 % useful and HMMA.16816 issues only every ~40 cycles per scheduler on B200 (32 per
          // warp = ~3100 cycles, profiles/r01_decode_step.md); 4 rows x 32 weights per thread are
          // ~260 instructions.  Warp w: K quarter w & 3, row parity w >> 2 of each 2-row stage.
          // The activations (f16-exact, hi plane only) are widened once per phase.
          const int kq = warp & 3, rsub = warp >> 2;
          float xr[BT][4][8];
#pragma unroll
          for (int b = 0; b < BT; ++b)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 u = *reinterpret_cast<const uint4 *>(xs + b * M3_XP_K4 + kq * 1024 + q * 256 + lane * 8);
              const __half2 *h2 = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(h2[e]);
                xr[b][q][2 * e] = f.x;
                xr[b][q][2 * e + 1] = f.y;
              }
            }
          uint4 wv[M3_GROUP][4];
#pragma unroll
          for (int j = 0; j < M3_GROUP; ++j) {
            const int r = j * 2 + rsub;
            const bool valid = j < gs && r < grows;
            const unsigned char *wp = ring + ((c_it + j) & (STAGES - 1)) * M3_STAGE_SMEM + rsub * M3_PITCH_K4 + kq * 2048 + lane * 16;
#pragma unroll
            for (int q = 0; q < 4; ++q) wv[j][q] = valid ? *reinterpret_cast<const uint4 *>(wp + q * 512) : make_uint4(0, 0, 0, 0);
          }
          float accf[M3_GROUP][BT];
#pragma unroll
          for (int j = 0; j < M3_GROUP; ++j) {
            float aq[4][BT];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float wf[8];
              const __half2 *h2 = reinterpret_cast<const __half2 *>(&wv[j][q]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(h2[e]);
                wf[2 * e] = f.x;
                wf[2 * e + 1] = f.y;
              }
#pragma unroll
              for (int b = 0; b < BT; ++b) {
                float t = wf[0] * xr[b][q][0];
#pragma unroll
                for (int e = 1; e < 8; ++e) t = fmaf(wf[e], xr[b][q][e], t);
                aq[q][b] = t;
              }
            }
#pragma unroll
            for (int b = 0; b < BT; ++b) accf[j][b] = (aq[0][b] + aq[2][b]) + (aq[1][b] + aq[3][b]);
          }
          trace(45);
          rel_it = c_it;
          rel_n = gs;
          c_it += gs;
          if (!a.defer) {
            __syncwarp();
            if (lane == 0)
              for (int j = 0; j < rel_n; ++j) mbar_arrive(&empty[(rel_it + j) & (STAGES - 1)]);
            rel_n = 0;
          }
#pragma unroll
          for (int j = 0; j < M3_GROUP; ++j)
#pragma unroll
            for (int b = 0; b < BT; ++b) {
              const float v = warp_sum(accf[j][b]);
              // rows of the other parity get a zero from this warp (the epilogue sums all 8 warps)
              if (lane == 0) {
                partial[(warp * 32 + j * 2 + rsub) * BT + b] = v;
                partial[(warp * 32 + j * 2 + (rsub ^ 1)) * BT + b] = 0.f;
              }
            }
        } else {
        const int mtiles = (grows + 15) >> 4;  // 1 or 2
        uint32_t aaddr[2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          int gr = mt * 16 + a_row;
          if (gr >= grows) gr = 0;  // rows past the slice: any resident row (results unused)
          const int slot = int((c_it + uint32_t(gr >> rlog)) & (STAGES - 1));
          aaddr[mt] = smem_u32(ring) + uint32_t(slot * M3_STAGE_SMEM + (gr & (rps - 1)) * pitch + (kbase + a_kofs) * 2);
        }
        // Batches of four independent (ldmatrix, mma) pairs: every load of a batch is issued before
        // its first mma, and the four mma of a batch feed four different accumulators, so the
        // chain per batch is one shared-memory latency + one mma latency.
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;
        auto ldb = [&](int ks, uint32_t &b0, uint32_t &b1) {
          b0 = 0; b1 = 0;
          if (col_on) {
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(b0) : "r"(xaddr + ks * 32));
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(b1) : "r"(xaddr + ks * 32 + 16));
          }
        };
        if (mtiles == 2) {
          // acc[0], acc[1]: row tile 0 (even / odd k step); acc[2], acc[3]: row tile 1
          for (int ks = 0; ks < ksteps; ks += 2) {
            uint32_t b[2][2], fa[4][4];
            ldb(ks, b[0][0], b[0][1]);
            ldb(ks + 1, b[1][0], b[1][1]);
            ldmatrix_x4(aaddr[0] + ks * 32, fa[0][0], fa[0][1], fa[0][2], fa[0][3]);
            ldmatrix_x4(aaddr[0] + ks * 32 + 32, fa[1][0], fa[1][1], fa[1][2], fa[1][3]);
            ldmatrix_x4(aaddr[1] + ks * 32, fa[2][0], fa[2][1], fa[2][2], fa[2][3]);
            ldmatrix_x4(aaddr[1] + ks * 32 + 32, fa[3][0], fa[3][1], fa[3][2], fa[3][3]);
            mma_16816(acc[0], fa[0][0], fa[0][1], fa[0][2], fa[0][3], b[0][0], b[0][1]);
            mma_16816(acc[1], fa[1][0], fa[1][1], fa[1][2], fa[1][3], b[1][0], b[1][1]);
            mma_16816(acc[2], fa[2][0], fa[2][1], fa[2][2], fa[2][3], b[0][0], b[0][1]);
            mma_16816(acc[3], fa[3][0], fa[3][1], fa[3][2], fa[3][3], b[1][0], b[1][1]);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) { acc[0][q] += acc[1][q]; acc[1][q] = acc[2][q] + acc[3][q]; }
        } else {
          for (int ks = 0; ks < ksteps; ks += 4) {
            uint32_t b[4][2], fa[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) ldb(ks + i, b[i][0], b[i][1]);
#pragma unroll
            for (int i = 0; i < 4; ++i) ldmatrix_x4(aaddr[0] + (ks + i) * 32, fa[i][0], fa[i][1], fa[i][2], fa[i][3]);
#pragma unroll
            for (int i = 0; i < 4; ++i) mma_16816(acc[i], fa[i][0], fa[i][1], fa[i][2], fa[i][3], b[i][0], b[i][1]);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[0][q] = (acc[0][q] + acc[1][q]) + (acc[2][q] + acc[3][q]);
        }
        trace(45);
        rel_it = c_it;  // released when the NEXT group starts (see below)
        rel_n = gs;
        c_it += gs;
        if (!a.defer) {  // A/B knob (TTS_MEGA_NODEFER=1): hand the slots back at once
          __syncwarp();
          if (lane == 0)
            for (int j = 0; j < rel_n; ++j) mbar_arrive(&empty[(rel_it + j) & (STAGES - 1)]);
          rel_n = 0;
        }
        // hi + lo planes: columns c and c + 4 sit two lanes apart
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          if (mt < mtiles) {
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[mt][q] += __shfl_xor_sync(0xffffffffu, acc[mt][q], 2);
            if ((lane & 3) < 2) {
              const int r = mt * 16 + (lane >> 2), c = (lane & 3) * 2;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int cc = c + (q & 1), rr = r + (q >> 1) * 8;
                if (cc < BT) partial[(warp * 32 + rr) * BT + cc] = acc[mt][q];
              }
            }
          }
        }
        }  // tensor-core path
        trace(46);
        bar_consumers();
        if (s0 == 0) { trace(30 + p); stamp(ph, 2); }
        // ---------------- epilogue: one output element per thread ----------------
        const bool act = tid < grows * BT && (tid % BT) < B;
        const int r = tid / BT, b = tid % BT, n = row0 + grow0 + r;
        float v = 0.f;
        if (act) {
#pragma unroll
          for (int w = 0; w < GV_WARPS; ++w) v += partial[(w * 32 + r) * BT + b];
          v += ep_bias;
          if (kind == 2) v = gelu16(v);
        }
        // (c_fc: rows n, n+1 of one candidate sit BT lanes apart; its slices start on even rows)
        const float v_next = __shfl_down_sync(0xffffffffu, v, BT);
        if (act) {
          if (kind == 0) {
            const __half hv16 = __float2half_rn(v);
            st_ll(a.ll_qkv + size_t(b) * 3072 + n, __half2float(hv16), out_tag);
            const int which = n >> 10, c = n & 1023;
            if (which != 0) {
              __half *cache = (which == 1 ? a.kc : a.vc) + size_t(li) * layer_kv;
              cache[(size_t(a.b0 + b) * kHeads + (c >> 6)) * size_t(a.P) * kHeadDim + size_t(a.n_past) * kHeadDim + (c & 63)] = hv16;
            }
          } else if (kind == 1) {
            const float o = hres[b * kDim + n] + v;
            uint2 *dst = (p == 1 ? a.ll_h2 : a.ll_h) + size_t(b) * kDim + n;
            for (int rr = 0; rr < nrep; ++rr) st_ll(dst + rr * h_rep, o, out_tag);
          } else if (kind == 2) {
            if ((r & 1) == 0) {
              const __half2 h2 = __floats2half2_rn(v, v_next);  // exact: gelu16 outputs are f16 values
              uint2 *dst = a.ll_m + size_t(b) * (kFF / 2) + (n >> 1);
              for (int rr = 0; rr < nrep; ++rr) st_ll_u32(dst + rr * m_rep, *reinterpret_cast<const uint32_t *>(&h2), out_tag);
            }
          } else {
            a.logits[size_t(a.b0 + b) * N + n] = v;
          }
        }
        if (s0 + M3_GROUP < n_stages) bar_consumers();  // the next group overwrites the partial tiles
      }
    }
    stamp(ph, 3);
  }
  trace(40);
}

}  // namespace tts
