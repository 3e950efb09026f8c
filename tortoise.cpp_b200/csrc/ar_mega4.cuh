// ar_mega4.cuh -- AR decode step as one persistent kernel for 5..16 candidates on ONE weight stream
// (f16 weights; BASELINE configs[2]: 16 candidates per GPU, configs[3]: 8 per GPU), and for 1..16 DIFFERENT
// utterances in the candidate slots (configs[4], tts_ar_prefill_multi: prompts right-aligned in the KV cache,
// Mega4Args::start[b] = first live key of slot b; nothing else in the kernel knows the difference).
//
// Same math (reference graph autoregressive_graph(fake_inputs=false), main.cpp:2668-3029; candidate
// batch axis main.cpp:5044) and the same cross-CTA protocol as ar_mega2/3.cuh: (value, tag) pairs
// through L2, no grid barrier, one producer warp streaming every weight row of the step through a
// shared-memory ring with TMA bulk copies.  What is different, and why:
//   * the MMA operands are SWAPPED.  ar_mega3 computes D[16 weight rows x 8] with the 8 columns =
//     (hi, lo) planes of <= 4 candidates; here the candidates are the M dimension:
//     D[16 x 8 weight rows] = X[16 x k16] . W^T[k16 x 8].  For <= 8 candidates rows 0-7 of the
//     A tile are the hi planes and rows 8-15 the lo planes (ONE mma.sync per 8 weight rows and k
//     step, result = D[c] + D[c + 8]); for 16 candidates a hi tile and a lo tile accumulate into
//     the same D.  Weight rows are consumed 8 at a time (no 16-row padding);
//   * LayerNorm is done by ONE WARP PER CANDIDATE (lane = 32 elements, shuffle-only statistics in
//     double like ggml.c:11935-11955): no cross-warp reduction, no candidate count in the register
//     budget; every CTA still normalises all candidates itself;
//   * polls are issued in rounds of at most 16 units (64 registers) per thread; only the first
//     round of a phase actually waits;
//   * the residual stream is kept only for the rows this CTA owns;
//   * the K = 4096 phase takes its 8 KB rows two per stage as before, but the 4096-wide f16 hidden
//     vectors of 16 candidates (128 KB) do not fit beside the ring: candidates are fed to the same
//     resident weight stages 8 at a time;
//   * attention keeps ar_mega3's one-CTA-per-(candidate, head) items (two rounds per layer at 16
//     candidates); a one-WARP-per-item variant (K / V rows straight from L2, no block barrier) measured
//     SLOWER on B200: 925 -> 1337 us per step at 8 candidates, 1470 -> 1840 at 16
//     (profiles/r02c_decode_warp_items_ab.txt) -- a single warp's dependent load chain is the longer pole;
//   * the prompt's K/V rows (identical for every candidate: the reference tiles B identical rows,
//     main.cpp:2640-2652) are stored ONCE, in candidate slot 0: attention item (b, head) reads rows
//     [0, n_prefix) from slot 0 and its own rows after that.
#pragma once
#include "ar_mega3.cuh"

namespace tts {

constexpr int M4_XP1 = kDim + 8, M4_XP4 = kFF + 8;   // halves between rows of the activation planes
constexpr int M4_XS_BYTES = 8 * M4_XP4 * 2;          // 65664: 8 hi rows at K = 4096 / 32 (hi + lo) rows at K = 1024 (66048)
constexpr int M4_XS_BYTES_MAX = 32 * M4_XP1 * 2 > M4_XS_BYTES ? 32 * M4_XP1 * 2 : M4_XS_BYTES;
constexpr int M4_OWN_ROWS = 8;                       // rows of an N = 1024 matrix one CTA owns (<= 7 at 148 CTAs)
constexpr int M4_REC = 68;                           // pairs per attention record: 64 acc, max, sum, 2 pad

template <int BT>
struct M4Cfg {
  static constexpr int kStages = BT == 16 ? 7 : 8;  // 16 KB ring stages that fit beside the planes
};

template <int BT>
__host__ __device__ inline size_t mega4_smem_bytes() {
  return size_t(M4Cfg<BT>::kStages) * M3_STAGE_SMEM + 256 /*mbarriers*/ + size_t(8) * 32 * BT * sizeof(float) /*partial tiles*/ +
         size_t(BT == 16 ? M4_XS_BYTES_MAX : M4_XS_BYTES) /*activation planes*/ + size_t(2) * 2 * kDim * sizeof(float) /*LN weights*/ +
         size_t(M4_OWN_ROWS) * BT * sizeof(float) /*residual rows*/ + (128 + 64 + 8 * 64 + 32) * sizeof(float);
  // (the K / V tiles of the attention items alias the activation planes, see below)
}

struct Mega4Args {
  const MegaLayer *layers;  // [30], device
  const float *lnf_w, *lnf_b, *lm0_w, *lm0_b, *lm_b;
  const void *lm_w;
  const float *mel_emb, *mel_pos;
  const int *tokens;
  // (value, tag) exchange buffers sized for 16 candidates: h / h2 / m / att hold nrep replicas
  uint2 *ll_h, *ll_h2, *ll_qkv, *ll_m, *ll_att;
  float *logits;
  __half *kc, *vc;  // [30][Bmax][16][P][64]
  int B, Bmax, P, n_past, pos_id;
  int n_prefix;     // K/V rows [0, n_prefix) of every candidate live in candidate slot 0
  // Utterance batching (slots = DIFFERENT prompts, n_prefix = 0): the prompts are right-aligned in the cache, slot b's
  // rows [0, start[b]) are padding that attention item (b, head) never looks at.  All zero for candidates of one prompt.
  int start[16];
  unsigned int epoch;
  int nrep;
};

__device__ __forceinline__ void ldmatrix_x2(uint32_t addr, uint32_t &r0, uint32_t &r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}

template <int BT>
static __global__ void __launch_bounds__(M2_THREADS, 1) ar_decode_mega4_kernel(Mega4Args a) {
  static_assert(BT == 8 || BT == 16, "5..8 candidates: stacked hi/lo tile; 9..16: a hi tile and a lo tile");
  constexpr int STAGES = M4Cfg<BT>::kStages;
  constexpr int NT = BT == 16 ? 2 : 1;  // A tiles of a K = 1024 phase
  constexpr int CPW = BT / 8;           // candidates per warp in the LayerNorm / embedding prologue
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *ring = smem;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + STAGES * M3_STAGE_SMEM);
  uint64_t *empty = full + STAGES;
  float *partial = reinterpret_cast<float *>(smem + STAGES * M3_STAGE_SMEM + 256);  // [8 warps][32 rows][BT]
  __half *xs = reinterpret_cast<__half *>(partial + 8 * 32 * BT);                    // activation planes
  float *lnw = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(xs) + (BT == 16 ? M4_XS_BYTES_MAX : M4_XS_BYTES));  // [2][2][1024]: {w, b} x 2 sets
  float *hown = lnw + 2 * 2 * kDim;                                                  // [M4_OWN_ROWS][BT] residual rows of this CTA
  // K / V tiles of the attention items ALIAS the activation planes: the LN1 planes are dead once the
  // QKV phase's MMAs are done, and attention_merge rewrites the planes only after the last item
  __half *kt = xs;                                                                   // [128][72] K tile
  __half *vt = kt + M2_KV_TILE * M2_KV_LD;                                           // [128][72] V tile
  float *sc = hown + M4_OWN_ROWS * BT;                                               // [128] scores
  float *qs = sc + 128;                                                              // [64] query
  float *pp = qs + 64;                                                               // [8][64] partial outputs
  float *redf = pp + 8 * 64;                                                         // [32] block reductions
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
  auto slice = [&](int N, int &row0, int &rows) {
    const int i = N == 3072 ? 0 : (N == kDim ? 1 : (N == kFF ? 2 : 3));
    row0 = s_slice[i][0];
    rows = s_slice[i][1];
  };

  if (warp >= M2_CONSUMERS / 32) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(M2_PRODUCER_REGS));
    // ---------------- weight stream: every slice of the step, in consumption order ------------
    if (warp == M2_CONSUMERS / 32 && lane == 0) {
      uint32_t it = 0;
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
          mbar_wait(&empty[slot], ((it / STAGES) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&full[slot], uint32_t(nr) * row_bytes);
          for (int r = 0; r < nr; ++r)
            bulk_g2s(ring + size_t(slot) * M3_STAGE_SMEM + r * pitch, src + size_t(r0 + r) * row_bytes, row_bytes, &full[slot]);
        }
      }
    }
    return;
  }

  // =========================== consumers (256 threads) ========================================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(M2_CONSUMER_REGS));
  const uint32_t tag_base = a.epoch << 8;
  auto tag_of = [&](int layer, int phase) { return tag_base + uint32_t(layer * 8 + phase); };
  uint32_t c_it = 0;    // ring stage counter (same order as the producer)
  uint32_t rel_it = 0;  // first stage of the group whose slots are still held
  int rel_n = 0;
  const int nrep = a.nrep;
  const int rep = cta % nrep;
  constexpr size_t NB = 16;  // candidates the exchange buffers are laid out for
  const size_t h_rep = NB * kDim, m_rep = NB * (kFF / 2), att_rep = NB * kHeads * M4_REC;

  auto poll_unit = [&](const uint2 *p, uint32_t tag) -> float2 {
    uint4 v = ld_ll(p);
    while (v.y != tag || v.w != tag) v = ld_ll(p);
    return make_float2(__uint_as_float(v.x), __uint_as_float(v.z));
  };

  int own_row0, own_rows;  // this CTA's rows of the N = 1024 matrices (c_proj, mlp c_proj): residual rows
  slice(kDim, own_row0, own_rows);

  // LayerNorm weights of the coming phase -> shared memory (requested before the poll)
  auto stage_ln = [&](int set, const float *w, const float *bb) {
    float *dw = lnw + set * 2 * kDim;
    cp_async_cg16(dw + tid * 4, w + tid * 4);
    cp_async_cg16(dw + kDim + tid * 4, bb + tid * 4);
  };

  // ---------------- row prologue: one warp per candidate ------------------------------------------
  // lane l holds elements {2 l + 64 j, 2 l + 64 j + 1 : j = 0..15} of its candidate's 1024-vector.
  // x -> (optionally) residual rows of this CTA -> LN (1 or 2 passes) -> hi / lo planes in xs.
  auto ln_rows_to_planes = [&](float (&x)[32], int b, int passes, bool keep) {
    if (keep) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int e = 2 * lane + 64 * j;
        if (e >= own_row0 && e < own_row0 + own_rows) hown[(e - own_row0) * BT + b] = x[2 * j];
        if (e + 1 >= own_row0 && e + 1 < own_row0 + own_rows) hown[(e + 1 - own_row0) * BT + b] = x[2 * j + 1];
      }
    }
    for (int pass = 0; pass < passes; ++pass) {
      // single pass, double accumulation; the mean is narrowed to float like ggml's (ggml.c:11935-11955)
      double s1 = 0.0, s2 = 0.0;
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        s1 += double(x[i]) + double(x[i + 1]);
        s2 += double(x[i]) * x[i] + double(x[i + 1]) * x[i + 1];
      }
      s1 = warp_sum_d(s1);
      s2 = warp_sum_d(s2);
      const double md = s1 * (1.0 / kDim);
      const float mean = float(md);
      double var = s2 * (1.0 / kDim) - 2.0 * md * double(mean) + double(mean) * double(mean);  // E[(x - mean_f)^2]
      if (var < 0) var = 0;
      const float rstd = 1.0f / sqrtf(float(var) + 1e-5f);
      const float *w = lnw + pass * 2 * kDim, *bb = w + kDim;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float2 w2 = *reinterpret_cast<const float2 *>(w + 2 * lane + 64 * j);
        const float2 b2 = *reinterpret_cast<const float2 *>(bb + 2 * lane + 64 * j);
        x[2 * j] = (x[2 * j] - mean) * rstd * w2.x + b2.x;
        x[2 * j + 1] = (x[2 * j + 1] - mean) * rstd * w2.y + b2.y;
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      __half2 hi, lo;
      split16(x[2 * j], x[2 * j + 1], hi, lo);
      *reinterpret_cast<__half2 *>(xs + b * M4_XP1 + 2 * lane + 64 * j) = hi;
      *reinterpret_cast<__half2 *>(xs + (BT + b) * M4_XP1 + 2 * lane + 64 * j) = lo;
    }
  };
  auto zero_planes = [&](int b) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      *reinterpret_cast<uint32_t *>(xs + b * M4_XP1 + 2 * lane + 64 * j) = 0u;
      *reinterpret_cast<uint32_t *>(xs + (BT + b) * M4_XP1 + 2 * lane + 64 * j) = 0u;
    }
  };

  // ---------------- attention items (as ar_mega3.cuh, plus the shared prompt rows) -----------------
  const int n_keys = a.n_past + 1;
  const int n_items = B * kHeads;
  const int n_tiles = (n_keys + M2_KV_TILE - 1) / M2_KV_TILE;
  const size_t layer_kv = size_t(a.Bmax) * kHeads * a.P * kHeadDim;
  auto prefetch_kv = [&](int li, int item, int t) {
    const int b = item / kHeads, head = item % kHeads;
    const int j0 = t * M2_KV_TILE;
    const int rows = min(j0 + M2_KV_TILE, a.n_past) - j0;
    const size_t lbase = size_t(li) * layer_kv;
    for (int u = tid; u < rows * 8; u += M2_CONSUMERS) {
      const int r = u >> 3, c = u & 7;
      const int slot = (j0 + r) < a.n_prefix ? 0 : b;  // prompt rows are stored once
      const size_t off = lbase + ((size_t(slot) * kHeads + head) * a.P + j0 + r) * kHeadDim + c * 8;
      cp_async_cg16(kt + r * M2_KV_LD + c * 8, a.kc + off);
      cp_async_cg16(vt + r * M2_KV_LD + c * 8, a.vc + off);
    }
  };
  auto first_tile = [&](int item) { return a.start[item / kHeads] / M2_KV_TILE; };
  auto attention_item = [&](int li, int item, bool prefetched) {
    const int b = item / kHeads, head = item % kHeads;
    const uint32_t tq = tag_of(li, 1);
    const uint2 *qkv = a.ll_qkv + size_t(b) * 3072;
    float Mr = -INFINITY, Lr = 0.f, orun = 0.f;
    const int start = a.start[b], t0 = start / M2_KV_TILE;  // (the tile of `start` holds a live key: start < n_keys)
    for (int t = t0; t < n_tiles; ++t) {
      const int j0 = t * M2_KV_TILE, j1 = min(n_keys, j0 + M2_KV_TILE), c = j1 - j0;
      const int cs = max(start - j0, 0);  // keys [0, cs) of this tile are padding
      const bool has_new = j1 == n_keys;
      if (t > t0 || !prefetched) prefetch_kv(li, item, t);
      if (tid < 32) {
        if (t == t0) {
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
          if (j < cs) dot = -INFINITY;  // padding: weight exp(-inf) = 0
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
      {
        float o0 = 0.f, o1 = 0.f;
        for (int j = cs + warp; j < c; j += GV_WARPS) {  // (padding rows may hold anything, NaN included: never read)
          const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(vt + j * M2_KV_LD + 2 * lane));
          const float p = sc[j];
          o0 = fmaf(p, f.x, o0);
          o1 = fmaf(p, f.y, o1);
        }
        pp[warp * 64 + 2 * lane] = o0;
        pp[warp * 64 + 2 * lane + 1] = o1;
      }
      bar_consumers();
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
      bar_consumers();
    }
    uint2 *rec = a.ll_att + size_t(b * kHeads + head) * M4_REC;
    const uint32_t to = tag_of(li, 2);
    if (tid < 66) {
      const float o = tid < 64 ? orun : (tid == 64 ? Mr : Lr);
      for (int r = 0; r < nrep; ++r) st_ll(rec + r * att_rep + tid, o, to);
    }
  };
  // normalised attention output of every (candidate, head) -> planes: thread t -> head t / 16, dims 4 (t % 16) .. + 3;
  // four candidates per round
  auto attention_merge = [&](int li) {
    const uint32_t tg = tag_of(li, 2);
    const int head = tid >> 4, d0 = (tid & 15) * 4;
    for (int b0 = 0; b0 < BT; b0 += 4) {
      uint4 v[4][3];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int u = 0; u < 3; ++u) v[i][u] = make_uint4(0, ~tg, 0, ~tg);
      for (;;) {
        bool pending = false;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint2 *rec = a.ll_att + rep * att_rep + size_t((b0 + i) * kHeads + head) * M4_REC;
          const int off[3] = {d0, d0 + 2, 64};
#pragma unroll
          for (int u = 0; u < 3; ++u)
            if (b0 + i < B && (v[i][u].y != tg || v[i][u].w != tg)) {
              v[i][u] = ld_ll(rec + off[u]);
              pending = true;
            }
        }
        if (!pending) break;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int b = b0 + i;
        const int k = head * kHeadDim + d0;
        uint2 hi2 = make_uint2(0u, 0u), lo2 = make_uint2(0u, 0u);
        if (b < B) {
          const float inv = 1.0f / __uint_as_float(v[i][2].z);
          __half2 h0, l0, h1, l1;
          split16(__uint_as_float(v[i][0].x) * inv, __uint_as_float(v[i][0].z) * inv, h0, l0);
          split16(__uint_as_float(v[i][1].x) * inv, __uint_as_float(v[i][1].z) * inv, h1, l1);
          hi2 = make_uint2(*reinterpret_cast<uint32_t *>(&h0), *reinterpret_cast<uint32_t *>(&h1));
          lo2 = make_uint2(*reinterpret_cast<uint32_t *>(&l0), *reinterpret_cast<uint32_t *>(&l1));
        }
        *reinterpret_cast<uint2 *>(xs + b * M4_XP1 + k) = hi2;
        *reinterpret_cast<uint2 *>(xs + (BT + b) * M4_XP1 + k) = lo2;
      }
    }
  };

  // =========================== the step: 30 x (QKV | c_proj | c_fc | mlp c_proj) + lm_head ==========
  constexpr int kPhases = kLayers * 4 + 1;
  for (int ph = 0; ph < kPhases; ++ph) {
    const bool tail = ph == kLayers * 4;
    const int li = tail ? kLayers - 1 : (ph >> 2), p = tail ? 0 : (ph & 3);
    const MegaLayer &l = s_layers[li];
    const int N = tail ? kMelVocab : (p == 0 ? 3072 : (p == 2 ? kFF : kDim));
    const bool k4 = p == 3;  // K = 4096
    const int kind = tail ? 3 : (p == 0 ? 0 : (p == 2 ? 2 : 1));  // 0 QKV, 1 residual, 2 GELU16, 3 logits
    const float *bias = tail ? a.lm_b : (p == 0 ? l.b_qkv : (p == 1 ? l.b_proj : (p == 2 ? l.b_fc : l.b_proj2)));
    const uint32_t out_tag = tag_of(li, p == 0 ? 1 : (p == 1 ? 3 : (p == 2 ? 4 : 5)));
    // K = 4096 with 16 candidates: two passes of 8 candidates over the same resident weight stages
    const int n_pass = (k4 && BT == 16) ? 2 : 1;

    // ---------------- prologue of the K = 1024 phases ----------------
    if (p == 0 || p == 2) {
      stage_ln(0, tail ? a.lnf_w : (p == 0 ? l.ln1_w : l.ln2_w), tail ? a.lnf_b : (p == 0 ? l.ln1_b : l.ln2_b));
      if (tail) stage_ln(1, a.lm0_w, a.lm0_b);
      asm volatile("cp.async.commit_group;" ::: "memory");
      const uint2 *src = ph == 0 ? nullptr : (p == 0 ? a.ll_h : a.ll_h2) + rep * h_rep;
      const uint32_t tg = p == 0 ? tag_of(tail ? kLayers - 1 : li - 1, 5) : tag_of(li, 3);
      bool ln_ready = false;
#pragma unroll 1
      for (int cw = 0; cw < CPW; ++cw) {
        const int b = warp + 8 * cw;
        float x[32];
        if (b < B) {
          if (ph == 0) {  // h = mel_emb[tok] + mel_pos[pos]   (main.cpp:2676-2691)
            const int tok = a.tokens[b];
            const float *e = a.mel_emb + size_t(tok) * kDim, *pe = a.mel_pos + size_t(a.pos_id) * kDim;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 e2 = *reinterpret_cast<const float2 *>(e + 2 * lane + 64 * j);
              const float2 p2 = *reinterpret_cast<const float2 *>(pe + 2 * lane + 64 * j);
              x[2 * j] = e2.x + p2.x;
              x[2 * j + 1] = e2.y + p2.y;
            }
          } else {
            // all 16 units of the candidate are re-requested in the same round until every tag matches
            uint4 v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = make_uint4(0, ~tg, 0, ~tg);
            const uint2 *row = src + size_t(b) * kDim + 2 * lane;
            for (;;) {
              bool pending = false;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (v[j].y != tg || v[j].w != tg) {
                  v[j] = ld_ll(row + 64 * j);
                  pending = true;
                }
              if (!pending) break;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              x[2 * j] = __uint_as_float(v[j].x);
              x[2 * j + 1] = __uint_as_float(v[j].z);
            }
          }
        }
        if (!ln_ready) {  // the LayerNorm weights were requested before the poll
          cp_async_wait_all();
          bar_consumers();
          ln_ready = true;
        }
        if (b < B) ln_rows_to_planes(x, b, tail ? 2 : 1, !tail);
        else zero_planes(b);
      }
    } else if (p == 1) {
      for (int item = cta; item < n_items; item += G) attention_item(li, item, item == cta);
      attention_merge(li);
    }

    for (int pass = 0; pass < n_pass; ++pass) {
      const int cb0 = pass * 8;  // first candidate of this pass (K = 4096 phase only)
      if (k4) {
        // input of the second MLP matrix: 4096 f16-exact values per candidate as {half2, tag} pairs:
        // 1024 units of 4 values per candidate, 8 candidates per pass, 32 units per thread in rounds of 8
        const uint32_t tg = tag_of(li, 4);
        if (pass > 0) bar_consumers();  // the previous pass still reads xs / partial
#pragma unroll 1
        for (int rnd = 0; rnd < 2; ++rnd) {
          // round rnd: candidates cb0 + 4 rnd .. + 3; thread owns units tid + 256 u (u = 0..3) of each:
          // 16 loads in flight per thread (only the first round of a phase really waits)
          uint4 v[4][4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int u = 0; u < 4; ++u) v[i][u] = make_uint4(0, ~tg, 0, ~tg);
          for (;;) {
            bool pending = false;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int b = cb0 + 4 * rnd + i;
              const uint2 *src = a.ll_m + rep * m_rep + size_t(b) * (kFF / 2);
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (b < B && (v[i][u].y != tg || v[i][u].w != tg)) {
                  v[i][u] = ld_ll(src + 2 * (tid + u * M2_CONSUMERS));
                  pending = true;
                }
            }
            if (!pending) break;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int b = cb0 + 4 * rnd + i, bl = 4 * rnd + i;
#pragma unroll
            for (int u = 0; u < 4; ++u)
              *reinterpret_cast<uint2 *>(xs + bl * M4_XP4 + 4 * (tid + u * M2_CONSUMERS)) =
                  b < B ? make_uint2(v[i][u].x, v[i][u].z) : make_uint2(0u, 0u);
          }
        }
      }
      bar_consumers();

      // ---------------- GEMV on the tensor cores: D[cand x 8 rows] = X[cand x K] . W^T[K x 8 rows] ----------------
      const int rps = k4 ? M3_ROWS_K4 : M3_ROWS_K1, pitch = k4 ? M3_PITCH_K4 : M3_PITCH_K1, xp = k4 ? M4_XP4 : M4_XP1;
      const int ksteps = k4 ? 32 : 8;  // 16-wide k steps of this warp's K / 8 slice
      const int kbase = warp * ksteps * 16;
      int row0, rows_cta;
      slice(N, row0, rows_cta);
      const int rlog = k4 ? 1 : 3;  // log2(rps)
      const int n_stages = (rows_cta + rps - 1) >> rlog;
      const int tiles = k4 ? 1 : NT;
      // A operand (activations): ldmatrix lane -> (row of the 16-row tile, k half)
      const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_kofs = (lane >> 4) * 8;
      // K = 4096: 8 hi rows only; rows 8-15 of the tile alias rows 0-7 (their results are dropped)
      const int a_row_eff = k4 ? (a_row & 7) : a_row;
      const uint32_t xaddr = smem_u32(xs) + uint32_t(a_row_eff * xp + kbase + a_kofs) * 2u;
      // B operand (weights): ldmatrix.x4 lane -> (row of the 8-row group, 8-wide k chunk of two k steps)
      const int b_row = lane & 7, b_kofs = (lane >> 3) * 8;
      uint32_t c_save = c_it;
      for (int s0 = 0; s0 < n_stages; s0 += M3_GROUP) {
        const int gs = min(M3_GROUP, n_stages - s0);
        const int grow0 = s0 * rps, grows = min(gs * rps, rows_cta - grow0);
        const int ngrp = (grows + 7) >> 3;  // 8-row groups of this stage group (<= 4; 1 at K = 4096)
        // deferred release (ar_mega3.cuh): the previous group's slots go back when this group's arithmetic starts
        __syncwarp();
        if (lane == 0)
          for (int j = 0; j < rel_n; ++j) mbar_arrive(&empty[(rel_it + j) % STAGES]);
        rel_n = 0;
        if (pass == 0) {
          for (;;) {
            bool ok = true;
#pragma unroll
            for (int j = 0; j < M3_GROUP; ++j)
              if (j < gs) ok = mbar_test_wait(&full[(c_it + j) % STAGES], ((c_it + j) / STAGES) & 1u) && ok;
            if (ok) break;
          }
        }
        uint32_t baddr[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          int r = g * 8 + b_row;
          if (r >= grows) r = 0;  // rows past the slice: any resident row (results unused)
          const uint32_t slot = (c_it + uint32_t(r >> rlog)) % STAGES;
          baddr[g] = smem_u32(ring) + uint32_t(slot * M3_STAGE_SMEM + (r & (rps - 1)) * pitch + (kbase + b_kofs) * 2);
        }
        float acc[4][4];
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[g][q] = 0.f;
        for (int ks = 0; ks < ksteps; ks += 2) {
          uint32_t fa[2][NT][4];  // [k step][tile]
#pragma unroll
          for (int kk = 0; kk < 2; ++kk)
#pragma unroll
            for (int t = 0; t < NT; ++t)
              if (t < tiles)
                ldmatrix_x4(xaddr + uint32_t(t * 16 * xp) * 2u + (ks + kk) * 32, fa[kk][t][0], fa[kk][t][1], fa[kk][t][2], fa[kk][t][3]);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (g < ngrp) {
              uint32_t fb[4];
              ldmatrix_x4(baddr[g] + ks * 32, fb[0], fb[1], fb[2], fb[3]);
#pragma unroll
              for (int t = 0; t < NT; ++t)
                if (t < tiles) {
                  mma_16816(acc[g], fa[0][t][0], fa[0][t][1], fa[0][t][2], fa[0][t][3], fb[0], fb[1]);
                  mma_16816(acc[g], fa[1][t][0], fa[1][t][1], fa[1][t][2], fa[1][t][3], fb[2], fb[3]);
                }
            }
          }
        }
        if (pass == n_pass - 1) {  // the stages are needed again by the next pass otherwise
          rel_it = c_it;
          rel_n = gs;
        }
        c_it += gs;
        // partial tiles: [warp][row of the group][candidate of the pass]
        {
          const int c = lane >> 2, r = (lane & 3) * 2;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (g < ngrp) {
              float *pt = partial + (warp * 32 + g * 8 + r) * BT;
              if (BT == 16 && !k4) {
                pt[c] = acc[g][0]; pt[BT + c] = acc[g][1]; pt[c + 8] = acc[g][2]; pt[BT + c + 8] = acc[g][3];
              } else if (k4) {
                pt[c] = acc[g][0]; pt[BT + c] = acc[g][1];  // rows 8-15 of the tile are aliases
              } else {
                pt[c] = acc[g][0] + acc[g][2]; pt[BT + c] = acc[g][1] + acc[g][3];  // hi + lo
              }
            }
          }
        }
        bar_consumers();
        // first K / V tile of this CTA's attention item: the planes it overwrites were last read by the
        // MMAs above (QKV rows of a CTA are one stage group); lands during the epilogue and the exchange
        if (p == 0 && !tail && cta < n_items) prefetch_kv(li, cta, first_tile(cta));
        // ---------------- epilogue ----------------
        const int ncand = k4 ? 8 : BT;  // candidates covered by the partial tiles of this pass
        for (int e0 = 0; e0 < grows * ncand; e0 += M2_CONSUMERS) {
          const int e = e0 + tid;
          const int r = e / ncand, bl = e % ncand, b = (k4 ? cb0 : 0) + bl, n = row0 + grow0 + r;
          const bool act = e < grows * ncand && b < B;
          float v = 0.f;
          if (act) {
#pragma unroll
            for (int w = 0; w < GV_WARPS; ++w) v += partial[(w * 32 + r) * BT + bl];
            v += bias[n];
            if (kind == 2) v = gelu16(v);
          }
          // (c_fc: rows n, n + 1 of one candidate sit `ncand` lanes apart; its slices start on even rows)
          const float v_next = __shfl_down_sync(0xffffffffu, v, BT == 16 ? 16 : 8);
          if (act) {
            if (kind == 0) {
              const __half hv16 = __float2half_rn(v);
              st_ll(a.ll_qkv + size_t(b) * 3072 + n, __half2float(hv16), out_tag);
              const int which = n >> 10, c = n & 1023;
              if (which != 0) {
                __half *cache = (which == 1 ? a.kc : a.vc) + size_t(li) * layer_kv;
                cache[(size_t(b) * kHeads + (c >> 6)) * size_t(a.P) * kHeadDim + size_t(a.n_past) * kHeadDim + (c & 63)] = hv16;
              }
            } else if (kind == 1) {
              const float o = hown[(n - own_row0) * BT + b] + v;
              uint2 *dst = (p == 1 ? a.ll_h2 : a.ll_h) + size_t(b) * kDim + n;
              for (int rr = 0; rr < nrep; ++rr) st_ll(dst + rr * h_rep, o, out_tag);
            } else if (kind == 2) {
              if ((r & 1) == 0) {
                const __half2 h2 = __floats2half2_rn(v, v_next);  // exact: gelu16 outputs are f16 values
                uint2 *dst = a.ll_m + size_t(b) * (kFF / 2) + (n >> 1);
                for (int rr = 0; rr < nrep; ++rr) st_ll_u32(dst + rr * m_rep, *reinterpret_cast<const uint32_t *>(&h2), out_tag);
              }
            } else {
              a.logits[size_t(b) * N + n] = v;
            }
          }
        }
        if (s0 + M3_GROUP < n_stages) bar_consumers();  // the next group overwrites the partial tiles
      }
      if (pass + 1 < n_pass) c_it = c_save;  // second pass: same stages again
    }
  }
}

}  // namespace tts
