// ar_mega2.cuh -- AR decode step as one persistent kernel (f32 parity mode; also the shared
// protocol pieces of ar_mega3.cuh / ar_mega4.cuh).  No device-wide barriers.
//
// Same math as wsgemv.cuh (reference graph autoregressive_graph(fake_inputs=false),
// main.cpp:2668-3029).  What changed, and why (measured with the in-kernel clock trace,
// profiles/r01_decode_step.md): with device-wide barriers a phase boundary cost 3000-4500 cycles
// (bar.sync -> MEMBAR.GPU + RED -> one thread polling the counter -> bar.sync -> the L2 round trip
// that finally fetches the activations) and the LayerNorm prologue another 3000 (two more dependent
// L2 round trips).  Here
//   * activations cross CTAs as (value, tag) pairs written with ONE 8-byte store per element
//     (the low-latency protocol of collective libraries): no fence, no counter, no separate
//     flag round trip -- a consumer polls the data itself (16-byte volatile loads = two pairs)
//     and a pair is valid once its tag equals the tag of (launch, layer, phase).  Tags are
//     unique per launch, so the buffers are never cleared;
//   * a phase therefore costs store -> L2 -> load, and everything constant a phase needs
//     (LayerNorm weights, bias) is requested BEFORE the poll;
//   * attention is split over (candidate, head, key range) items; the K/V tile of a CTA's item
//     is prefetched into shared memory with cp.async while the QKV phase runs; each item emits
//     an unnormalised partial (acc[64], max, sum) and the c_proj prologue merges the partials
//     (exact log-sum-exp merge), so the merge costs no extra exchange;
//   * per-stage dot products are kept per lane and reduced with shuffles once per phase chunk
//     (independent chains) instead of once per stage;
//   * every vector that ALL CTAs read (h, h2, m, attention partials) is written to nrep replicas
//     (<= M2_REP, default 2) and CTA c polls replica c % nrep: 148 CTAs spinning on the same 64 L2
//     lines queue on the slices that own them; more replicas multiply the producers' stores
//     (same-box A/B of the whole step: 8 -> 726 us, 4 -> 680, 2 -> 665, 1 -> 692);
//   * the step is ONE loop over 121 GEMV phases with a single copy of the GEMV body: the
//     straight-line version (5 inlined copies, 160 KB of SASS) re-fetched its code from L2 in
//     every phase;
//   * the MLP hidden vector (f16-exact after the reference's fp16 GELU table) crosses as
//     {half2, tag}: half the bytes of the largest per-phase poll;
//   * up to 4 candidates ride on one weight stream (register-resident activations).
// Weight streaming is unchanged: a dedicated producer warp walks the whole step's weight slices
// of this CTA, in consumption order, through a ring of 16 KB stages with 1-D TMA bulk copies.
#pragma once
#include "ar_kernels.cuh"
#include "wsgemv.cuh"

namespace tts {

struct MegaLayer {
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
  const void *w_qkv, *w_proj, *w_fc, *w_proj2;
  const float *b_qkv, *b_proj, *b_fc, *b_proj2;
};

constexpr int M2_CONSUMERS = 256;
constexpr int M2_THREADS = M2_CONSUMERS + 128;  // + one warpgroup: warp 8 = weight-stream producer, warps 9-11 exit at once
constexpr int M2_CONSUMER_REGS = 232, M2_PRODUCER_REGS = 24;  // setmaxnreg budgets (the launch allocates 168 x 384)
constexpr int M2_SMAX = 8;       // max key splits per (candidate, head)
constexpr int M2_REC = 68;       // pairs per attention partial record: 64 acc, max, sum, 2 pad
constexpr int M2_KV_TILE = 128;  // keys per attention item
constexpr int M2_KV_LD = 72;     // halves per K/V row in shared memory (144 B: conflict-free 16-byte reads)
constexpr int M2_REP = 8;        // replicas of the all-to-all exchange vectors

struct Mega2Args {
  const MegaLayer *layers;  // [30], device
  const float *lnf_w, *lnf_b, *lm0_w, *lm0_b, *lm_b;
  const void *lm_w;
  const float *mel_emb, *mel_pos;
  const int *tokens;
  // (value, tag) exchange buffers; h / h2 / m / att hold M2_REP replicas ([rep][...]); m is
  // {half2, tag} (kFF / 2 pairs per candidate)
  uint2 *ll_h, *ll_h2, *ll_qkv, *ll_m, *ll_att;
  float *logits;
  __half *kc, *vc;  // [30][Bmax][16][P][64]
  int B, Bmax, P, n_past, pos_id;
  unsigned int epoch;  // unique per launch (1 .. 2^24-1)
  long long *dbg;      // optional clock trace (TTS_MEGA_TRACE=1: CTA 0, cycle stamps; =2: every CTA, 4 globaltimer stamps per phase)
  int dbg_mode;
  int b0;  // first candidate of this launch (more than 4 candidates = several launches per step, 4 at a time)
  int evict_first;  // 1: weight stream with the L2 evict-first hint (TTS_MEGA_NOEVICT=1 turns it off)
  int keys_per_split;  // attention: keys per (candidate, head) item before splitting (TTS_MEGA_KPS)
  int poll_spin;  // cycles between two polls of a missing tag (TTS_MEGA_SPIN)
  int defer;  // 1: deferred ring-slot release (default)
  int nrep;  // replicas of the all-to-all vectors in use (1..M2_REP; TTS_MEGA_REP)
};

template <int BT>
struct M2Cfg {
  static constexpr int kStages = BT == 4 ? 6 : 8;  // ring depth (16 KB stages)
  static constexpr int kChunk = BT == 1 ? 4 : (BT == 2 ? 2 : 1);  // stages loaded and reduced together (register budget)
};

template <int BT>
__host__ __device__ inline size_t mega2_smem_bytes() {
  return size_t(M2Cfg<BT>::kStages) * GV_STAGE_BYTES + 256 /*mbarriers*/ +
         size_t(GV_MAX_ROWS_PER_CTA) * 8 * BT * sizeof(float) /*partials*/ + 64 * sizeof(double) /*LN sums*/ +
         size_t(BT) * kFF * sizeof(float) /*phase input*/ + size_t(BT) * kDim * sizeof(float) /*residual*/ +
         size_t(2) * M2_KV_TILE * M2_KV_LD * sizeof(__half) /*K, V tile*/ + (128 + 64 + 8 * 64 + 32) * sizeof(float);
}

__device__ __forceinline__ uint4 ld_ll(const uint2 *p) {  // two (value, tag) pairs
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_ll_u32(uint2 *p, uint32_t bits, uint32_t tag) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(bits), "r"(tag) : "memory");
}
__device__ __forceinline__ void st_ll(uint2 *p, float val, uint32_t tag) { st_ll_u32(p, __float_as_uint(val), tag); }
__device__ __forceinline__ void cp_async_cg16(void *dst_smem, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// Between two polls of a tag that has not arrived.  Measured (tools/xchg_bench.cu): __nanosleep(40)
// costs ~1800 cycles per call on B200 whatever its argument, i.e. it quantises every exchange to
// multiples of ~1 us; a plain re-poll (one L2 round trip, ~300 cycles) is the cheapest wait.
__device__ __forceinline__ void poll_backoff(int spin) {
  if (spin > 0) {  // a short clock spin thins out the re-polls (fewer requests queued on the hot L2 lines)
    const long long t0 = clock64();
    while (clock64() - t0 < spin) {
    }
  }
}
__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

template <typename WT, int BT>
static __global__ void __launch_bounds__(M2_THREADS, 1) ar_decode_mega2_kernel(Mega2Args a) {
  constexpr int E = WTraits<WT>::kElemsPer16B;
  constexpr int KS = 32 * 4 * E;  // K elements of one warp's 2 KB slice
  constexpr int STAGES = M2Cfg<BT>::kStages;
  constexpr int CH = M2Cfg<BT>::kChunk;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char *ring = smem;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + STAGES * GV_STAGE_BYTES);
  uint64_t *empty = full + STAGES;
  float *partial = reinterpret_cast<float *>(smem + STAGES * GV_STAGE_BYTES + 256);
  double *red = reinterpret_cast<double *>(partial + GV_MAX_ROWS_PER_CTA * 8 * BT);
  float *xin = reinterpret_cast<float *>(red + 64);  // [BT][kFF] input vector of the running phase
  float *hres = xin + BT * kFF;                      // [BT][kDim] residual stream
  __half *kt = reinterpret_cast<__half *>(hres + BT * kDim);  // [128][72] K tile
  __half *vt = kt + M2_KV_TILE * M2_KV_LD;                    // [128][72] V tile
  float *sc = reinterpret_cast<float *>(vt + M2_KV_TILE * M2_KV_LD);  // [128] scores
  float *qs = sc + 128;                                               // [64] query
  float *pp = qs + 64;                                                // [8][64] partial outputs
  float *redf = pp + 8 * 64;                                          // [32] block reductions

  // the layer table (pointers) and this CTA's row slices live in shared memory: a pointer fetched
  // from global memory in front of every dependent load, or an integer division per phase, stalls
  // the in-order issue of a phase that is only a few thousand cycles long
  __shared__ MegaLayer s_layers[kLayers];
  __shared__ int s_slice[4][2];  // {row0, rows} for N = 3072, 1024, 4096, 8194

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
  // rows of a matrix owned by this CTA
  auto slice = [&](int N, int &row0, int &rows) {
    const int i = N == 3072 ? 0 : (N == kDim ? 1 : (N == kFF ? 2 : 3));
    row0 = s_slice[i][0];
    rows = s_slice[i][1];
  };

  if (warp >= M2_CONSUMERS / 32) {
    // the producer warpgroup hands its registers to the consumers (9 warps would put three on one
    // scheduler and cap everybody at 168 registers: the GEMV body spills)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(M2_PRODUCER_REGS));
    // ---------------- weight stream: every slice of the step, in consumption order ------------
    if (warp == M2_CONSUMERS / 32 && lane == 0) {
      const uint64_t pol = l2_policy_evict_first();  // see ar_mega3.cuh
      long it = 0;
      for (int sid = 0; sid <= 120; ++sid) {
        int N, K, row0, rows;
        const void *W;
        seg_shape(sid, N, K, W);
        slice(N, row0, rows);
        const size_t row_bytes = size_t(K) * sizeof(WT);
        const unsigned char *src = reinterpret_cast<const unsigned char *>(W) + size_t(row0) * row_bytes;
        const size_t total = size_t(rows) * row_bytes;
        for (size_t off = 0; off < total; off += GV_STAGE_BYTES, ++it) {
          const int slot = int(it % STAGES);
          mbar_wait(&empty[slot], uint32_t((it / STAGES) & 1) ^ 1u);
          const uint32_t bytes = uint32_t(min(size_t(GV_STAGE_BYTES), total - off));
          mbar_arrive_expect_tx(&full[slot], bytes);
          if (a.evict_first) bulk_g2s_hint(ring + size_t(slot) * GV_STAGE_BYTES, src + off, bytes, &full[slot], pol);
          else bulk_g2s(ring + size_t(slot) * GV_STAGE_BYTES, src + off, bytes, &full[slot]);
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
  long c_it = 0;  // ring stage counter (same order as the producer)
  long rel_it = 0;  // first stage of the group whose slots are still held
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
      *reinterpret_cast<float4 *>(xin + b * kFF + head * kHeadDim + d0) =
          make_float4(o[0] * inv, o[1] * inv, o[2] * inv, o[3] * inv);
    }
  };

  // =========================== the step: 30 x (QKV | c_proj | c_fc | mlp c_proj) + lm_head ==========
  // One loop, one copy of the GEMV body.  Phase p of layer li: prologue (what xin holds), then
  // xin x this CTA's rows of the phase's matrix, then the epilogue that feeds the next phase.
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
        *reinterpret_cast<float2 *>(xin + b * kFF + 2 * tid) = make_float2(hv[b][0], hv[b][1]);
        *reinterpret_cast<float2 *>(xin + b * kFF + 512 + 2 * tid) = make_float2(hv[b][2], hv[b][3]);
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
          const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&v[u].x));
          const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&v[u].z));
          *reinterpret_cast<float4 *>(xin + b * kFF + 4 * (tid + u * M2_CONSUMERS)) = make_float4(f0.x, f0.y, f1.x, f1.y);
        }
      }
    }
    bar_consumers();
    trace(20 + p);
    stamp(ph, 1);

    // ---------------- GEMV: xin (shared) x this CTA's weight rows (ring) ----------------
    // kind: 0 = QKV (f16 round trip, KV append), 1 = residual, 2 = GELU16, 3 = logits
    const int N = tail ? kMelVocab : (p == 0 ? 3072 : (p == 2 ? kFF : kDim));
    const int K = p == 3 ? kFF : kDim;
    const int kind = tail ? 3 : (p == 0 ? 0 : (p == 2 ? 2 : 1));
    const float *bias = tail ? a.lm_b : (p == 0 ? l.b_qkv : (p == 1 ? l.b_proj : (p == 2 ? l.b_fc : l.b_proj2)));
    const uint32_t out_tag = tag_of(li, p == 0 ? 1 : (p == 1 ? 3 : (p == 2 ? 4 : 5)));
    {
      // (all shapes are powers of two: no runtime division)
      constexpr int WPR1 = kDim / KS, WPR4 = kFF / KS;  // warps per weight row
      const bool k4 = p == 3;
      const int wpr = k4 ? WPR4 : WPR1, rps = k4 ? GV_WARPS / WPR4 : GV_WARPS / WPR1;
      const int rlog = 31 - __clz(rps);
      int row0, rows_cta;
      slice(N, row0, rows_cta);
      const int n_stages = (rows_cta + rps - 1) >> rlog;
      const int ks = warp & (wpr - 1), rsub = warp >> (31 - __clz(wpr));
      float ep_bias = 0.f;
      if (tid < rows_cta * BT) ep_bias = bias[row0 + tid / BT];
      float xr[BT][4][E];
#pragma unroll
      for (int b = 0; b < BT; ++b)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = ks * KS + j * (32 * E) + lane * E;
#pragma unroll
          for (int e4 = 0; e4 < E; e4 += 4) {
            const float4 v = *reinterpret_cast<const float4 *>(xin + b * kFF + k + e4);
            xr[b][j][e4 + 0] = v.x; xr[b][j][e4 + 1] = v.y; xr[b][j][e4 + 2] = v.z; xr[b][j][e4 + 3] = v.w;
          }
        }
      trace(40);
      // CH stages at a time: every wait, then every shared-memory load, then the arithmetic (the
      // stages of a phase are resident when it starts; serialising wait -> load -> FMA chain ->
      // arrive per stage was ~600 cycles per stage)
      for (int s0 = 0; s0 < n_stages; s0 += CH) {
        const int gs = min(CH, n_stages - s0);
        // Deferred release (see ar_mega3.cuh): the previous group's slots go back to the producer
        // only now, after the exchange that followed it, so the refill burst of the weight stream
        // overlaps this group's arithmetic instead of the latency-critical exchange.
        __syncwarp();
        if (lane == 0)
          for (int j = 0; j < rel_n; ++j) mbar_arrive(&empty[int((rel_it + j) % STAGES)]);
        rel_n = 0;
#pragma unroll
        for (int j = 0; j < CH; ++j)
          if (j < gs) mbar_wait(&full[int((c_it + j) % STAGES)], uint32_t(((c_it + j) / STAGES) & 1));
        trace(50);
        // straight-line, branch-free: rows past the slice read zeros, so the scheduler can interleave
        // the 4 CH BT independent FMA chains (a phase is latency-bound: ~3 cycles per instruction
        // with two warps per scheduler and dependent chains)
        uint4 wv[CH][4];
        bool valid[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const int r = (s0 + j) * rps + rsub;
          valid[j] = j < gs && r < rows_cta;
          const unsigned char *wp = ring + size_t((c_it + j) % STAGES) * GV_STAGE_BYTES + warp * 2048 + lane * 16;
#pragma unroll
          for (int q = 0; q < 4; ++q) wv[j][q] = valid[j] ? *reinterpret_cast<const uint4 *>(wp + q * 512) : make_uint4(0, 0, 0, 0);
        }
        float acc[CH][BT];
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          float aq[4][BT];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float wf[E];
            if constexpr (sizeof(WT) == 4) {
              wf[0] = __uint_as_float(wv[j][q].x);
              wf[1] = __uint_as_float(wv[j][q].y);
              wf[2] = __uint_as_float(wv[j][q].z);
              wf[3] = __uint_as_float(wv[j][q].w);
            } else {
              const __half2 *h2 = reinterpret_cast<const __half2 *>(&wv[j][q]);
#pragma unroll
              for (int qd = 0; qd < 4; ++qd) {
                const float2 f = __half22float2(h2[qd]);
                wf[2 * qd] = f.x;
                wf[2 * qd + 1] = f.y;
              }
            }
#pragma unroll
            for (int b = 0; b < BT; ++b) {
              float t = wf[0] * xr[b][q][0];
#pragma unroll
              for (int e = 1; e < E; ++e) t = fmaf(wf[e], xr[b][q][e], t);
              aq[q][b] = t;
            }
          }
#pragma unroll
          for (int b = 0; b < BT; ++b) acc[j][b] = (aq[0][b] + aq[2][b]) + (aq[1][b] + aq[3][b]);
        }
        rel_it = c_it;  // released when the NEXT group of stages starts (see above)
        rel_n = gs;
        c_it += gs;
        if (!a.defer) {  // A/B knob (TTS_MEGA_NODEFER=1): hand the slots back at once
          __syncwarp();
          if (lane == 0)
            for (int j = 0; j < rel_n; ++j) mbar_arrive(&empty[int((rel_it + j) % STAGES)]);
          rel_n = 0;
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) {
#pragma unroll
          for (int b = 0; b < BT; ++b) {
            const float v = warp_sum(acc[j][b]);
            if (lane == 0 && valid[j]) partial[(((s0 + j) * rps + rsub) * 8 + ks) * BT + b] = v;
          }
        }
      }
      trace(45);
      bar_consumers();
      trace(30 + p);
      stamp(ph, 2);
      // ---------------- epilogue: one output element per thread ----------------
      const bool act = tid < rows_cta * BT && (tid % BT) < B;
      const int r = tid / BT, b = tid % BT, n = row0 + r;
      float v = 0.f;
      if (act) {
        for (int w = 0; w < wpr; ++w) v += partial[(r * 8 + w) * BT + b];
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
#pragma unroll
          for (int rr = 0; rr < nrep; ++rr) st_ll(dst + rr * h_rep, o, out_tag);
        } else if (kind == 2) {
          if ((r & 1) == 0) {
            const __half2 h2 = __floats2half2_rn(v, v_next);  // exact: gelu16 outputs are f16 values
            uint2 *dst = a.ll_m + size_t(b) * (kFF / 2) + (n >> 1);
#pragma unroll
            for (int rr = 0; rr < nrep; ++rr) st_ll_u32(dst + rr * m_rep, *reinterpret_cast<const uint32_t *>(&h2), out_tag);
          }
        } else {
          a.logits[size_t(a.b0 + b) * N + n] = v;
        }
      }
    }
    stamp(ph, 3);
  }
  trace(40);
}

}  // namespace tts
