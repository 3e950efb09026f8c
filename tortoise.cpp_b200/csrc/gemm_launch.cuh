// gemm_launch.cuh -- picks the tensor-core GEMM implementation for one TGemmArgs problem:
// tcgen05/TMEM (tc5gemm.cuh) whenever K is a multiple of its 64-wide K block, otherwise the
// HMMA (wmma) kernel.  TTS_NO_TCGEN05=1 forces the HMMA kernel (A/B testing, parity tests).
#pragma once
#include <stdlib.h>

#include "tc5v2.cuh"

namespace tts {

static inline bool tc5_enabled() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("TTS_NO_TCGEN05");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

static inline bool tc5_v1_forced() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("TTS_TC5_V1");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// returns the number of partial entries per (sequence, group) when the fused GroupNorm
// statistics (g.gn_partial) were produced, 0 otherwise
static inline int launch_gemm(const Launcher &L, const TGemmArgs &g_in) {
  TGemmArgs g = g_in;
  if (tc5_enabled() && !tc5_v1_forced() && tc5v2_supported(g)) return launch_tc5v2(L, g);
  if (g.Tseq) throw ArgError("internal: per-sequence lengths need the tcgen05 GEMM path");
  if (tc5_enabled() && g.K % T5_BK == 0) {
    const bool alo = g.Alo != nullptr, wlo = g.Wlo != nullptr;
    const int nseq = g.M / g.T;                      // M = nseq * T always
    const int mt = (g.T + T5_BM - 1) / T5_BM;        // tiles per sequence
    // BN = 64 halves the A re-reads; BN = 32 doubles the CTA count when the grid would be small
    const bool want_gn = g.gn_partial != nullptr && g.N == 1024;
    if (!want_gn) g.gn_partial = nullptr;
    g.gn_mtiles = mt;
    const bool bn64 = !want_gn && ((g.N + 63) / 64) * mt * nseq >= 120;  // measured: BN = 32 (more CTAs) wins for N = 1024
    // deep pipeline (8 x 20-24 KB stages) for single-plane operands (the convolutions): these
    // GEMMs are small (M = 2S ~ 400 rows) and latency-bound; 4 stages when lo planes double a stage
    const bool deep = !alo && !wlo;
    auto go = [&](auto kern, int BN, int stages) {
      ensure_smem_attr(kern, tc5_smem_bytes(BN, stages, true, true) > 227 * 1024 ? tc5_smem_bytes(BN, stages, alo, wlo)
                                                                                  : tc5_smem_bytes(BN, stages, true, true));
      L(kern, dim3((g.N + BN - 1) / BN, mt, nseq), dim3(T5_THREADS), tc5_smem_bytes(BN, stages, alo, wlo), g);
    };
    if (bn64) {
      if (deep) go(tc5gemm_kernel<64, 8>, 64, 8);
      else go(tc5gemm_kernel<64, 4>, 64, 4);
    } else {
      if (deep) go(tc5gemm_kernel<32, 8>, 32, 8);
      else go(tc5gemm_kernel<32, 4>, 32, 4);
    }
    return want_gn ? mt : 0;
  }
  ensure_smem_attr(tgemm_kernel, tgemm_smem_bytes());
  dim3 grid((g.N + TG_BN - 1) / TG_BN, (g.M + TG_BM - 1) / TG_BM);
  L(tgemm_kernel, grid, dim3(128), tgemm_smem_bytes(), g);
  return 0;
}

}  // namespace tts
