// host_math.cpp -- see host_math.h.  Compiled with -ffp-contract=off: the reference's host
// code is built for baseline x86-64 (no FMA), so no fused multiply-adds may appear here.
#include "host_math.h"

#include <cmath>
#include <numeric>
#include <functional>

namespace tts_host {

std::vector<int> timestep_map(int n_steps) {
  // guided-diffusion space_timesteps with one section: frac_stride = (4000-1)/(n-1),
  // cur_idx accumulates in double, entries are round(cur_idx).  For n = 80 this reproduces
  // the literal table at main.cpp:5641-5648 (checked in tests/test_host_cpu.py).
  std::vector<int> m;
  if (n_steps <= 1) { m.push_back(0); return m; }
  const double frac = double(4000 - 1) / double(n_steps - 1);
  double cur = 0.0;
  for (int i = 0; i < n_steps; ++i) {
    m.push_back(int(std::round(cur)));
    cur += frac;
  }
  return m;
}

std::vector<DdpmStep> ddpm_schedule(int n_steps) {
  const int n0 = 4000;
  // get_beta_schedule (main.cpp:5390-5400): note the (float) cast of (end - start)
  std::vector<double> betas;
  {
    const double scale = 1000.0 / n0;
    const double beta_start = scale * 0.0001, beta_end = scale * 0.02;
    for (int i = 0; i < n0; ++i) betas.push_back(beta_start + i * (float)(beta_end - beta_start) / (n0 - 1));
  }
  auto cumprod = [](const std::vector<double> &b) {  // get_alphas_cumulative_product (5370-5388): 1.0f - beta
    std::vector<double> a;
    for (double x : b) a.push_back(1.0f - x);
    std::vector<double> r(a.size());
    std::partial_sum(a.begin(), a.end(), r.begin(), std::multiplies<double>());
    return r;
  };
  std::vector<double> acp = cumprod(betas);
  const std::vector<int> tmap = timestep_map(n_steps);
  // respacing (main.cpp:5662-5670): last_alpha_cumprod is a FLOAT in the reference
  float last = 1.0;
  std::vector<double> nb;
  for (int i : tmap) {
    nb.push_back(1 - (acp[i] / last));
    last = acp[i];
  }
  acp = cumprod(nb);
  const int n = n_steps;
  std::vector<double> prev(n);
  prev[0] = 1.0f;
  for (int i = 1; i < n; ++i) prev[i] = acp[i - 1];
  std::vector<double> post_var(n), post_logvar(n), c1(n), c2(n), sr(n), srm1(n);
  for (int i = 0; i < n; ++i) {
    post_var[i] = nb[i] * (1.0 - prev[i]) / (1.0 - acp[i]);
    c1[i] = nb[i] * std::sqrt(prev[i]) / (1.0 - acp[i]);
    c2[i] = (1.0 - prev[i]) * std::sqrt(1.0 - nb[i]) / (1.0 - acp[i]);
    sr[i] = std::sqrt(1.0f / acp[i]);
    srm1[i] = std::sqrt(1.0f / acp[i] - 1);
  }
  post_logvar[0] = n > 1 ? std::log(post_var[1]) : std::log(post_var[0]);
  for (int i = 1; i < n; ++i) post_logvar[i] = std::log(post_var[i]);
  std::vector<DdpmStep> out(n);
  for (int d = 0; d < n; ++d) {
    const int idx = n - 1 - d;
    DdpmStep &s = out[d];
    s.max_log = (float)std::log(nb[idx]);       // "max_log" (main.cpp:5988)
    s.min_log = (float)post_logvar[idx];        // "min_log" (main.cpp:5990)
    s.cfk = 2.0f * (1 - (float)idx / float(n)); // main.cpp:5992-5994
    s.sqrt_recip = (float)sr[idx];
    s.sqrt_recipm1 = (float)srm1[idx];
    s.coef1 = (float)c1[idx];
    s.coef2 = (float)c2[idx];
    s.last = idx == 0;
    s.timestep = tmap[idx];
  }
  return out;
}

void timestep_embedding(int t, float *out) {
  const int half = 512, max_period = 10000;
  for (int i = 0; i < half; ++i) {
    float freq = std::exp(-std::log((double)max_period) * static_cast<float>(i) / half);
    float arg = static_cast<float>(t) * freq;
    // the reference's unqualified cos(arg)/sin(arg) resolve to the C double functions; the
    // result is narrowed on push_back (pinned bit-exactly by tests/golden/hostfn.npz)
    out[i] = (float)std::cos((double)arg);
    out[half + i] = (float)std::sin((double)arg);
  }
}

std::vector<int> relative_position_table(int n) {
  std::vector<int> t(n > 0 ? n : 1);
  for (int rp = 0; rp < (int)t.size(); ++rp) {
    int val_if_large = 8 + (int)(std::log(float(rp) / 8) / std::log(64.0 / 8.0) * (16.0 - 8.0));
    if (val_if_large > 15) val_if_large = 15;
    t[rp] = rp < 8 ? rp : val_if_large;
  }
  return t;
}

std::vector<int> relative_position_buckets(int n) {
  const std::vector<int> t = relative_position_table(n);
  std::vector<int> m(size_t(n) * n);
  for (int i = 0; i < n; ++i)
    for (int c = 0; c < n; ++c) m[size_t(i) * n + c] = (i < c ? 16 : 0) + t[std::abs(c - i)];
  return m;
}

std::vector<int> upscale_index(int L, int S) {
  std::vector<int> idx(S);
  const float sf0 = (float)S / L;
  for (int64_t i0 = 0; i0 < S; ++i0) idx[i0] = (int)(int64_t)(i0 / sf0);
  return idx;
}

}  // namespace tts_host
