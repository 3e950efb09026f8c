// sampler.cpp -- RNG + logits post-processing + sampling, padding and trimming: the host
// logic of the AR stage (main.cpp:4510-4532, 4562-4720, 4753-4806, 4873-4915).
//
// Two implementations of process_logits_and_sample:
//  * sample_literal(): the reference's algorithm step for step (full sorts of 8194 floats);
//  * sample_fast(): same results in O(V + k log k).  Only the >= 50 logits that survive top-k
//    can have non-zero probability, exp(lowest()) == 0 exactly, and adding 0.0f is exact, so
//    every float sum the reference forms over 8194 entries equals the same sum over the
//    survivors in the same order.  If two survivors tie (the only case where std::sort's
//    unspecified order of equal keys could matter) it defers to sample_literal().
#include <algorithm>
#include <functional>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <limits>
#include <random>
#include <vector>

#include "rng.h"

namespace tts_host {

constexpr int V = 8194;

static int multinomial(Rng &r, const float *probs, int n) {
  float sample = r.distribution(r.generator);
  sample = r.distribution(r.generator);  // two draws, second used (main.cpp:4708-4709)
  float cum = 0;
  for (int i = 0; i < n; ++i) {
    cum += probs[i];
    if (cum >= sample) return i;
  }
  return V - 1;
}

static void penalise(std::vector<float> &lg, const int32_t *prev, int n_prev) {
  // gather -> apply_penalty(2.0) -> scatter (main.cpp:4770-4775): every scatter writes a value
  // derived from the ORIGINAL logit, so duplicates in prev are idempotent.
  std::vector<float> orig(n_prev);
  for (int i = 0; i < n_prev; ++i) orig[i] = lg[prev[i]];
  for (int i = 0; i < n_prev; ++i) lg[prev[i]] = orig[i] < 0 ? orig[i] * 2.0f : orig[i] / 2.0f;
}

int sample_literal_one(Rng &r, const float *logits_row, const int32_t *prev, int n_prev) {
  std::vector<float> lg(logits_row, logits_row + V);
  penalise(lg, prev, n_prev);
  const float temp = 0.8;
  for (float &x : lg) x /= temp;
  {  // top_k_inplace(50)
    std::vector<float> s(lg);
    std::sort(s.begin(), s.end());
    const float kth = s[s.size() - 50];
    for (float &x : lg)
      if (x < kth) x = std::numeric_limits<float>::lowest();
  }
  {  // top_p_inplace
    std::vector<std::pair<float, int>> pairs;
    for (int i = 0; i < V; ++i) pairs.push_back(std::make_pair(lg[i], i));
    std::sort(pairs.begin(), pairs.end(),
              [](const std::pair<float, int> &a, const std::pair<float, int> &b) { return a.first < b.first; });
    std::vector<float> sl(V);
    float sum = 0;
    for (int i = 0; i < V; ++i) {
      sl[i] = std::exp(pairs[i].first);
      sum += sl[i];
    }
    for (int i = 0; i < V; ++i) sl[i] /= sum;
    for (int i = 1; i < V; ++i) sl[i] += sl[i - 1];
    for (int i = 0; i < V - 1; ++i)
      if (sl[i] <= 0.2) lg[pairs[i].second] = std::numeric_limits<float>::lowest();
  }
  float sum = 0;
  for (float &x : lg) {
    x = std::exp(x);
    sum += x;
  }
  for (float &x : lg) x /= sum;
  return multinomial(r, lg.data(), V);
}

// 50th largest value of x[0..V) = root of a min-heap of the 50 largest seen so far: one pass in
// which almost every element fails a single (well-predicted) comparison.  (An AVX2 variant testing
// 8 elements per instruction measured no faster than this loop -- ~10 us per row either way.)
static float kth_largest_50(const float *x) {
  float heap[50];
  for (int i = 0; i < 50; ++i) heap[i] = x[i];
  std::make_heap(heap, heap + 50, std::greater<float>());
  for (int i = 50; i < V; ++i) {
    const float v = x[i];
    if (v > heap[0]) {
      std::pop_heap(heap, heap + 50, std::greater<float>());
      heap[49] = v;
      std::push_heap(heap, heap + 50, std::greater<float>());
    }
  }
  return heap[0];
}

// Shared tail of the fast paths: top-p, final softmax and the multinomial draw over the top-k
// survivors `surv` ((value / temp, index) in INDEX order).  Returns -1 when it must defer to the
// literal path (ties among survivors: the only case where std::sort's unspecified order of equal
// keys could matter).  Draws from the generator only after it has decided not to defer.
static int finish_from_survivors(Rng &r, const std::vector<std::pair<float, int>> &surv, float *logprob) {
  std::vector<std::pair<float, int>> asc(surv);
  std::sort(asc.begin(), asc.end());  // by value, then index
  for (size_t i = 1; i < asc.size(); ++i)
    if (asc[i].first == asc[i - 1].first) return -1;
  // top-p over the ascending survivors; the V - |surv| masked entries contribute exact zeros
  // and sort before every survivor, so index i of the full sorted array = i - n_masked here.
  const int ns = int(asc.size());
  std::vector<float> e(ns);
  float sum = 0;
  for (int i = 0; i < ns; ++i) {
    e[i] = std::exp(asc[i].first);
    sum += e[i];
  }
  std::vector<int> dead;  // indices cut by top-p (a handful)
  float cum = 0;
  for (int i = 0; i < ns; ++i) {
    const float p = e[i] / sum;
    cum = (i == 0) ? p : cum + p;  // masked prefix sums to exactly 0.0f
    if (i < ns - 1 && cum <= 0.2) dead.push_back(asc[i].second);  // the last (largest) entry is never cut
  }
  // final softmax + multinomial in index order over what is left
  std::vector<std::pair<float, int>> fin;
  float fsum = 0;
  for (const auto &pr : surv)
    if (std::find(dead.begin(), dead.end(), pr.second) == dead.end()) {
      const float ex = std::exp(pr.first);
      fin.push_back(std::make_pair(ex, pr.second));
      fsum += ex;
    }
  float sample = r.distribution(r.generator);
  sample = r.distribution(r.generator);
  int pick = V - 1;
  float pickp = 0.f;
  if (!(0.0f < sample)) {
    pick = 0;  // cumulative 0 >= sample already holds at index 0 (main.cpp:4713-4716)
    pickp = (!fin.empty() && fin[0].second == 0) ? fin[0].first / fsum : 0.f;
  } else {
    float c = 0;
    bool found = false;
    for (const auto &pr : fin) {
      const float p = pr.first / fsum;
      c += p;
      if (c >= sample) {
        pick = pr.second;
        pickp = p;
        found = true;
        break;
      }
    }
    if (!found) {
      pick = V - 1;
      pickp = 0.f;
      for (const auto &pr : fin)
        if (pr.second == V - 1) pickp = pr.first / fsum;
    }
  }
  if (logprob) *logprob = pickp > 0.f ? std::log(pickp) : -std::numeric_limits<float>::infinity();
  return pick;
}

// returns -1 when it must defer to the literal path
//
// Cost matters: the sampler sits between two decode steps (a step is ~0.45 ms on B200, and 16
// candidates are sampled one after the other).  The 50th-largest value is found with a 50-entry
// min-heap in ONE pass over the penalised raw logits (division by the temperature is monotone,
// so it is applied to the handful of candidates around the threshold only; the survivor test
// itself is done on the divided values exactly as the reference does it).
static int sample_fast_one(Rng &r, const float *logits_row, const int32_t *prev, int n_prev, float *logprob) {
  thread_local std::vector<float> lg;
  lg.assign(logits_row, logits_row + V);
  penalise(lg, prev, n_prev);
  const float temp = 0.8;
  const float kth_raw = kth_largest_50(lg.data());
  const float kth = kth_raw / temp;  // == s[V - 50] of the divided array (x -> x / temp is monotone)
  // survivors: !(x / temp < kth).  x >= kth_raw always survives; below it only values whose
  // quotient rounds up to kth can -- they are within a few ulp of kth_raw.
  const float slack = std::fabs(kth_raw) * 4e-7f + 1e-37f;
  const float lo = kth_raw - slack;
  std::vector<std::pair<float, int>> surv;  // (value / temp, index), in index order
  surv.reserve(64);
  for (int i = 0; i < V; ++i) {
    const float x = lg[i];
    if (x >= lo) {
      const float q = x / temp;
      if (!(q < kth)) surv.push_back(std::make_pair(q, i));
    }
  }
  if ((int)surv.size() == V) return -1;  // degenerate: nothing was cut
  return finish_from_survivors(r, surv, logprob);
}

// The same sampler fed with the device-side pre-selection (tts_ar_step_topk): `n` (value, index)
// pairs holding the n largest RAW logits of the row (ties at the cut: lowest indices).  Exactness
// argument: let m = the smallest raw value among the entries; every entry of the row that is NOT in
// the set has raw <= m, and the repetition penalty only lowers a value, so after the penalty it is
// still <= m.  The top-k threshold (50th largest penalised value) computed over the set is >= the
// (50 + #penalised)-th largest raw value; if m lies strictly below the survivor window
// [kth_raw - slack, inf) no outside entry can be a survivor, the threshold over the set equals the
// threshold over the row, and everything downstream only looks at survivors -- the result is the
// one sample_fast_one produces on the full row.  Returns -2 when that cannot be shown (the caller
// fetches the full row), -1 on ties among survivors (literal path, also on the full row).
int sample_sparse_one(Rng &r, const float *vals, const int32_t *idx, int n, const int32_t *prev, int n_prev,
                      float *logprob) {
  if (n < 52) return -2;
  std::vector<std::pair<int, float>> ent(n);  // (index, penalised value), sorted by index below
  float m = std::numeric_limits<float>::infinity();
  for (int i = 0; i < n; ++i) {
    if (idx[i] < 0 || idx[i] >= V) return -2;
    float x = vals[i];
    m = std::min(m, x);
    for (int j = 0; j < n_prev; ++j)
      if (prev[j] == idx[i]) {  // every scatter derives from the ORIGINAL logit: duplicates are idempotent
        x = vals[i] < 0 ? vals[i] * 2.0f : vals[i] / 2.0f;
        break;
      }
    ent[i] = std::make_pair(int(idx[i]), x);
  }
  std::sort(ent.begin(), ent.end());
  for (int i = 1; i < n; ++i)
    if (ent[i].first == ent[i - 1].first) return -2;  // malformed set
  std::vector<float> pv(n);
  for (int i = 0; i < n; ++i) pv[i] = ent[i].second;
  std::nth_element(pv.begin(), pv.begin() + 49, pv.end(), std::greater<float>());
  const float temp = 0.8;
  const float kth_raw = pv[49];
  const float kth = kth_raw / temp;
  const float slack = std::fabs(kth_raw) * 4e-7f + 1e-37f;
  const float lo = kth_raw - slack;
  if (!(m < lo)) return -2;  // an entry outside the set could reach the survivor window
  std::vector<std::pair<float, int>> surv;
  surv.reserve(64);
  for (int i = 0; i < n; ++i) {
    const float x = ent[i].second;
    if (x >= lo) {
      const float q = x / temp;
      if (!(q < kth)) surv.push_back(std::make_pair(q, ent[i].first));
    }
  }
  return finish_from_survivors(r, surv, logprob);
}

int sample_one(Rng &r, const float *logits_row, const int32_t *prev, int n_prev, float *logprob) {
  // (the fast path draws from the generator only after it has decided not to defer)
  const int s = sample_fast_one(r, logits_row, prev, n_prev, logprob);
  if (s >= 0) return s;
  // the literal path does not track the picked probability: report "unknown" (NaN) so that the
  // caller's mean-log-prob score skips this step instead of counting a perfect 0
  if (logprob) *logprob = std::numeric_limits<float>::quiet_NaN();
  return sample_literal_one(r, logits_row, prev, n_prev);
}

void apply_padding(std::vector<int32_t> &vec) {  // main.cpp:4510-4532
  while (!vec.empty() && vec.back() == 8139) vec.pop_back();  // sic: 8139, not 8193
  for (size_t i = vec.size(); i < 500; ++i) vec.push_back(83);
  vec[vec.size() - 3] = 45;
  vec[vec.size() - 2] = 45;
  vec[vec.size() - 1] = 248;
  vec.push_back(8193);
  vec.insert(vec.begin(), 8192);
}

int trim_count(const int32_t *codes500) {  // main.cpp:4894-4911
  int calm = 0;
  for (int c = 0; c < 500; ++c) {
    calm = codes500[c] == 83 ? calm + 1 : 0;
    if (calm > 8) return c;
  }
  return 500;
}

}  // namespace tts_host
