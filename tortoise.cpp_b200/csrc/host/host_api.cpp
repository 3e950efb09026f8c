// host_api.cpp -- extern "C" surface of the host-only pieces (include/tortoise_host.h).
#include <cstring>
#include <map>
#include <random>
#include <string>
#include <vector>

#include "../../../include/tortoise_host.h"
#include "host_math.h"
#include "rng.h"

namespace tts_host {
bool scrape_vocab(const std::string &path, std::map<std::string, int32_t> &vocab);
std::vector<int32_t> tokenize_message(const std::map<std::string, int32_t> &vocab, std::string message, bool warn);
int sample_one(Rng &r, const float *logits_row, const int32_t *prev, int n_prev, float *logprob);
int sample_sparse_one(Rng &r, const float *vals, const int32_t *idx, int n, const int32_t *prev, int n_prev, float *logprob);
int sample_literal_one(Rng &r, const float *logits_row, const int32_t *prev, int n_prev);
void apply_padding(std::vector<int32_t> &vec);
int trim_count(const int32_t *codes500);
bool write_wav(const char *path, const float *data, int64_t n, int sample_rate);
}  // namespace tts_host

extern "C" {

tts_rng *tts_rng_create(uint32_t seed) { return new tts_rng(seed); }
void tts_rng_seed(tts_rng *r, uint32_t seed) { r->r.generator.seed(seed); }
void tts_rng_free(tts_rng *r) { delete r; }
float tts_rng_uniform(tts_rng *r) { return r->r.distribution(r->r.generator); }
void tts_rng_normal(tts_rng *r, float *out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = r->r.normal(r->r.generator);  // double -> float, main.cpp:4698
}

int tts_host_tokenize(const char *path, const char *message, int32_t *out, int cap) {
  std::map<std::string, int32_t> vocab;
  if (!tts_host::scrape_vocab(path, vocab)) return -2;
  const std::vector<int32_t> ids = tts_host::tokenize_message(vocab, message, true);
  for (int i = 0; i < int(ids.size()) && i < cap; ++i) out[i] = ids[i];
  return int(ids.size());
}
int tts_host_vocab_size(const char *path) {
  std::map<std::string, int32_t> vocab;
  if (!tts_host::scrape_vocab(path, vocab)) return -2;
  return int(vocab.size());
}

int tts_host_sample(tts_rng *r, const float *logits, const int32_t *prev, int n_prev, int B, int32_t *out,
                    float *logprob) {
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < n_prev; ++i)
      if (prev[b * n_prev + i] < 0 || prev[b * n_prev + i] >= 8194) return -1;
  for (int b = 0; b < B; ++b)
    out[b] = tts_host::sample_one(r->r, logits + size_t(b) * 8194, prev + size_t(b) * n_prev, n_prev,
                                  logprob ? logprob + b : nullptr);
  return 0;
}
int tts_host_sample_sparse(tts_rng *r, const float *vals, const int32_t *idx, int n, const int32_t *prev, int n_prev,
                           int32_t *sample_out, float *logprob) {
  if (!r || !vals || !idx || !prev || !sample_out) return -1;
  for (int i = 0; i < n_prev; ++i)
    if (prev[i] < 0 || prev[i] >= 8194) return -1;
  const int s = tts_host::sample_sparse_one(r->r, vals, idx, n, prev, n_prev, logprob);
  if (s < 0) return 1;  // the full row is needed (nothing was drawn from the generator)
  *sample_out = s;
  return 0;
}
int tts_host_sample_reference_order(tts_rng *r, const float *logits, const int32_t *prev, int n_prev, int B,
                                    int32_t *out) {
  for (int b = 0; b < B; ++b)
    out[b] = tts_host::sample_literal_one(r->r, logits + size_t(b) * 8194, prev + size_t(b) * n_prev, n_prev);
  return 0;
}

int tts_host_apply_padding(const int32_t *seq, int n, int32_t *out502) {
  if (n < 0 || n > 500) return -5;
  std::vector<int32_t> v(seq, seq + n);
  tts_host::apply_padding(v);
  if (v.size() != 502) return -5;  // only possible when trailing 8139s were stripped from a full sequence
  memcpy(out502, v.data(), 502 * 4);
  return 0;
}
int tts_host_trim_count(const int32_t *codes500) { return tts_host::trim_count(codes500); }

int tts_host_write_wav(const char *path, const float *data, int64_t n, int rate) {
  return tts_host::write_wav(path, data, n, rate) ? 0 : -2;
}

int tts_host_timestep_map(int n_steps, int32_t *out) {
  const std::vector<int> m = tts_host::timestep_map(n_steps);
  for (size_t i = 0; i < m.size(); ++i) out[i] = m[i];
  return int(m.size());
}
void tts_host_timestep_embedding(int t, float *out) { tts_host::timestep_embedding(t, out); }
void tts_host_relative_position_buckets(int n, int32_t *out) {
  const std::vector<int> b = tts_host::relative_position_buckets(n);
  memcpy(out, b.data(), b.size() * 4);
}
int tts_host_ddpm_schedule(int n_steps, float *out) {
  const auto s = tts_host::ddpm_schedule(n_steps);
  for (size_t i = 0; i < s.size(); ++i) {
    float *o = out + i * 9;
    o[0] = s[i].cfk; o[1] = s[i].sqrt_recip; o[2] = s[i].sqrt_recipm1; o[3] = s[i].coef1; o[4] = s[i].coef2;
    o[5] = s[i].min_log; o[6] = s[i].max_log; o[7] = float(s[i].last); o[8] = float(s[i].timestep);
  }
  return int(s.size());
}

}  // extern "C"
