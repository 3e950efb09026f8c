// host_math.h -- pure-C++ restatement of the reference's host-side arithmetic that feeds
// the device path (no CUDA here; linked into both libtortoise_host.so and libtortoise_b200.so).
#pragma once
#include <cstdint>
#include <vector>

namespace tts_host {

struct DdpmStep {  // coefficients of ONE sampling step, already narrowed to float like main.cpp:5988-6013
  float cfk, sqrt_recip, sqrt_recipm1, coef1, coef2, min_log, max_log;
  int last;
  int timestep;  // original-schedule timestep fed to the time embedding
};

// timestep_map of the respaced schedule: literal table for 80 (main.cpp:5641-5648) which
// equals guided-diffusion's accumulating rule; the rule is used for other step counts.
std::vector<int> timestep_map(int n_steps);

// The reference's DDPM schedule (main.cpp:5370-5494, 5650-5716) for n_steps sampling steps,
// returned in SAMPLING ORDER: element i is diffusion_index i (timestep index n_steps-1-i).
std::vector<DdpmStep> ddpm_schedule(int n_steps);

// generate_timestep_embedding (main.cpp:5496-5521): [cos | sin] of float(t) * freq_k
void timestep_embedding(int t, float *out1024);

// rp -> bucket part of get_relative_position_buckets (main.cpp:4722-4749), rp in [0, n)
std::vector<int> relative_position_table(int n);
// full table [n*n], row = i (query), col = c (key): exactly what the reference uploads
std::vector<int> relative_position_buckets(int n);

// nearest-neighbour source index of ggml_upscale_ext (ggml.c:15527-15568) for L -> S
std::vector<int> upscale_index(int L, int S);

}  // namespace tts_host
