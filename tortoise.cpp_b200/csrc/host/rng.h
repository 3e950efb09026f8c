// rng.h -- the reference's three RNG globals (main.cpp:35-50) bundled per stream.
#pragma once
#include <cstdint>
#include <random>
namespace tts_host {
struct Rng {
  std::mt19937 generator;
  std::uniform_real_distribution<float> distribution{0.0, 1.0};
  std::normal_distribution<double> normal{0.0, 1.0};
  explicit Rng(uint32_t seed) : generator(seed) {}
};
}  // namespace tts_host

// opaque handle of the C-ABI (include/tortoise_host.h)
struct tts_rng {
  tts_host::Rng r;
  explicit tts_rng(uint32_t s) : r(s) {}
};
