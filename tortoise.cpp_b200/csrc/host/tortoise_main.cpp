// tortoise_main.cpp -- the drop-in `tortoise` executable: same flags, defaults, relative
// model paths and output format as the reference CLI (main.cpp:6528-6584).  Flags are parsed
// pairwise left to right and the last argv element is never treated as a flag (`i < argc-1`).
// Extra, non-reference flags (all default to reference behaviour): --candidates N,
// --steps N, --dtype f32|f16, --models DIR, --device N, --bench-json, --normalize (spell out
// numbers / symbols and lower-case the message before tokenisation).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "../../../include/tortoise_b200.h"
#include "../../../include/tortoise_host.h"

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv) {
  std::string message = "this is a test message.";
  std::string voicePath = "../models/mol.bin";
  std::string outputPath = "./output.wav";
  std::string models = "../models";
  bool seeded = false;
  uint32_t seed = 0;
  int candidates = 1, steps = 80, dtype = TTS_DTYPE_F32, device = 0;
  bool bench_json = false;
  for (int i = 1; i < argc - 1; ++i) {
    const std::string a = argv[i];
    if (a == "--voice") voicePath = argv[i + 1];
    else if (a == "--message") message = argv[i + 1];
    else if (a == "--output") outputPath = argv[i + 1];
    else if (a == "--seed") { seed = uint32_t(std::stoi(argv[i + 1])); seeded = true; }
    else if (a == "--candidates") candidates = std::stoi(argv[i + 1]);
    else if (a == "--steps") steps = std::stoi(argv[i + 1]);
    else if (a == "--dtype") dtype = std::string(argv[i + 1]) == "f16" ? TTS_DTYPE_F16 : TTS_DTYPE_F32;
    else if (a == "--models") models = argv[i + 1];
    else if (a == "--device") device = std::stoi(argv[i + 1]);
  }
  bool normalize = false;
  for (int i = 1; i < argc; ++i) {
    if (std::string(argv[i]) == "--bench-json") bench_json = true;
    if (std::string(argv[i]) == "--normalize") normalize = true;
  }
  if (normalize) {
    std::vector<char> buf(16 * message.size() + 64);
    if (tts_host_normalize_text(message.c_str(), buf.data(), int(buf.size())) >= 0) message = buf.data();
  }
  if (!seeded) {  // wall-clock milliseconds, like the reference's global initialiser (main.cpp:39-47)
    seed = unsigned(std::chrono::duration_cast<std::chrono::milliseconds>(
                        std::chrono::system_clock::now().time_since_epoch()).count());
  }
  tts_rng *rng = tts_rng_create(seed);

  std::vector<int32_t> tokens(2048);
  const int T = tts_host_tokenize((models + "/tokenizer.json").c_str(), message.c_str(), tokens.data(), 2048);
  if (T < 0) { fprintf(stderr, "failed to load %s/tokenizer.json\n", models.c_str()); return 1; }
  tokens.resize(T);

  std::vector<float> voice(1024, 0.f);
  {
    std::ifstream vf(voicePath, std::ios::binary);
    if (!vf.is_open()) { fprintf(stderr, "Error: Unable to open file %s\n", voicePath.c_str()); }
    else vf.read(reinterpret_cast<char *>(voice.data()), 1024 * sizeof(float));
  }

  tts_config cfg{};
  cfg.device = device;
  cfg.dtype = dtype;
  cfg.max_batch = candidates > 4 ? candidates : 4;
  cfg.max_positions = 404 + (candidates > 4 ? 256 : 0);
  cfg.parity_quirks = 1;
  tts_ctx *ctx = nullptr;
  if (tts_init(&cfg, &ctx) != TTS_OK) { fprintf(stderr, "tts_init: %s\n", tts_last_error(nullptr)); return 1; }
  const double t0 = now_s();
  if (tts_load_ar(ctx, (models + "/ggml-model.bin").c_str()) != TTS_OK ||
      tts_load_diffusion(ctx, (models + "/ggml-diffusion-model.bin").c_str()) != TTS_OK ||
      tts_load_vocoder(ctx, (models + "/ggml-vocoder-model.bin").c_str()) != TTS_OK) {
    fprintf(stderr, "failed to load model: %s\n", tts_last_error(ctx));
    return 1;
  }
  const double t1 = now_s();
  const int B = candidates;
  std::vector<int32_t> codes(size_t(B) * 500), nlat(B);
  std::vector<float> latents(size_t(B) * 500 * 1024), score(B);
  int32_t ar_steps = 0;
  tts_ar_options opt{};
  opt.per_candidate_stop = B > 4 ? 1 : 0;
  int rc = tts_host_autoregressive(ctx, rng, tokens.data(), T, voice.data(), B, &opt, codes.data(), latents.data(),
                                   nlat.data(), score.data(), &ar_steps);
  if (rc != TTS_OK) { fprintf(stderr, "autoregressive: %s\n", tts_last_error(ctx)); return 1; }
  printf("tokens sampled: %d\n", ar_steps);
  const double t2 = now_s();
  // the reference diffuses candidate 0 (main.cpp:6575); with --candidates > 1 the best mean
  // log-probability wins (extension; identical for one candidate)
  int best = 0;
  for (int b = 1; b < B; ++b)
    if (score[b] > score[best]) best = b;
  const int L = nlat[best];
  int32_t S = 0;
  std::vector<float> mel(size_t(100) * (L * 4 * 24000 / 22050));
  rc = tts_host_diffusion(ctx, rng, latents.data() + size_t(best) * 500 * 1024, L, steps, mel.data(), &S);
  if (rc != TTS_OK) { fprintf(stderr, "diffusion: %s\n", tts_last_error(ctx)); return 1; }
  const double t3 = now_s();
  std::vector<float> audio(size_t(S + 10) * 256 - 6);
  rc = tts_host_vocoder(ctx, rng, mel.data(), S, audio.data());
  if (rc != TTS_OK) { fprintf(stderr, "vocoder: %s\n", tts_last_error(ctx)); return 1; }
  const double t4 = now_s();
  if (tts_host_write_wav(outputPath.c_str(), audio.data(), (int64_t)audio.size(), 24000) != 0) {
    fprintf(stderr, "Error opening output file.\n");
    return 1;
  }
  printf("WAV file saved successfully. :^)\n");
  if (bench_json) {
    const double audio_s = audio.size() / 24000.0;
    printf("{\"load_s\": %.3f, \"ar_s\": %.4f, \"diffusion_s\": %.4f, \"vocoder_s\": %.4f, \"audio_s\": %.3f, "
           "\"rtf\": %.3f, \"ar_steps\": %d, \"candidates\": %d, \"winner\": %d, \"launches\": %lld, \"codes\": [",
           t1 - t0, t2 - t1, t3 - t2, t4 - t3, audio_s, audio_s / (t4 - t1), ar_steps, B, best,
           (long long)tts_launch_count(ctx));
    // sampled mel codes of the diffused candidate up to and including the stop token (parity tests)
    for (int i = 0; i < 500; ++i) {
      const int code = codes[size_t(best) * 500 + i];
      printf("%s%d", i ? ", " : "", code);
      if (code == TTS_MEL_STOP) break;
    }
    printf("]}\n");
  }
  tts_free(ctx);
  tts_rng_free(rng);
  return 0;
}
