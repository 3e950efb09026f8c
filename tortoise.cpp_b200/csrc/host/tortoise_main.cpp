// tortoise_main.cpp -- the drop-in `tortoise` executable: same flags, defaults, relative
// model paths and output format as the reference CLI (main.cpp:6528-6584).  Flags are parsed
// pairwise left to right and the last argv element is never treated as a flag (`i < argc-1`).
// Extra, non-reference flags (all default to reference behaviour): --candidates N,
// --steps N, --dtype f32|f16, --models DIR, --device N, --bench-json, --normalize (spell out
// numbers / symbols and lower-case the message before tokenisation), --gpus G (shard the
// candidates over G GPUs of this box: one context + one host thread per GPU, weights replicated,
// ONE NCCL all-gather of (score, length) per candidate for the selection -- tts_gather_select --
// then latent pass + diffusion + vocoder for the winner on the GPU that owns it; the communicators are
// built in a background thread while the models load and the candidates decode).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <thread>
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
  int candidates = 1, steps = 80, dtype = TTS_DTYPE_F32, device = 0, gpus = 1;
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
    else if (a == "--gpus") gpus = std::stoi(argv[i + 1]);
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

  if (gpus < 1) gpus = 1;
  if (candidates < gpus) candidates = gpus;
  const int per = (candidates + gpus - 1) / gpus;  // candidates per GPU (the last GPU may idle a few)
  const int B = per;
  std::vector<tts_ctx *> ctxs(gpus, nullptr);
  const double t0 = now_s();
  std::vector<int> load_rc(gpus, TTS_OK);
  for (int g = 0; g < gpus; ++g) {
    tts_config cfg{};
    cfg.device = device + g;
    cfg.dtype = dtype;
    cfg.max_batch = per > 4 ? per : 4;
    cfg.max_positions = 404 + (per > 4 ? 256 : 0);
    cfg.parity_quirks = 1;
    load_rc[g] = tts_init(&cfg, &ctxs[g]);
    if (load_rc[g] != TTS_OK) {
      fprintf(stderr, "GPU %d: %s\n", device + g, tts_last_error(nullptr));
      return 1;
    }
  }
  // the NCCL communicators of the selection step are built in the background while the models load and the
  // candidates decode: ncclCommInitAll over 8 GPUs takes tens of seconds on some hosts (topology probing), the
  // decode loop of 8 candidates per GPU well under one
  tts_group *grp = nullptr;
  int grp_rc = TTS_OK;
  double grp_s = 0.0;
  std::thread grp_thread;
  if (gpus > 1)
    grp_thread = std::thread([&] {
      const double g0 = now_s();
      grp_rc = tts_group_init_local(ctxs.data(), gpus, &grp);
      grp_s = now_s() - g0;
    });
  struct Joiner {  // every early return below leaves through here
    std::thread &t;
    ~Joiner() { if (t.joinable()) t.join(); }
  } joiner{grp_thread};
  auto load_one = [&](int g) {
    int rc = tts_load_ar(ctxs[g], (models + "/ggml-model.bin").c_str());
    if (rc == TTS_OK) rc = tts_load_diffusion(ctxs[g], (models + "/ggml-diffusion-model.bin").c_str());
    if (rc == TTS_OK) rc = tts_load_vocoder(ctxs[g], (models + "/ggml-vocoder-model.bin").c_str());
    load_rc[g] = rc;
  };
  {
    std::vector<std::thread> th;
    for (int g = 1; g < gpus; ++g) th.emplace_back(load_one, g);
    load_one(0);
    for (auto &t : th) t.join();
  }
  for (int g = 0; g < gpus; ++g)
    if (load_rc[g] != TTS_OK) {
      fprintf(stderr, "GPU %d: %s\n", device + g, ctxs[g] ? tts_last_error(ctxs[g]) : tts_last_error(nullptr));
      if (grp_thread.joinable()) grp_thread.join();
      return 1;
    }
  const double t1 = now_s();
  // AR stage: every GPU decodes its own candidates (GPU g draws from generator(seed + g); g = 0 is the
  // reference's stream, so --gpus 1 is bit-identical to the single-GPU run)
  std::vector<std::vector<int32_t>> codes(gpus, std::vector<int32_t>(size_t(B) * 500)), nlat(gpus, std::vector<int32_t>(B));
  std::vector<std::vector<float>> score(gpus, std::vector<float>(B));
  std::vector<int32_t> ar_steps_g(gpus, 0);
  std::vector<int> ar_rc(gpus, TTS_OK);
  std::vector<tts_rng *> rngs(gpus, nullptr);
  rngs[0] = rng;
  for (int g = 1; g < gpus; ++g) rngs[g] = tts_rng_create(seed + uint32_t(g));
  std::vector<float> latents(size_t(gpus > 1 ? 1 : B) * 500 * 1024);
  auto ar_one = [&](int g) {
    tts_ar_options opt{};
    opt.per_candidate_stop = B > 4 ? 1 : 0;
    opt.skip_latents = gpus > 1 ? 1 : 0;  // multi-GPU: select first, latent pass for the winner only
    ar_rc[g] = tts_host_autoregressive(ctxs[g], rngs[g], tokens.data(), T, voice.data(), B, &opt, codes[g].data(),
                                       gpus > 1 ? nullptr : latents.data(), nlat[g].data(), score[g].data(), &ar_steps_g[g]);
  };
  {
    std::vector<std::thread> th;
    for (int g = 1; g < gpus; ++g) th.emplace_back(ar_one, g);
    ar_one(0);
    for (auto &t : th) t.join();
  }
  for (int g = 0; g < gpus; ++g)
    if (ar_rc[g] != TTS_OK) { fprintf(stderr, "autoregressive (GPU %d): %s\n", device + g, tts_last_error(ctxs[g])); return 1; }
  int32_t ar_steps = 0;
  for (int g = 0; g < gpus; ++g) ar_steps = ar_steps_g[g] > ar_steps ? ar_steps_g[g] : ar_steps;
  printf("tokens sampled: %d\n", ar_steps);
  // selection.  The reference diffuses candidate 0 (main.cpp:6575); with more candidates the best mean
  // log-probability wins (extension; identical for one candidate).  Across GPUs: one NCCL all-gather.
  int owner = 0, best = 0;
  const double t_ar = now_s();  // the decode loops of every GPU are done
  double wait_grp_s = 0.0;
  if (gpus > 1) {
    grp_thread.join();
    wait_grp_s = now_s() - t_ar;
    if (grp_rc != TTS_OK) { fprintf(stderr, "NCCL group init failed\n"); return 1; }
    std::vector<float> sc(size_t(gpus) * B);
    std::vector<int32_t> ln(size_t(gpus) * B);
    for (int g = 0; g < gpus; ++g)
      for (int b = 0; b < B; ++b) {
        const bool live = g * B + b < candidates;  // padding candidates of the last GPU never win
        sc[size_t(g) * B + b] = live ? score[g][b] : -1e30f;
        ln[size_t(g) * B + b] = nlat[g][b];
      }
    int32_t winner = 0;
    if (tts_gather_select(grp, sc.data(), ln.data(), B, &winner, nullptr, nullptr) != TTS_OK) {
      fprintf(stderr, "gather/select: %s\n", tts_group_last_error(grp));
      return 1;
    }
    tts_group_free(grp);
    owner = winner / B;
    best = winner % B;
    int32_t n = 0;
    if (tts_host_latents(ctxs[owner], tokens.data(), T, voice.data(), codes[owner].data() + size_t(best) * 500, latents.data(), &n) != TTS_OK) {
      fprintf(stderr, "latents: %s\n", tts_last_error(ctxs[owner]));
      return 1;
    }
    printf("selected candidate %d of %d (GPU %d)\n", winner, candidates, device + owner);
  } else {
    for (int b = 1; b < B; ++b)
      if (score[0][b] > score[0][best]) best = b;
  }
  tts_ctx *ctx = ctxs[owner];
  rng = rngs[owner];
  const float *win_latents = latents.data() + (gpus > 1 ? 0 : size_t(best) * 500 * 1024);
  const double t2 = now_s();
  const int L = nlat[owner][best];
  int32_t S = 0;
  std::vector<float> mel(size_t(100) * (L * 4 * 24000 / 22050));
  int rc = tts_host_diffusion(ctx, rng, win_latents, L, steps, mel.data(), &S);
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
    // ar_s: the decode loops (all GPUs in parallel); select_s: all-gather + argmax + the winner's latent pass;
    // nccl_init_s: communicator construction (runs beside load + decode), nccl_wait_s: what of it was NOT hidden
    printf("{\"load_s\": %.3f, \"ar_s\": %.4f, \"select_s\": %.4f, \"nccl_init_s\": %.3f, \"nccl_wait_s\": %.3f, "
           "\"diffusion_s\": %.4f, \"vocoder_s\": %.4f, \"audio_s\": %.3f, "
           "\"rtf\": %.3f, \"ar_steps\": %d, \"candidates\": %d, \"winner\": %d, \"launches\": %lld, \"codes\": [",
           t1 - t0, t_ar - t1, t2 - t_ar - wait_grp_s, grp_s, wait_grp_s, t3 - t2, t4 - t3, audio_s,
           audio_s / (t4 - t1 - wait_grp_s), ar_steps, candidates, owner * B + best, (long long)tts_launch_count(ctx));
    // sampled mel codes of the diffused candidate up to and including the stop token (parity tests)
    for (int i = 0; i < 500; ++i) {
      const int code = codes[owner][size_t(best) * 500 + i];
      printf("%s%d", i ? ", " : "", code);
      if (code == TTS_MEL_STOP) break;
    }
    printf("]}\n");
  }
  for (int g = 0; g < gpus; ++g) {
    tts_free(ctxs[g]);
    tts_rng_free(rngs[g]);
  }
  return 0;
}
