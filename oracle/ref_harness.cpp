// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Drives the UNMODIFIED reference (balisujohn/tortoise.cpp main.cpp + its vendored ggml,
// compiled from /root/reference by oracle/Makefile into oracle/_ref/ref_harness) stage
// by stage and dumps the tensors our CUDA path is compared against.  No reference source
// is copied into this repo: main.cpp is #include'd from where it lies, with `main`
// renamed, and two ggml entry points are intercepted by macro so that
//   * every ggml_backend_graph_compute (main.cpp:5186,5247,5342,5838,5955,6112) is timed
//     and counted (lets bench.py bound a CPU sample to N graph runs), and
//   * every ggml_backend_tensor_get of a graph result (logits main.cpp:4767, diffusion
//     "output" main.cpp:5840/5960, latents 5352, audio 6122) can be dumped to disk.
//
// Usage (cwd must contain ../models/{tokenizer.json,ggml-*.bin} exactly like ./tortoise):
//   ref_harness shapes  <out.json>                    tensor name/shape manifest of the 3 files
//   ref_harness tokenize <text>                        prints "255,<ids>,0"
//   ref_harness ar   <text> <voice.bin> <B> <seed> <outdir> [max_computes|-1] [force_len]
//   ref_harness diff <latents.f32> <seed> <outdir> [max_computes]
//   ref_harness voc  <mel.f32> <seed> <outdir>
//   ref_harness full <text> <voice.bin> <seed> <outdir>   (== ./tortoise, with dumps + timings)
#include "ggml/ggml-alloc.h"
#include "ggml/ggml-backend.h"
#include "ggml/ggml.h"

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <sys/stat.h>
#include <vector>

namespace hx {
static std::string g_outdir;          // empty => no dumps
static std::string g_stage;           // "ar" | "diff" | "voc"
static int g_compute_count = 0;       // graph computes so far in this stage
static int g_max_computes = -1;       // exit(0) after this many (bounded CPU sample)
static int g_get_count = 0;
static double g_compute_seconds = 0;
static std::vector<double> g_compute_times;
static bool g_dump_all_logits = true;
// forced-length AR runs (long-context goldens): while fewer than g_force_len logit rows have been
// read back, the stop token's logit (8193) is replaced by -1e30 IN TRANSIT, after the true row was
// dumped -- the reference's code is untouched, its sampler just never sees a winning stop logit.
// Mirrors tts_ar_options.forced_codes of our own stage driver.
static int g_force_len = 0;
static int g_batch = 1;

static void write_file(const std::string &path, const void *p, size_t n) {
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) { fprintf(stderr, "harness: cannot write %s\n", path.c_str()); exit(2); }
  fwrite(p, 1, n, f);
  fclose(f);
}

static void report_and_exit_if_capped() {
  if (g_max_computes >= 0 && g_compute_count >= g_max_computes) {
    printf("\nHARNESS_CAPPED stage=%s computes=%d compute_seconds=%.6f\n", g_stage.c_str(),
           g_compute_count, g_compute_seconds);
    printf("HARNESS_COMPUTE_TIMES");
    for (double t : g_compute_times) printf(" %.6f", t);
    printf("\n");
    fflush(stdout);
    exit(0);
  }
}

static enum ggml_status graph_compute(ggml_backend_t backend, struct ggml_cgraph *gf) {
  auto t0 = std::chrono::steady_clock::now();
  enum ggml_status st = ggml_backend_graph_compute(backend, gf);
  auto t1 = std::chrono::steady_clock::now();
  double dt = std::chrono::duration<double>(t1 - t0).count();
  g_compute_seconds += dt;
  g_compute_times.push_back(dt);
  g_compute_count++;
  return st;
}

static void tensor_get(const struct ggml_tensor *t, void *data, size_t offset, size_t size) {
  ggml_backend_tensor_get(t, data, offset, size);
  if (!g_outdir.empty()) {
    char name[256];
    // name by stage + running compute index so step-wise comparison is possible
    snprintf(name, sizeof name, "%s/%s_get%04d_c%04d.f32", g_outdir.c_str(), g_stage.c_str(),
             g_get_count, g_compute_count);
    bool dump = true;
    if (g_stage == "ar" && !g_dump_all_logits && g_get_count > 4) dump = false;
    if (dump) write_file(name, data, size);
  }
  if (g_stage == "ar" && g_force_len > 0 && size == size_t(g_batch) * 8194 * 4 && g_get_count < g_force_len)
    for (int b = 0; b < g_batch; ++b) static_cast<float *>(data)[size_t(b) * 8194 + 8193] = -1e30f;
  g_get_count++;
  report_and_exit_if_capped();
}
// inputs the reference uploads per graph run (x_t of each diffusion pass main.cpp:5808/5925,
// vocoder noise main.cpp:6105) are dumped so single passes can be checked teacher-forced.
static int g_set_count = 0;
static void tensor_set(struct ggml_tensor *t, const void *data, size_t offset, size_t size) {
  ggml_backend_tensor_set(t, data, offset, size);
  if (!g_outdir.empty() && (std::string(t->name) == "noise_tensor" ||
                            std::string(t->name) == "vocoder_noise_tensor")) {
    char name[256];
    snprintf(name, sizeof name, "%s/%s_set_%s_c%04d.f32", g_outdir.c_str(), g_stage.c_str(), t->name,
             g_compute_count);
    write_file(name, data, size);
  }
  g_set_count++;
}
} // namespace hx

#define ggml_backend_tensor_set(t, d, o, s) hx::tensor_set((t), (d), (o), (s))
#define ggml_backend_graph_compute(b, g) hx::graph_compute((b), (g))
#define ggml_backend_tensor_get(t, d, o, s) hx::tensor_get((t), (d), (o), (s))
#define main reference_main
#ifndef HX_MAIN
#define HX_MAIN "main.cpp" // resolved through -I/root/reference ; never copied
#endif
#include HX_MAIN // (-DHX_MAIN=... selects the "patched reference" TU of oracle/patch_steps.py)
#undef main
#undef ggml_backend_graph_compute
#undef ggml_backend_tensor_get
#undef ggml_backend_tensor_set

static void begin_stage(const char *s) {
  hx::g_stage = s;
  hx::g_compute_count = 0;
  hx::g_get_count = 0;
  hx::g_compute_seconds = 0;
  hx::g_compute_times.clear();
}

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static std::vector<float> read_f32(const std::string &p) {
  FILE *f = fopen(p.c_str(), "rb");
  if (!f) { fprintf(stderr, "harness: cannot read %s\n", p.c_str()); exit(2); }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<float> v(n / 4);
  if (fread(v.data(), 4, v.size(), f) != v.size()) { exit(2); }
  fclose(f);
  return v;
}

static std::vector<gpt_vocab::id> tokenize_like_main(std::string message) {
  gpt_vocab vocab;
  gpt_vocab_init("../models/tokenizer.json", vocab);
  replaceAll(message, " ", "[SPACE]");
  std::vector<gpt_vocab::id> tokens = ::gpt_tokenize(vocab, message);
  tokens.insert(tokens.begin(), 255);
  tokens.push_back(0);
  return tokens;
}

template <class M> static void dump_shapes(FILE *out, const char *file, M &model, bool last) {
  fprintf(out, "  \"%s\": [\n", file);
  size_t k = 0;
  for (auto &kv : model.tensors) {
    ggml_tensor *t = kv.second;
    fprintf(out, "    {\"name\": \"%s\", \"ne\": [", kv.first.c_str());
    int nd = ggml_n_dims(t);
    for (int i = 0; i < nd; i++) fprintf(out, "%s%d", i ? ", " : "", (int)t->ne[i]);
    fprintf(out, "]}%s\n", ++k == model.tensors.size() ? "" : ",");
  }
  fprintf(out, "  ]%s\n", last ? "" : ",");
}

static void write_magic_only(const char *path) {
  uint32_t magic = 0x67676d6c;
  hx::write_file(path, &magic, 4);
}

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: see header of oracle/ref_harness.cpp\n"); return 2; }
  std::string cmd = argv[1];
  setvbuf(stdout, NULL, _IOLBF, 0);

  if (cmd == "shapes") {
    // A file holding only the magic makes each loader build its full name->tensor map and
    // read zero records (main.cpp:811-888 loop ends at EOF), which we then enumerate.
    mkdir("/tmp/_hx_shapes", 0755);
    write_magic_only("/tmp/_hx_shapes/m.bin");
    FILE *out = fopen(argv[2], "w");
    fprintf(out, "{\n");
    { autoregressive_model m; if (!autoregressive_model_load("/tmp/_hx_shapes/m.bin", m)) return 1;
      dump_shapes(out, "ggml-model.bin", m, false); }
    { diffusion_model m; if (!diffusion_model_load("/tmp/_hx_shapes/m.bin", m)) return 1;
      dump_shapes(out, "ggml-diffusion-model.bin", m, false); }
    { vocoder_model m; if (!vocoder_model_load("/tmp/_hx_shapes/m.bin", m)) return 1;
      dump_shapes(out, "ggml-vocoder-model.bin", m, true); }
    fprintf(out, "}\n");
    fclose(out);
    return 0;
  }

  if (cmd == "tokenize") {
    auto toks = tokenize_like_main(argv[2]);
    for (size_t i = 0; i < toks.size(); i++) printf("%s%d", i ? "," : "", toks[i]);
    printf("\n");
    return 0;
  }

  if (cmd == "ar") {
    std::string text = argv[2], voice = argv[3];
    int B = atoi(argv[4]);
    generator.seed(atoi(argv[5]));
    hx::g_outdir = argv[6];
    if (argc > 7) hx::g_max_computes = atoi(argv[7]);
    if (argc > 8) hx::g_force_len = atoi(argv[8]);
    hx::g_batch = B;
    mkdir(hx::g_outdir.c_str(), 0755);
    auto tokens = tokenize_like_main(text);
    hx::write_file(hx::g_outdir + "/tokens.i32", tokens.data(), tokens.size() * 4);
    begin_stage("ar");
    double t0 = now_s();
    auto res = autoregressive(tokens, voice, B);
    double t1 = now_s();
    for (int b = 0; b < B; b++) {
      hx::write_file(hx::g_outdir + "/trimmed_latents_" + std::to_string(b) + ".f32",
                     res.first[b].data(), res.first[b].size() * 4);
      // NB: trim_latents() (main.cpp:4873) erased first/last: 500 codes remain
      hx::write_file(hx::g_outdir + "/codes_" + std::to_string(b) + ".i32", res.second[b].data(),
                     res.second[b].size() * 4);
    }
    printf("\nHARNESS_DONE stage=ar wall=%.3f computes=%d compute_seconds=%.6f\n", t1 - t0,
           hx::g_compute_count, hx::g_compute_seconds);
    printf("HARNESS_COMPUTE_TIMES");
    for (double t : hx::g_compute_times) printf(" %.6f", t);
    printf("\n");
    return 0;
  }

  if (cmd == "diff") {
    auto lat = read_f32(argv[2]);
    generator.seed(atoi(argv[3]));
    hx::g_outdir = argv[4];
    if (argc > 5) hx::g_max_computes = atoi(argv[5]);
    mkdir(hx::g_outdir.c_str(), 0755);
    begin_stage("diff");
    double t0 = now_s();
    auto mel = diffusion(lat);
    double t1 = now_s();
    hx::write_file(hx::g_outdir + "/mel.f32", mel.data(), mel.size() * 4);
    printf("\nHARNESS_DONE stage=diff wall=%.3f computes=%d compute_seconds=%.6f\n", t1 - t0,
           hx::g_compute_count, hx::g_compute_seconds);
    return 0;
  }

  if (cmd == "voc") {
    auto mel = read_f32(argv[2]);
    generator.seed(atoi(argv[3]));
    hx::g_outdir = argv[4];
    mkdir(hx::g_outdir.c_str(), 0755);
    begin_stage("voc");
    double t0 = now_s();
    auto audio = vocoder(mel);
    double t1 = now_s();
    hx::write_file(hx::g_outdir + "/audio.f32", audio.data(), audio.size() * 4);
    printf("\nHARNESS_DONE stage=voc wall=%.3f computes=%d compute_seconds=%.6f\n", t1 - t0,
           hx::g_compute_count, hx::g_compute_seconds);
    return 0;
  }

  if (cmd == "full") {
    std::string text = argv[2], voice = argv[3];
    generator.seed(atoi(argv[4]));
    hx::g_outdir = argv[5];
    hx::g_dump_all_logits = true;
    mkdir(hx::g_outdir.c_str(), 0755);
    auto tokens = tokenize_like_main(text);
    hx::write_file(hx::g_outdir + "/tokens.i32", tokens.data(), tokens.size() * 4);
    begin_stage("ar");
    double t0 = now_s();
    auto res = autoregressive(tokens, voice, 1);
    double t1 = now_s();
    double ar_compute = hx::g_compute_seconds; int ar_n = hx::g_compute_count;
    hx::write_file(hx::g_outdir + "/trimmed_latents_0.f32", res.first[0].data(), res.first[0].size() * 4);
    hx::write_file(hx::g_outdir + "/codes_0.i32", res.second[0].data(), res.second[0].size() * 4);
    begin_stage("diff");
    auto mel = diffusion(res.first[0]);
    double t2 = now_s();
    double df_compute = hx::g_compute_seconds;
    hx::write_file(hx::g_outdir + "/mel.f32", mel.data(), mel.size() * 4);
    begin_stage("voc");
    auto audio = vocoder(mel);
    double t3 = now_s();
    hx::write_file(hx::g_outdir + "/audio.f32", audio.data(), audio.size() * 4);
    writeWav((hx::g_outdir + "/output.wav").c_str(), audio, 24000);
    printf("\nHARNESS_DONE stage=full ar_wall=%.3f diff_wall=%.3f voc_wall=%.3f ar_computes=%d "
           "ar_compute_s=%.3f diff_compute_s=%.3f voc_compute_s=%.3f samples=%zu\n",
           t1 - t0, t2 - t1, t3 - t2, ar_n, ar_compute, df_compute, hx::g_compute_seconds,
           audio.size());
    return 0;
  }

  if (cmd == "sample") {
    // ref_harness sample <logits.f32 [B*8194]> <prev_inputs.i32 [B*n]> <B> <seed> <n_draws> <out.i32>
    // Runs the reference's own process_logits_and_sample (main.cpp:4753) n_draws times on
    // the same logits (RNG advancing), through a one-node CPU graph holding the logits.
    auto logits = read_f32(argv[2]);
    std::vector<int> prev;
    { auto raw = read_f32(argv[3]); prev.resize(raw.size()); memcpy(prev.data(), raw.data(), raw.size() * 4); }
    int B = atoi(argv[4]);
    generator.seed(atoi(argv[5]));
    int n_draws = atoi(argv[6]);
    ggml_backend_t backend = ggml_backend_cpu_init();
    struct ggml_init_params ip = {ggml_tensor_overhead() * 8 + ggml_graph_overhead(), NULL, true};
    struct ggml_context *ctx = ggml_init(ip);
    struct ggml_tensor *t = ggml_new_tensor_2d(ctx, GGML_TYPE_F32, 8194, B);
    ggml_backend_buffer_t buf = ggml_backend_alloc_ctx_tensors(ctx, backend);
    ggml_backend_tensor_set(t, logits.data(), 0, logits.size() * 4);
    struct ggml_cgraph *gf = ggml_new_graph(ctx);
    gf->nodes[0] = t;
    gf->n_nodes = 1;
    std::vector<int> out;
    for (int d = 0; d < n_draws; d++) {
      std::vector<int> sm = process_logits_and_sample(gf, prev, d, B);
      out.insert(out.end(), sm.begin(), sm.end());
    }
    hx::write_file(argv[7], out.data(), out.size() * 4);
    (void)buf;
    return 0;
  }

  if (cmd == "hostfn") {
    // ref_harness hostfn <outdir>: golden vectors of the reference's pure host functions
    std::string od = argv[2];
    mkdir(od.c_str(), 0755);
    std::vector<int> tm = {0, 51, 101, 152, 1012, 2025, 3037, 3948, 3999};
    std::vector<float> te;
    for (int t : tm) { auto e = generate_timestep_embedding({t}, 1024, 10000); te.insert(te.end(), e.begin(), e.end()); }
    hx::write_file(od + "/timestep_embeddings.f32", te.data(), te.size() * 4);
    hx::write_file(od + "/timestep_values.i32", tm.data(), tm.size() * 4);
    for (int n : {26, 113, 300}) {
      auto b = get_relative_position_buckets(n);
      hx::write_file(od + "/buckets_" + std::to_string(n) + ".i32", b.data(), b.size() * 4);
    }
    std::vector<int> seq = {5, 6, 7, 8139, 83, 83, 8193};
    apply_padding(seq);
    hx::write_file(od + "/apply_padding_a.i32", seq.data(), seq.size() * 4);
    std::vector<int> seq2 = {100, 200, 8139, 8139};
    apply_padding(seq2);
    hx::write_file(od + "/apply_padding_b.i32", seq2.data(), seq2.size() * 4);
    std::vector<float> mel = {-1.0f, -0.5f, 0.0f, 0.25f, 1.0f};
    denormalize_tacotron_mel(mel);
    hx::write_file(od + "/denorm_mel.f32", mel.data(), mel.size() * 4);
    generator.seed(0);
    auto nn = sample_normal_noise(1000);
    hx::write_file(od + "/normal_seed0_1000.f32", nn.data(), nn.size() * 4);
    std::vector<float> uu(1000);
    generator.seed(7);
    for (auto &u : uu) u = distribution(generator);
    hx::write_file(od + "/uniform_seed7_1000.f32", uu.data(), uu.size() * 4);
    std::vector<float> wav = {0.0f, 0.5f, -0.5f, 1.0f};
    writeWav((od + "/tiny.wav").c_str(), wav, 24000);
    return 0;
  }

  fprintf(stderr, "unknown command %s\n", cmd.c_str());
  return 2;
}
