// loader.cu -- container parsing + upload helpers shared by the three model loaders.
// Format contract: reference parsers autoregressive_model_load (main.cpp:482-897),
// diffusion_model_load (931-1634), vocoder_model_load (1665-2021); SURVEY.md App. B.
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "engine.h"

namespace tts {

bool Container::open(const std::string &p, std::string &err) {
  path = p;
  tensors.clear();
  order.clear();
  FILE *f = fopen(p.c_str(), "rb");
  if (!f) {
    err = "cannot open '" + p + "'";
    return false;
  }
  uint32_t magic = 0;
  if (fread(&magic, 4, 1, f) != 1 || magic != 0x67676d6c) {
    err = "bad magic in '" + p + "' (expected 0x67676d6c, main.cpp:494-500)";
    fclose(f);
    return false;
  }
  fseek(f, 0, SEEK_END);
  const long fsize = ftell(f);
  fseek(f, 4, SEEK_SET);
  while (true) {
    int32_t hdr[3];
    if (fread(hdr, 4, 3, f) != 3) break;  // EOF ends the record list (main.cpp:821-823)
    const int n_dims = hdr[0], name_len = hdr[1], ttype = hdr[2];
    if (n_dims < 1 || n_dims > 4 || name_len <= 0 || name_len > 512) {
      err = "malformed record header in '" + p + "'";
      fclose(f);
      return false;
    }
    if (ttype != 0) {
      err = "only f32 tensors (ttype 0) are supported, as in the reference (main.cpp:682-729)";
      fclose(f);
      return false;
    }
    HostTensor t;
    t.ne.resize(n_dims);
    if (fread(t.ne.data(), 4, n_dims, f) != size_t(n_dims)) { err = "truncated record"; fclose(f); return false; }
    std::string name(name_len, 0);
    if (fread(&name[0], 1, name_len, f) != size_t(name_len)) { err = "truncated record"; fclose(f); return false; }
    t.nelem = 1;
    bool bad_dim = false;
    for (int d : t.ne) {
      // a zero / negative dim, or a product that cannot fit the file, is a malformed record (a
      // negative dim would otherwise wrap the size_t product past the bounds check below)
      if (d <= 0 || t.nelem > size_t(fsize) / size_t(d)) { bad_dim = true; break; }
      t.nelem *= size_t(d);
    }
    if (bad_dim) {
      err = "tensor '" + name + "' has an invalid shape in '" + p + "'";
      fclose(f);
      return false;
    }
    t.offset = size_t(ftell(f));
    if (t.nelem > (size_t(fsize) - t.offset) / 4) {
      err = "tensor '" + name + "' runs past the end of '" + p + "'";
      fclose(f);
      return false;
    }
    fseek(f, long(t.nelem * 4), SEEK_CUR);
    tensors[name] = t;
    order.push_back(name);
  }
  fclose(f);
  return true;
}

static void ensure_staging(tts_ctx *c, size_t bytes) {
  if (c->staging_bytes < bytes) {
    if (c->staging) ctx_free_host(c, c->staging);
    TTS_CUDA_TRY(ctx_malloc_host(c, &c->staging, bytes));
    c->staging_bytes = bytes;
  }
  if (c->d_scratch_bytes < bytes) {
    if (c->d_scratch) ctx_free(c, c->d_scratch);
    TTS_CUDA_TRY(ctx_malloc(c, &c->d_scratch, bytes));
    c->d_scratch_bytes = bytes;
  }
}

// Reads tensor `name` into the pinned staging buffer AND the device scratch buffer.
void read_tensor_to_staging(tts_ctx *c, const Container &ct, const std::string &name, size_t *nelem) {
  auto it = ct.tensors.find(name);
  if (it == ct.tensors.end()) throw ArgError("tensor '" + name + "' missing from " + ct.path, TTS_EIO);
  const HostTensor &t = it->second;
  ensure_staging(c, t.nelem * 4);
  FILE *f = fopen(ct.path.c_str(), "rb");
  if (!f) throw ArgError("cannot reopen " + ct.path, TTS_EIO);
  fseek(f, long(t.offset), SEEK_SET);
  const size_t got = fread(c->staging, 4, t.nelem, f);
  fclose(f);
  if (got != t.nelem) throw ArgError("short read of '" + name + "'", TTS_EIO);
  TTS_CUDA_TRY(cudaMemcpyAsync(c->d_scratch, c->staging, t.nelem * 4, cudaMemcpyHostToDevice, c->stream));
  TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (nelem) *nelem = t.nelem;
}

static bool shape_matches(const std::vector<int> &got, const std::vector<int> &want) {
  // the reference checks element count and ne[0], ne[1] only (main.cpp:842-855); trailing
  // 1-dims may be present or absent in the file.
  size_t ng = 1, nw = 1;
  for (int d : got) ng *= size_t(d);
  for (int d : want) nw *= size_t(d);
  if (ng != nw) return false;
  for (size_t i = 0; i < 2; ++i) {
    const int g = i < got.size() ? got[i] : 1, w = i < want.size() ? want[i] : 1;
    if (g != w) return false;
  }
  return true;
}

// Uploads an f32 tensor unchanged; returns a fresh device allocation.
float *upload_f32(tts_ctx *c, const Container &ct, const std::string &name, const std::vector<int> &expect_ne) {
  auto it = ct.tensors.find(name);
  if (it == ct.tensors.end()) throw ArgError("tensor '" + name + "' missing from " + ct.path, TTS_EIO);
  if (!shape_matches(it->second.ne, expect_ne))
    throw ArgError("tensor '" + name + "' has wrong shape in model file", TTS_EIO);
  size_t n = 0;
  read_tensor_to_staging(c, ct, name, &n);
  float *d = nullptr;
  TTS_CUDA_TRY(ctx_malloc(c, &d, n * 4));
  TTS_CUDA_TRY(cudaMemcpyAsync(d, c->d_scratch, n * 4, cudaMemcpyDeviceToDevice, c->stream));
  TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
  return d;
}

}  // namespace tts
