// dist.cu -- the path's only exchange step: the final candidate gather / selection over NCCL
// (north star: "NCCL over NVLink only for the final candidate gather/selection").
//
// Candidates are sharded over GPUs with no data-path collective: every GPU decodes its own candidates
// against replicated weights.  At the end each rank contributes (score, n_codes) per candidate to ONE
// ncclAllGather; the argmax is taken on the host (first maximum wins, NaN scores lose) and the winner's
// owner alone runs latent pass + diffusion + vocoder -- no tensor moves.  The reference has no scorer
// and no multi-GPU path (it diffuses candidate 0, main.cpp:6575); with one rank and one candidate the
// result is candidate 0.
//
// Two ways to form the group: one process driving several GPUs (the C++ CLI, `tortoise --gpus N`:
// ncclCommInitAll over the contexts' devices) or one process per GPU (bench.py under torchrun:
// ncclCommInitRank with a unique id the launcher distributes).  libnccl.so.2 is opened with dlopen on
// first use, so the single-GPU library has no NCCL dependency.
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "common.cuh"
#include "engine.h"

namespace {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi &nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    auto sym = [&](const char *n) { return dlsym(h, n); };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.AllGather && api.GroupStart && api.GroupEnd &&
             api.CommDestroy && api.GetErrorString;
  });
  return api;
}

}  // namespace

struct tts_group {
  int world = 0;                    // ranks in the group
  int rank0 = 0;                    // global rank of local member 0
  std::vector<tts_ctx *> ctx;       // local members (all of them in single-process mode, one in rank mode)
  std::vector<ncclComm_t> comm;
  std::vector<int32_t *> d_send, d_recv;
  int cap = 0;                      // (score, len) pairs per rank the device buffers hold
  std::string err;
};

static int gfail(tts_group *g, int code, const std::string &msg) {
  if (g) g->err = msg;
  return code;
}

#define TTS_NCCL_TRY(g, expr)                                                            \
  do {                                                                                   \
    ncclResult_t _r = (expr);                                                            \
    if (_r != ncclSuccess) return gfail(g, TTS_ECUDA, std::string(#expr) + ": " + nccl().GetErrorString(_r)); \
  } while (0)

extern "C" {

int tts_nccl_unique_id(char *out128) {
  if (!out128) return TTS_EINVAL;
  if (!nccl().ok) return TTS_ENODEV;
  static_assert(sizeof(ncclUniqueId) == 128, "id is passed around as 128 opaque bytes");
  ncclUniqueId id;
  if (nccl().GetUniqueId(&id) != ncclSuccess) return TTS_ECUDA;
  memcpy(out128, &id, 128);
  return TTS_OK;
}

int tts_group_init_local(tts_ctx **ctxs, int n, tts_group **out) {
  if (!ctxs || n < 1 || !out) return TTS_EINVAL;
  *out = nullptr;
  if (!nccl().ok) return TTS_ENODEV;
  tts_group *g = new tts_group();
  g->world = n;
  g->rank0 = 0;
  std::vector<int> devs(n);
  for (int i = 0; i < n; ++i) {
    if (!ctxs[i]) { delete g; return TTS_EINVAL; }
    g->ctx.push_back(ctxs[i]);
    devs[i] = ctxs[i]->cfg.device;
  }
  g->comm.resize(n);
  if (nccl().CommInitAll(g->comm.data(), n, devs.data()) != ncclSuccess) { delete g; return TTS_ECUDA; }
  g->d_send.assign(n, nullptr);
  g->d_recv.assign(n, nullptr);
  // First collective = connection setup (seconds): pay it here, on streams of our own, so that a caller who builds
  // the group in the background (the CLI does, while models load) finds tts_gather_select warm.
  {
    std::vector<cudaStream_t> st(n, nullptr);
    std::vector<int32_t *> snd(n, nullptr), rcv(n, nullptr);
    bool ok = true;
    for (int i = 0; i < n && ok; ++i) {
      ok = cudaSetDevice(devs[i]) == cudaSuccess && cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking) == cudaSuccess &&
           cudaMalloc(&snd[i], 8) == cudaSuccess && cudaMalloc(&rcv[i], size_t(n) * 8) == cudaSuccess &&
           cudaMemsetAsync(snd[i], 0, 8, st[i]) == cudaSuccess;
    }
    if (ok) {
      ok = nccl().GroupStart() == ncclSuccess;
      for (int i = 0; i < n && ok; ++i) {
        cudaSetDevice(devs[i]);
        ok = nccl().AllGather(snd[i], rcv[i], 2, ncclInt32, g->comm[i], st[i]) == ncclSuccess;
      }
      ok = nccl().GroupEnd() == ncclSuccess && ok;
    }
    for (int i = 0; i < n; ++i) {
      cudaSetDevice(devs[i]);
      if (st[i]) { cudaStreamSynchronize(st[i]); cudaStreamDestroy(st[i]); }
      if (snd[i]) cudaFree(snd[i]);
      if (rcv[i]) cudaFree(rcv[i]);
    }
    if (!ok) { cudaGetLastError(); tts_group_free(g); return TTS_ECUDA; }
  }
  *out = g;
  return TTS_OK;
}

int tts_group_init_rank(tts_ctx *ctx, int rank, int world, const char *id128, tts_group **out) {
  if (!ctx || !id128 || !out || world < 1 || rank < 0 || rank >= world) return TTS_EINVAL;
  *out = nullptr;
  if (!nccl().ok) return TTS_ENODEV;
  if (cudaSetDevice(ctx->cfg.device) != cudaSuccess) return TTS_ECUDA;
  tts_group *g = new tts_group();
  g->world = world;
  g->rank0 = rank;
  g->ctx.push_back(ctx);
  g->comm.resize(1);
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  if (nccl().CommInitRank(&g->comm[0], world, id, rank) != ncclSuccess) { delete g; return TTS_ECUDA; }
  g->d_send.assign(1, nullptr);
  g->d_recv.assign(1, nullptr);
  *out = g;
  return TTS_OK;
}

const char *tts_group_last_error(const tts_group *g) { return g ? g->err.c_str() : ""; }

// scores / lens: [n_local][per] (n_local = local members: all ranks in single-process mode, 1 in rank mode).
// Outputs (any may be NULL): *winner = global candidate index rank * per + i of the best score,
// scores_all / lens_all [world][per].
int tts_gather_select(tts_group *g, const float *scores, const int32_t *lens, int per, int32_t *winner, float *scores_all,
                      int32_t *lens_all) {
  if (!g || !scores || !lens || per < 1) return TTS_EINVAL;
  const int nl = int(g->ctx.size());
  if (per > g->cap) {
    for (int i = 0; i < nl; ++i) {
      tts_ctx *c = g->ctx[i];
      if (cudaSetDevice(c->cfg.device) != cudaSuccess) return gfail(g, TTS_ECUDA, "cudaSetDevice failed");
      if (g->d_send[i]) tts::ctx_free(c, g->d_send[i]);
      if (g->d_recv[i]) tts::ctx_free(c, g->d_recv[i]);
      if (tts::ctx_malloc(c, &g->d_send[i], size_t(per) * 2 * 4) != cudaSuccess ||
          tts::ctx_malloc(c, &g->d_recv[i], size_t(g->world) * per * 2 * 4) != cudaSuccess)
        return gfail(g, TTS_ECUDA, "device allocation failed");
    }
    g->cap = per;
  }
  std::vector<int32_t> send(size_t(per) * 2);
  for (int i = 0; i < nl; ++i) {
    tts_ctx *c = g->ctx[i];
    if (cudaSetDevice(c->cfg.device) != cudaSuccess) return gfail(g, TTS_ECUDA, "cudaSetDevice failed");
    for (int k = 0; k < per; ++k) {
      memcpy(&send[2 * k], &scores[size_t(i) * per + k], 4);  // bit pattern of the f32 score
      send[2 * k + 1] = lens[size_t(i) * per + k];
    }
    if (cudaMemcpyAsync(g->d_send[i], send.data(), send.size() * 4, cudaMemcpyHostToDevice, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess)
      return gfail(g, TTS_ECUDA, "upload failed");
  }
  TTS_NCCL_TRY(g, nccl().GroupStart());
  for (int i = 0; i < nl; ++i) {
    tts_ctx *c = g->ctx[i];
    cudaSetDevice(c->cfg.device);
    TTS_NCCL_TRY(g, nccl().AllGather(g->d_send[i], g->d_recv[i], size_t(per) * 2, ncclInt32, g->comm[i], c->stream));
  }
  TTS_NCCL_TRY(g, nccl().GroupEnd());
  for (int i = 0; i < nl; ++i) {
    cudaSetDevice(g->ctx[i]->cfg.device);
    if (cudaStreamSynchronize(g->ctx[i]->stream) != cudaSuccess) return gfail(g, TTS_ECUDA, "all-gather failed");
  }
  std::vector<int32_t> all(size_t(g->world) * per * 2);
  cudaSetDevice(g->ctx[0]->cfg.device);
  if (cudaMemcpy(all.data(), g->d_recv[0], all.size() * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
    return gfail(g, TTS_ECUDA, "download failed");
  int best = 0;
  float best_s = -INFINITY;
  for (int k = 0; k < g->world * per; ++k) {
    float s;
    memcpy(&s, &all[2 * k], 4);
    if (scores_all) scores_all[k] = s;
    if (lens_all) lens_all[k] = all[2 * k + 1];
    if (!isnan(s) && s > best_s) {
      best_s = s;
      best = k;
    }
  }
  if (winner) *winner = best;
  return TTS_OK;
}

void tts_group_free(tts_group *g) {
  if (!g) return;
  for (size_t i = 0; i < g->comm.size(); ++i) {
    cudaSetDevice(g->ctx[i]->cfg.device);
    if (g->d_send[i]) tts::ctx_free(g->ctx[i], g->d_send[i]);
    if (g->d_recv[i]) tts::ctx_free(g->ctx[i], g->d_recv[i]);
    if (nccl().ok && g->comm[i]) nccl().CommDestroy(g->comm[i]);
  }
  delete g;
}

}  // extern "C"
