// vocoder.cu -- UnivNet generator: loader + forward (reference: vocoder_model_load
// main.cpp:1665-2021, vocoder_graph main.cpp:4068-4483, vocoder() main.cpp:6044-6127).
// Time-major activations; every ggml_conv_1d (F16 x F16 -> F32) runs as an implicit GEMM on
// tensor cores; the location-variable convolution is ONE fused kernel per layer (per-frame
// 32->64 k3 conv + bias + sigmoid*tanh gate + residual) instead of the reference's
// unfold -> mul_mat -> 31 adds -> split -> gate chain (main.cpp:4378-4455).
#include <set>

#include "diff_kernels.cuh"
#include "engine.h"
#include "gemm_launch.cuh"

namespace tts {

struct VocBlock {
  float *convt_w, *convt_b;                       // [32 in][32 out][K] f32 (ConvTranspose1d stays F32, A-5)
  __half *kp_in_w;  float *kp_in_b;               // k5 100(pad 128) -> 64
  __half *kp_r1_w[3], *kp_r3_w[3]; float *kp_r1_b[3], *kp_r3_b[3];
  __half *kp_kernel_w; float *kp_kernel_b;        // k3 64 -> 24576
  __half *kp_bias_w;   float *kp_bias_b;          // k3 64 -> 256
  __half *cb_w[4]; float *cb_b[4];                // k3 dilated 32 -> 32
};
struct VocModel {
  bool loaded = false;
  __half *pre_w, *post_w;
  float *pre_b, *post_b;
  VocBlock blk[3];
  int capN0 = 0;
  float *mel_dev = nullptr, *noise_dev = nullptr, *C = nullptr, *CO = nullptr, *KT = nullptr, *BT = nullptr;
  float *X = nullptr, *X2 = nullptr, *Y = nullptr, *audio = nullptr;
  __half *MEL16 = nullptr, *C16 = nullptr, *Z16 = nullptr, *A16 = nullptr;
  float *h_pin = nullptr;
  size_t h_pin_bytes = 0;
};

constexpr int VOC_HALO = 27;

// normalised mel [100][S] -> denormalised (main.cpp:5575-5584) + 10 pad frames of -11.5129
// (main.cpp:6051-6054) -> f16 [N0 + 4][128] (halo 2 for the k5 conv, channels padded to 128)
__global__ void __launch_bounds__(128) voc_mel_kernel(const float *mel, __half *out, int S) {
  pdl_launch_dependents();
  pdl_wait();
  const int N0 = S + 10, row = blockIdx.x, t = row - 2, c = threadIdx.x;
  float v = 0.f;
  if (t >= 0 && t < N0 && c < 100) {
    if (t < S) {
      const float MX = 2.3143386840820312f, MN = -11.512925148010254f;
      v = __fadd_rn(__fmul_rn(__fdiv_rn(__fadd_rn(mel[size_t(c) * S + t], 1.0f), 2.0f), __fsub_rn(MX, MN)), MN);
    } else {
      v = float(-11.5129);
    }
  }
  out[size_t(row) * 128 + c] = __float2half_rn(v);
}

// noise z [64][N0] (channel-major as drawn) -> reflect-pad 3 (ggml_pad_reflect_1d,
// ggml.c:13993-14028) -> f16 [N0 + 6][64]
__global__ void __launch_bounds__(64) voc_noise_kernel(const float *z, __half *out, int N0) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x, c = threadIdx.x;
  int t = row - 3;
  if (t < 0) t = -t;
  if (t >= N0) t = 2 * (N0 - 1) - t;
  out[size_t(row) * 64 + c] = __float2half_rn(z[size_t(c) * N0 + t]);
}

// f32 [T][C] -> (optional leaky relu 0.2) -> f16 [T + 2H][C], zero halo rows
__global__ void __launch_bounds__(256) voc_act_f16_kernel(const float *X, __half *out, int T, int C, int H,
                                                          int lrelu) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t n = size_t(T + 2 * H) * C;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const int row = int(i / C), c = int(i % C), t = row - H;
    float v = 0.f;
    if (t >= 0 && t < T) {
      v = X[size_t(t) * C + c];
      if (lrelu) v = v > 0.f ? v : 0.2f * v;
    }
    out[i] = __float2half_rn(v);
  }
}

// ConvTranspose1d(32 -> 32, kernel K, stride s) on leaky_relu(x), cropped by p on both sides,
// + bias (main.cpp:4145-4167; ggml_compute_forward_conv_transpose_1d_f32 ggml.c:14955-15052).
// x [L][32] -> out [L*s][32].  One thread per (output frame, out channel).
__global__ void __launch_bounds__(256) voc_convt_kernel(const float *x, const float *w, const float *bias,
                                                        float *out, int L, int K, int s, int p) {
  pdl_launch_dependents();
  pdl_wait();
  const int n = L * s * 32;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
    const int oc = idx & 31, j = idx >> 5;
    const int jp = j + p;
    float acc = 0.f;
    for (int k = jp % s; k < K; k += s) {
      const int i = (jp - k) / s;
      if (i < 0 || i >= L) continue;
      float v = 0.f;
      for (int ic = 0; ic < 32; ++ic) {
        float xv = x[size_t(i) * 32 + ic];
        xv = xv > 0.f ? xv : 0.2f * xv;
        v = fmaf(xv, w[(size_t(ic) * 32 + oc) * K + k], v);
      }
      acc += v;
    }
    out[idx] = acc + bias[oc];
  }
}

// Fused location-variable convolution layer (main.cpp:4378-4455):
//   o[oc, l*hop+s] = Bt[l][layer*64+oc] + sum_ic sum_k ypad[l*hop+s+k-1][ic] * Kt[l][layer*6144+ic*192+oc*3+k]
//   x[l*hop+s][c] += sigmoid(o[c]) * tanh(o[32+c])
// grid (N0, nchunk) x 256 threads = 32 channels x 8 positions; the frame's 6144 kernel
// taps live in shared memory.
__global__ void __launch_bounds__(256) voc_lvc_kernel(const float *Y, const float *KT, const float *BT, float *X,
                                                      int N0, int hop, int layer) {
  __shared__ float Ks[6144];
  __shared__ float ys[10][32];
  pdl_launch_dependents();
  pdl_wait();
  const int l = blockIdx.x, tid = threadIdx.x, c = tid & 31, sl = tid >> 5;
  const int T = N0 * hop;
  const float *kp = KT + size_t(l) * 24576 + layer * 6144;
  for (int i = tid; i < 6144 / 4; i += 256)
    reinterpret_cast<float4 *>(Ks)[i] = reinterpret_cast<const float4 *>(kp)[i];
  const float b1 = BT[size_t(l) * 256 + layer * 64 + c], b2 = BT[size_t(l) * 256 + layer * 64 + 32 + c];
  const int nchunks = hop / 8;
  for (int ch = blockIdx.y; ch < nchunks; ch += gridDim.y) {
    const int s0 = ch * 8;
    __syncthreads();
    for (int i = tid; i < 10 * 32; i += 256) {
      const int r = i >> 5, ic = i & 31;
      const int t = l * hop + s0 + r - 1;
      ys[r][ic] = (t >= 0 && t < T) ? Y[size_t(t) * 32 + ic] : 0.f;
    }
    __syncthreads();
    float a1 = 0.f, a2 = 0.f;
#pragma unroll 4
    for (int ic = 0; ic < 32; ++ic) {
      const float y0 = ys[sl][ic], y1 = ys[sl + 1][ic], y2 = ys[sl + 2][ic];
      const float *k1 = Ks + ic * 192 + c * 3, *k2 = k1 + 96;
      float p1 = y0 * k1[0];
      p1 = fmaf(y1, k1[1], p1);
      p1 = fmaf(y2, k1[2], p1);
      float p2 = y0 * k2[0];
      p2 = fmaf(y1, k2[1], p2);
      p2 = fmaf(y2, k2[2], p2);
      a1 += p1;
      a2 += p2;
    }
    const float o1 = a1 + b1, o2 = a2 + b2;
    const float gate = (1.0f / (1.0f + expf(-o1))) * tanhf(o2);
    const size_t xi = size_t(l * hop + s0 + sl) * 32 + c;
    X[xi] += gate;
  }
}

void voc_load(tts_ctx *c, const char *path) {
  Container ct;
  std::string err;
  if (!ct.open(path, err)) throw ArgError(err, TTS_EIO);
  if (c->voc && c->voc->loaded) throw ArgError("vocoder model already loaded in this context");
  if (!c->voc) c->voc = new VocModel();
  VocModel &m = *c->voc;
  std::set<std::string> known;
  auto f32 = [&](const std::string &n, std::vector<int> ne) { known.insert(n); return upload_f32(c, ct, n, ne); };
  auto convw = [&](const std::string &n, int OC, int IC, int K, int ICpad) {
    known.insert(n);
    auto it = ct.tensors.find(n);
    if (it == ct.tensors.end()) throw ArgError("tensor '" + n + "' missing from " + ct.path, TTS_EIO);
    if (it->second.nelem != size_t(OC) * IC * K || it->second.ne[0] != K ||
        (it->second.ne.size() > 1 ? it->second.ne[1] : 1) != IC)
      throw ArgError("tensor '" + n + "' has wrong shape in model file", TTS_EIO);
    size_t nel = 0;
    read_tensor_to_staging(c, ct, n, &nel);
    __half *d = nullptr;
    TTS_CUDA_TRY(ctx_malloc(c, &d, size_t(K) * OC * ICpad * 2));
    conv_weight_kernel<<<1024, 256, 0, c->stream>>>(c->d_scratch, d, OC, IC, K, ICpad);
    TTS_CUDA_TRY(cudaGetLastError());
    TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return d;
  };
  m.pre_w = convw("conv_pre.weight", 32, 64, 7, 64);
  m.pre_b = f32("conv_pre.bias", {32});
  const int Ks[3] = {16, 16, 8};
  for (int i = 0; i < 3; ++i) {
    const std::string p = "res_stack." + std::to_string(i) + ".";
    VocBlock &b = m.blk[i];
    b.convt_w = f32(p + "convt_pre.1.weight", {Ks[i], 32, 32});
    b.convt_b = f32(p + "convt_pre.1.bias", {32});
    b.kp_in_w = convw(p + "kernel_predictor.input_conv.0.weight", 64, 100, 5, 128);
    b.kp_in_b = f32(p + "kernel_predictor.input_conv.0.bias", {64});
    for (int r = 0; r < 3; ++r) {
      const std::string q = p + "kernel_predictor.residual_convs." + std::to_string(r) + ".";
      b.kp_r1_w[r] = convw(q + "1.weight", 64, 64, 3, 64);
      b.kp_r1_b[r] = f32(q + "1.bias", {64});
      b.kp_r3_w[r] = convw(q + "3.weight", 64, 64, 3, 64);
      b.kp_r3_b[r] = f32(q + "3.bias", {64});
    }
    b.kp_kernel_w = convw(p + "kernel_predictor.kernel_conv.weight", 24576, 64, 3, 64);
    b.kp_kernel_b = f32(p + "kernel_predictor.kernel_conv.bias", {24576});
    b.kp_bias_w = convw(p + "kernel_predictor.bias_conv.weight", 256, 64, 3, 64);
    b.kp_bias_b = f32(p + "kernel_predictor.bias_conv.bias", {256});
    for (int l = 0; l < 4; ++l) {
      const std::string q = p + "conv_blocks." + std::to_string(l) + ".1.";
      b.cb_w[l] = convw(q + "weight", 32, 32, 3, 32);
      b.cb_b[l] = f32(q + "bias", {32});
    }
  }
  m.post_w = convw("conv_post.1.weight", 1, 32, 7, 32);
  m.post_b = f32("conv_post.1.bias", {1});
  for (const auto &n : ct.order)
    if (!known.count(n)) throw ArgError("unknown tensor '" + n + "' in model file", TTS_EIO);
  m.loaded = true;
}

void voc_free(tts_ctx *c) {
  if (c->voc) {
    delete c->voc;  // its buffers belong to the context's allocation registry (tts_free)
    c->voc = nullptr;
  }
}

template <typename T>
static void vgrow(tts_ctx *c, T **p, size_t n) {
  if (*p) ctx_free(c, *p);
  TTS_CUDA_TRY(ctx_malloc(c, p, n * sizeof(T)));
}

static void vtg(tts_ctx *c, const Launcher &L, const __half *X16, const __half *W, const float *bias, float *C, int M,
                int N, int K, int ldc, int epi, int taps, int dil, int pad, int halo, int T) {
  TGemmArgs g{X16, nullptr, W, nullptr, bias, C, nullptr, nullptr, M, N, K, K, ldc, 0, epi, taps, dil, pad, halo, T};
  launch_gemm(L, g);
}

void voc_run(tts_ctx *c, const float *mel, int S, const float *noise, float *audio) {
  if (!c->voc || !c->voc->loaded) throw ArgError("vocoder model not loaded");
  if (S < 1 || S > 4096) throw ArgError("bad mel length", TTS_ELIMIT);
  VocModel &m = *c->voc;
  const int N0 = S + 10;
  const size_t Tmax = size_t(N0) * 256;
  if (N0 > m.capN0) {
    vgrow(c, &m.mel_dev, size_t(100) * S + 16);
    vgrow(c, &m.noise_dev, size_t(N0) * 64);
    vgrow(c, &m.C, size_t(N0) * 64);
    vgrow(c, &m.CO, size_t(N0) * 64);
    vgrow(c, &m.KT, size_t(N0) * 24576);
    vgrow(c, &m.BT, size_t(N0) * 256);
    vgrow(c, &m.X, Tmax * 32);
    vgrow(c, &m.X2, Tmax * 32);
    vgrow(c, &m.Y, Tmax * 32);
    vgrow(c, &m.audio, Tmax);
    vgrow(c, &m.MEL16, size_t(N0 + 4) * 128);
    vgrow(c, &m.C16, size_t(N0 + 2) * 64);
    vgrow(c, &m.Z16, size_t(N0 + 6) * 64);
    vgrow(c, &m.A16, (Tmax + 2 * VOC_HALO) * 32);
    m.capN0 = N0;
  }
  Launcher L{c->stream, c->use_pdl, &c->launches};
  TTS_CUDA_TRY(cudaEventRecord(c->ev0, c->stream));
  TTS_CUDA_TRY(cudaMemcpyAsync(m.mel_dev, mel, size_t(100) * S * 4, cudaMemcpyHostToDevice, c->stream));
  TTS_CUDA_TRY(cudaMemcpyAsync(m.noise_dev, noise, size_t(N0) * 64 * 4, cudaMemcpyHostToDevice, c->stream));
  L(voc_mel_kernel, dim3(N0 + 4), dim3(128), 0, (const float *)m.mel_dev, m.MEL16, S);
  L(voc_noise_kernel, dim3(N0 + 6), dim3(64), 0, (const float *)m.noise_dev, m.Z16, N0);
  // conv_pre: k7 over the reflect-padded noise, 64 -> 32 (main.cpp:4114-4130)
  vtg(c, L, m.Z16, m.pre_w, m.pre_b, m.X, N0, 32, 64, 32, E_BIAS, 7, 1, 3, 3, N0);
  const int strides[3] = {8, 8, 4}, pads[3] = {4, 4, 2}, Ks[3] = {16, 16, 8}, hops[3] = {8, 64, 256};
  const int dils[4] = {1, 3, 9, 27};
  int Tlen = N0;
  float *x = m.X, *x2 = m.X2;
  auto blocks = [](size_t n) { return dim3(unsigned(std::min<size_t>((n + 255) / 256, 148 * 8))); };
  for (int i = 0; i < 3; ++i) {
    VocBlock &b = m.blk[i];
    // upsample: leaky relu -> ConvTranspose1d -> crop -> + bias
    L(voc_convt_kernel, blocks(size_t(Tlen) * strides[i] * 32), dim3(256), 0, (const float *)x,
      (const float *)b.convt_w, (const float *)b.convt_b, x2, Tlen, Ks[i], strides[i], pads[i]);
    std::swap(x, x2);
    Tlen *= strides[i];
    // kernel predictor on the padded mel (main.cpp:4169-4324)
    vtg(c, L, m.MEL16, b.kp_in_w, b.kp_in_b, m.C, N0, 64, 128, 64, E_BIAS_LRELU, 5, 1, 2, 2, N0);
    for (int r = 0; r < 3; ++r) {
      L(voc_act_f16_kernel, blocks(size_t(N0 + 2) * 64), dim3(256), 0, (const float *)m.C, m.C16, N0, 64, 1, 0);
      vtg(c, L, m.C16, b.kp_r1_w[r], b.kp_r1_b[r], m.CO, N0, 64, 64, 64, E_BIAS_LRELU, 3, 1, 1, 1, N0);
      L(voc_act_f16_kernel, blocks(size_t(N0 + 2) * 64), dim3(256), 0, (const float *)m.CO, m.C16, N0, 64, 1, 0);
      vtg(c, L, m.C16, b.kp_r3_w[r], b.kp_r3_b[r], m.C, N0, 64, 64, 64, E_BIAS_LRELU_RESID, 3, 1, 1, 1, N0);
    }
    L(voc_act_f16_kernel, blocks(size_t(N0 + 2) * 64), dim3(256), 0, (const float *)m.C, m.C16, N0, 64, 1, 0);
    vtg(c, L, m.C16, b.kp_kernel_w, b.kp_kernel_b, m.KT, N0, 24576, 64, 24576, E_BIAS, 3, 1, 1, 1, N0);
    vtg(c, L, m.C16, b.kp_bias_w, b.kp_bias_b, m.BT, N0, 256, 64, 256, E_BIAS, 3, 1, 1, 1, N0);
    for (int l = 0; l < 4; ++l) {
      L(voc_act_f16_kernel, blocks(size_t(Tlen + 2 * VOC_HALO) * 32), dim3(256), 0, (const float *)x, m.A16, Tlen,
        32, VOC_HALO, 1);
      vtg(c, L, m.A16, b.cb_w[l], b.cb_b[l], m.Y, Tlen, 32, 32, 32, E_BIAS_LRELU, 3, dils[l], dils[l], VOC_HALO,
          Tlen);
      const int nchunk = std::min(hops[i] / 8, 4);
      L(voc_lvc_kernel, dim3(N0, nchunk), dim3(256), 0, (const float *)m.Y, (const float *)m.KT, (const float *)m.BT,
        x, N0, hops[i], l);
    }
  }
  // conv_post: leaky relu -> k7 32 -> 1, no padding, no tanh (main.cpp:4459-4475)
  L(voc_act_f16_kernel, blocks(size_t(Tlen) * 32), dim3(256), 0, (const float *)x, m.A16, Tlen, 32, 0, 1);
  const int n_out = Tlen - 6;
  vtg(c, L, m.A16, m.post_w, m.post_b, m.audio, n_out, 1, 32, 1, E_BIAS, 7, 1, 0, 0, n_out);
  if (m.h_pin_bytes < size_t(n_out) * 4) {
    if (m.h_pin) ctx_free_host(c, m.h_pin);
    TTS_CUDA_TRY(ctx_malloc_host(c, &m.h_pin, size_t(n_out) * 4));
    m.h_pin_bytes = size_t(n_out) * 4;
  }
  TTS_CUDA_TRY(cudaMemcpyAsync(m.h_pin, m.audio, size_t(n_out) * 4, cudaMemcpyDeviceToHost, c->stream));
  TTS_CUDA_TRY(cudaEventRecord(c->ev1, c->stream));
  TTS_CUDA_TRY(cudaStreamSynchronize(c->stream));
  TTS_CUDA_TRY(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  c->total_ms += c->last_ms;
  memcpy(audio, m.h_pin, size_t(n_out) * 4);
}

}  // namespace tts
