"""oracle/tortoise_oracle.py -- CPU restatement (numpy) of the reference's hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under tortoise.cpp_b200/ may import this module; it is
used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the CHECKER.

Every function cites the reference lines (balisujohn/tortoise.cpp @ 0eb5a5d) it restates.
The restatement is pinned against the real reference (oracle/_ref/ref_harness, built from
/root/reference by oracle/Makefile) through the fixtures in tests/golden/ -- see
tests/golden/make_golden.py and tests/test_oracle_vs_golden.py.

Numerics follow the reference's ggml CPU kernels: f32 storage, LayerNorm / GroupNorm with
double accumulation (ggml.c:11905-11958, 12229-12304), the F16 round trip on AR q/k/v
(main.cpp:2789-2790), the fp16-table GELU (ggml.c:2193-2218), F16 x F16 -> F32 convolutions
(ggml.c:6493-6508) -- dot-product summation ORDER differs (numpy/BLAS), which is why float
results are compared with tolerances (logits 1e-3, latents/mel/audio 1e-2, the reference's
own bar at main.cpp:6183-6231) while integer results (tokens, codes, buckets) are exact.
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------- helpers
def h16(x: np.ndarray) -> np.ndarray:
    """f32 -> f16 -> f32 round trip (ggml_cpy F32->F16->F32, main.cpp:2789-2790)."""
    return np.asarray(x, dtype=F32).astype(np.float16).astype(F32)


def layer_norm(x: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """ggml_norm over the last axis: double sums, float result (ggml.c:11935-11955)."""
    x = np.asarray(x, dtype=F32)
    mean = (x.astype(np.float64).sum(-1, keepdims=True) / x.shape[-1]).astype(F32)
    d = x - mean
    var = ((d * d).astype(np.float64).sum(-1, keepdims=True) / x.shape[-1]).astype(F32)
    scale = (F32(1.0) / np.sqrt(var + F32(eps))).astype(F32)
    return (d * scale).astype(F32)


def gelu16(x: np.ndarray) -> np.ndarray:
    """ggml CPU GELU = fp16 lookup table (ggml.c:2193-2218, table at ggml.c:3333)."""
    x = np.asarray(x, dtype=F32)
    xh = h16(x)
    g = F32(0.5) * xh * (F32(1.0) + np.tanh(F32(0.79788456080286535587989211986876) * xh *
                                            (F32(1.0) + F32(0.044715) * xh * xh)).astype(F32))
    out = h16(g.astype(F32))
    out = np.where(x <= F32(-10.0), F32(0.0), out)
    out = np.where(x >= F32(10.0), x, out)
    return out.astype(F32)


def softmax_rows(s: np.ndarray) -> np.ndarray:
    """ggml_soft_max (ggml.c:14069-14169): max-subtract, expf, double sum, scale."""
    s = np.asarray(s, dtype=F32)
    m = s.max(-1, keepdims=True)
    e = np.exp((s - m).astype(F32)).astype(F32)
    tot = e.astype(np.float64).sum(-1, keepdims=True)
    return (e * (1.0 / tot).astype(F32)).astype(F32)


# ----------------------------------------------------------------------------- AR stage
class AROracle:
    """GPT-2 mel-token decoder.  weights: dict from
    synth_weights.read_container(ggml-model.bin); matrices keep the FILE orientation."""

    def __init__(self, weights: dict, weight_dtype: str = "f32"):
        self.w = weights
        self.round_w = (lambda a: a) if weight_dtype == "f32" else (lambda a: h16(a))
        self.kcache = None  # list per layer of [B][n][1024]
        self.vcache = None

    def _lw(self, i: int, name: str) -> np.ndarray:
        return self.w[f"inference_model.transformer.h.{i}.{name}"]

    def _mat(self, i: int, name: str) -> np.ndarray:
        return self.round_w(self._lw(i, name))

    def _layer(self, i: int, h: np.ndarray, n_past: int) -> np.ndarray:
        """One transformer layer on h [B][R][1024] whose first row sits at position n_past
        (main.cpp:2718-2983 / 2200-2470)."""
        B, R, _ = h.shape
        a = layer_norm(h) * self._lw(i, "ln_1.weight") + self._lw(i, "ln_1.bias")
        # c_attn stored [in][out] (GPT-2 Conv1D): qkv = a @ W + b, then the F16 round trip
        qkv = h16((a.reshape(B * R, 1024) @ self._mat(i, "attn.c_attn.weight")).reshape(B, R, 3072)
                  + self._lw(i, "attn.c_attn.bias"))
        q, k, v = qkv[..., :1024], qkv[..., 1024:2048], qkv[..., 2048:]
        if n_past == 0:
            self.kcache[i], self.vcache[i] = k.copy(), v.copy()
        else:
            self.kcache[i] = np.concatenate([self.kcache[i], k], axis=1)
            self.vcache[i] = np.concatenate([self.vcache[i], v], axis=1)
        K, V = self.kcache[i], self.vcache[i]
        n = K.shape[1]
        qh = q.reshape(B, R, 16, 64).transpose(0, 2, 1, 3)
        kh = K.reshape(B, n, 16, 64).transpose(0, 2, 1, 3)
        vh = V.reshape(B, n, 16, 64).transpose(0, 2, 1, 3)
        s = (qh @ kh.transpose(0, 1, 3, 2)).astype(F32) * F32(0.125)  # 1/sqrt(64), main.cpp:2868
        qpos = n_past + np.arange(R)[:, None]
        kpos = np.arange(n)[None, :]
        s = np.where(kpos > qpos, -np.inf, s).astype(F32)  # ggml_diag_mask_inf(n_past)
        p = softmax_rows(s)
        o = (p @ vh).astype(F32).transpose(0, 2, 1, 3).reshape(B, R, 1024)
        h = h + ((o.reshape(B * R, 1024) @ self._mat(i, "attn.c_proj.weight")).reshape(B, R, 1024)
                 + self._lw(i, "attn.c_proj.bias")).astype(F32)
        m = layer_norm(h) * self._lw(i, "ln_2.weight") + self._lw(i, "ln_2.bias")
        f = gelu16((m.reshape(B * R, 1024) @ self._mat(i, "mlp.c_fc.weight")).reshape(B, R, 4096)
                   + self._lw(i, "mlp.c_fc.bias"))
        h = h + ((f.reshape(B * R, 4096) @ self._mat(i, "mlp.c_proj.weight")).reshape(B, R, 1024)
                 + self._lw(i, "mlp.c_proj.bias")).astype(F32)
        return h.astype(F32)

    def _final_z(self, h: np.ndarray) -> np.ndarray:
        """Double final norm (SURVEY A-1; main.cpp:2985-3003 / 2475-2499)."""
        w = self.w
        y = layer_norm(h) * w["inference_model.transformer.ln_f.weight"] + w["inference_model.transformer.ln_f.bias"]
        return (layer_norm(y) * w["inference_model.lm_head.0.weight"] + w["inference_model.lm_head.0.bias"]).astype(F32)

    def _logits(self, h_last: np.ndarray) -> np.ndarray:
        z = self._final_z(h_last)
        W = self.round_w(self.w["inference_model.lm_head.1.weight"])  # [8194][1024]
        return (z @ W.T + self.w["inference_model.lm_head.1.bias"]).astype(F32)

    def _stack(self, h: np.ndarray, n_past: int) -> np.ndarray:
        for i in range(30):
            h = self._layer(i, h, n_past)
        return h

    def prefill(self, text: np.ndarray, voice: np.ndarray, B: int) -> np.ndarray:
        """autoregressive_graph(fake_inputs=true) (main.cpp:2586-2666): rows
        [voice | text_emb+text_pos | mel_emb[8192]+mel_pos[0]], tiled over B. -> logits [B][8194]"""
        w = self.w
        T = len(text)
        rows = [np.asarray(voice, dtype=F32)[None, :],
                w["text_embedding.weight"][text] + w["text_pos_embedding.emb.weight"][:T],
                (w["mel_embedding.weight"][8192] + w["mel_pos_embedding.emb.weight"][0])[None, :]]
        h = np.concatenate(rows, 0).astype(F32)[None].repeat(B, 0)
        self.kcache, self.vcache = [None] * 30, [None] * 30
        self.n_past = T + 2
        h = self._stack(h, 0)
        return self._logits(h[:, -1, :])

    def step(self, tokens: np.ndarray, pos_id: int) -> np.ndarray:
        """autoregressive_graph(fake_inputs=false) (main.cpp:2668-2692, 2718-3029)."""
        w = self.w
        h = (w["mel_embedding.weight"][np.asarray(tokens)] + w["mel_pos_embedding.emb.weight"][pos_id]).astype(F32)
        h = self._stack(h[:, None, :], self.n_past)
        self.n_past += 1
        return self._logits(h[:, 0, :])

    def latents(self, text: np.ndarray, voice: np.ndarray, codes: np.ndarray, n_keep: int = 500,
                parity_quirks: bool = True) -> np.ndarray:
        """autoregressive_latent_graph (main.cpp:2053-2519) on codes [B][502]; returns
        [B][n_keep][1024].  Mel position table quirk A-4 (main.cpp:5326-5333)."""
        w = self.w
        codes = np.asarray(codes)
        B, T = codes.shape[0], len(text)
        pos = np.zeros(B * 502, dtype=np.int64)
        if parity_quirks:
            per = 502 * B // 4
            for i in range(B):
                for c in range(per):
                    if i * per + c < pos.size:
                        pos[i * per + c] = c
        else:
            pos[:] = np.tile(np.arange(502), B)
        pos = pos.reshape(B, 502)
        text_rows = w["text_embedding.weight"][text] + w["text_pos_embedding.emb.weight"][:T]
        hs = []
        for b in range(B):
            mel_rows = w["mel_embedding.weight"][codes[b, :n_keep]] + w["mel_pos_embedding.emb.weight"][pos[b, :n_keep]]
            hs.append(np.concatenate([np.asarray(voice, dtype=F32)[None, :], text_rows, mel_rows], 0))
        h = np.stack(hs).astype(F32)
        self.kcache, self.vcache = [None] * 30, [None] * 30
        h = self._stack(h, 0)
        return self._final_z(h[:, 1 + T:, :])


# ----------------------------------------------------------------------------- host-side integer logic
def apply_padding(seq: list[int]) -> list[int]:
    """main.cpp:4510-4532 (including the 8139 typo: nothing is ever stripped in practice)."""
    v = list(seq)
    while v and v[-1] == 8139:
        v.pop()
    assert len(v) <= 500
    v += [83] * (500 - len(v))
    v[-3:] = [45, 45, 248]
    return [8192] + v + [8193]


def trim_count(codes500: list[int]) -> int:
    """Number of latent frames trim_latents keeps (main.cpp:4894-4911)."""
    calm = 0
    for c, code in enumerate(codes500):
        calm = calm + 1 if code == 83 else 0
        if calm > 8:
            return c
    return 500


def relative_position_buckets(n: int) -> np.ndarray:
    """get_relative_position_buckets (main.cpp:4722-4749): [i (query)][c (key)] int32."""
    i = np.arange(n)[:, None]
    c = np.arange(n)[None, :]
    rp = np.abs(c - i)
    base = np.where(i < c, 16, 0)
    with np.errstate(divide="ignore"):
        large = 8 + (np.log(rp.astype(F32) / F32(8)) / math.log(64.0 / 8.0) * (16.0 - 8.0)).astype(np.int64)
    large = np.minimum(large, 15)
    return (base + np.where(rp < 8, rp, large)).astype(np.int32)


def timestep_embedding(t: int, dim: int = 1024, max_period: int = 10000) -> np.ndarray:
    """generate_timestep_embedding (main.cpp:5496-5521): [cos | sin], freq computed in
    double then narrowed to float, arg = float(t) * freq in float."""
    half = dim // 2
    i = np.arange(half, dtype=F32)
    freq = np.exp(-math.log(max_period) * i.astype(np.float64) / half).astype(F32)
    arg = (F32(t) * freq).astype(F32)
    a64 = arg.astype(np.float64)  # the reference's cos/sin are the double versions, narrowed afterwards
    return np.concatenate([np.cos(a64), np.sin(a64)]).astype(F32)


# ----------------------------------------------------------------------------- shared conv helpers
def conv16(x: np.ndarray, w: np.ndarray, b: np.ndarray | None, pad: int = 0, dil: int = 1) -> np.ndarray:
    """ggml_conv_1d (ggml.c:6493-6508): im2col rounds the input to F16, the kernel was cast to
    F16 (main.cpp:3163-3166), products accumulate in F32.  x [C_in][T] channel-major (the
    reference's layout), w [C_out][C_in][K] (file order), stride 1.  Returns [C_out][T_out]."""
    x = h16(x)
    w = h16(w)
    cout, cin, k = w.shape
    T = x.shape[1]
    xp = np.zeros((cin, T + 2 * pad), dtype=F32)
    xp[:, pad:pad + T] = x
    tout = T + 2 * pad - dil * (k - 1)
    out = np.zeros((cout, tout), dtype=F32)
    for j in range(k):
        out += w[:, :, j] @ xp[:, j * dil:j * dil + tout]
    if b is not None:
        out = out + np.asarray(b, dtype=F32)[:, None]
    return out.astype(F32)


def group_norm32(x: np.ndarray, eps: float = 1e-6) -> np.ndarray:
    """ggml_group_norm, 32 groups over channels x time (ggml.c:12229-12304), x [C][T]."""
    C, T = x.shape
    g = x.reshape(32, (C // 32) * T).astype(F32)
    mean = (g.astype(np.float64).sum(1, keepdims=True) / g.shape[1]).astype(F32)
    d = g - mean
    var = ((d * d).astype(np.float64).sum(1, keepdims=True) / g.shape[1]).astype(F32)
    return (d * (F32(1.0) / np.sqrt(var + F32(eps)))).astype(F32).reshape(C, T)


def silu(x):
    x = np.asarray(x, dtype=F32)
    return (x / (F32(1.0) + np.exp(-x).astype(F32))).astype(F32)


def lrelu(x):
    x = np.asarray(x, dtype=F32)
    return np.where(x > 0, x, F32(0.2) * x).astype(F32)


# ----------------------------------------------------------------------------- diffusion stage
class DiffusionOracle:
    """diffusion_graph (main.cpp:3066-4044) + the DDPM host math of diffusion()
    (main.cpp:5370-5612, 5641-5716, 5970-6030).  Activations [1024][T] like the reference."""

    def __init__(self, weights: dict):
        self.w = weights

    def _attn(self, p: str, x: np.ndarray) -> np.ndarray:
        w = self.w
        C, T = x.shape
        a = group_norm32(x) * w[p + "norm.weight"][:, None] + w[p + "norm.bias"][:, None]
        qkv = conv16(a, w[p + "qkv.weight"][:, :, None], w[p + "qkv.bias"])  # [3072][T]
        qkv = qkv.reshape(16, 192, T)
        q, k, v = qkv[:, :64], qkv[:, 64:128], qkv[:, 128:]
        s = np.einsum("hdi,hdj->hij", q, k).astype(F32) * F32(0.125)
        buckets = relative_position_buckets(T)  # [i query][j key]
        bias = w[p + "relative_pos_embeddings.relative_attention_bias.weight"]  # [32][16]
        s = s + F32(8.0) * bias[buckets].transpose(2, 0, 1)
        pr = softmax_rows(s)
        o = np.einsum("hij,hdj->hdi", pr, v).astype(F32).reshape(1024, T)
        return (x + (w[p + "proj_out.weight"] @ o + w[p + "proj_out.bias"][:, None])).astype(F32)

    def _res(self, p: str, x: np.ndarray, temb: np.ndarray) -> np.ndarray:
        w = self.w
        h = silu(group_norm32(x) * w[p + "in_layers.0.weight"][:, None] + w[p + "in_layers.0.bias"][:, None])
        h = conv16(h, w[p + "in_layers.2.weight"][:, :, None], w[p + "in_layers.2.bias"])
        e = (w[p + "emb_layers.1.weight"] @ silu(temb) + w[p + "emb_layers.1.bias"]).astype(F32)
        scale, shift = e[:1024], e[1024:]
        h = group_norm32(h) * w[p + "out_layers.0.weight"][:, None] + w[p + "out_layers.0.bias"][:, None]
        h = silu(h * (scale + F32(1.0))[:, None] + shift[:, None])
        h = conv16(h, w[p + "out_layers.3.weight"], w[p + "out_layers.3.bias"], pad=1)
        return (x + h).astype(F32)

    def code_embedding(self, latents: np.ndarray, S: int, conditioning_free: bool) -> np.ndarray:
        w = self.w
        if conditioning_free:
            return np.repeat(w["unconditioned_embedding"][:, None], S, axis=1).astype(F32)
        L = latents.shape[0]
        c = conv16(latents.T, w["latent_conditioner.0.weight"], w["latent_conditioner.0.bias"], pad=1)
        for i in range(1, 5):
            c = self._attn(f"latent_conditioner.{i}.", c)
        c = group_norm32(c) * w["code_norm.weight"][:, None] + w["code_norm.bias"][:, None]
        cl = w["diffusion_conditioning_latent"].reshape(-1)
        c = c * (cl[:1024] + F32(1.0))[:, None] + cl[1024:][:, None]
        sf = F32(S) / F32(L)  # ggml_upscale_ext: i00 = i0 / sf0 (float), truncated (ggml.c:15527-15568)
        idx = (np.arange(S, dtype=F32) / sf).astype(np.int64)
        return c[:, idx].astype(F32)

    def eps(self, latents: np.ndarray, x: np.ndarray, timestep: int, conditioning_free: bool) -> np.ndarray:
        """one denoiser evaluation: x [100][S] -> [200][S]"""
        w = self.w
        S = x.shape[1]
        te = timestep_embedding(timestep)
        temb = (w["time_embed.0.weight"] @ te + w["time_embed.0.bias"]).astype(F32)
        temb = (w["time_embed.2.weight"] @ silu(temb) + w["time_embed.2.bias"]).astype(F32)
        c = self.code_embedding(latents, S, conditioning_free)
        for i in range(3):
            c = self._res(f"conditioning_timestep_integrator.{i}.resblk.", c, temb)
            c = self._attn(f"conditioning_timestep_integrator.{i}.attn.", c)
        h = conv16(x, w["inp_block.weight"], w["inp_block.bias"], pad=1)
        h = conv16(np.concatenate([h, c], 0), w["integrating_conv.weight"][:, :, None], w["integrating_conv.bias"])
        for i in range(10):
            h = self._res(f"layers.{i}.resblk.", h, temb)
            h = self._attn(f"layers.{i}.attn.", h)
        for i in range(10, 13):
            h = self._res(f"layers.{i}.", h, temb)
        h = silu(group_norm32(h) * w["out.0.weight"][:, None] + w["out.0.bias"][:, None])
        return conv16(h, w["out.2.weight"], w["out.2.bias"], pad=1)


def ddpm_schedule(n_steps: int = 80):
    """Respaced linear-beta schedule of diffusion() (main.cpp:5390-5400, 5650-5716), in
    sampling order (index d = diffusion_index); float-narrowed like main.cpp:5988-6013."""
    n0 = 4000
    scale = 1000.0 / n0
    bs, be = scale * 0.0001, scale * 0.02
    # `i * (float)(end - start) / (n - 1)` is evaluated in FLOAT (int * float, float / int) and only then
    # added to the double start value (main.cpp:5396-5398)
    betas = np.array([bs + float(F32(F32(i) * F32(be - bs)) / F32(n0 - 1)) for i in range(n0)], dtype=np.float64)
    acp = np.cumprod(1.0 - betas)
    frac = (n0 - 1) / (n_steps - 1)
    tmap = [int(round(i * frac)) if n_steps != 80 else None for i in range(n_steps)]
    cur = 0.0
    tmap = []
    for _ in range(n_steps):
        tmap.append(int(math.floor(cur + 0.5)))
        cur += frac
    last = F32(1.0)
    nb = []
    for i in tmap:
        nb.append(1 - (acp[i] / float(last)))
        last = F32(acp[i])
    nb = np.array(nb)
    acp = np.cumprod(1.0 - nb)
    prev = np.concatenate([[1.0], acp[:-1]])
    post_var = nb * (1.0 - prev) / (1.0 - acp)
    post_logvar = np.log(np.concatenate([[post_var[1]], post_var[1:]]))
    c1 = nb * np.sqrt(prev) / (1.0 - acp)
    c2 = (1.0 - prev) * np.sqrt(1.0 - nb) / (1.0 - acp)
    out = []
    for d in range(n_steps):
        idx = n_steps - 1 - d
        out.append(dict(cfk=F32(2.0) * (F32(1) - F32(idx) / F32(n_steps)), sqrt_recip=F32(np.sqrt(1.0 / acp[idx])),
                        sqrt_recipm1=F32(np.sqrt(1.0 / acp[idx] - 1)), coef1=F32(c1[idx]), coef2=F32(c2[idx]),
                        min_log=F32(post_logvar[idx]), max_log=F32(np.log(nb[idx])), last=idx == 0,
                        timestep=tmap[idx]))
    return out


def ddpm_step(x, out_c, out_u, noise, k) -> np.ndarray:
    """main.cpp:5970-6030 on [100][S] arrays (out_* are [200][S])."""
    eps_c, var_raw, eps_u = out_c[:100], out_c[100:], out_u[:100]
    frac = (var_raw + F32(1)) / F32(2)
    logvar = (frac * k["min_log"] + (F32(1) - frac) * k["max_log"]).astype(F32)  # argument swap, SURVEY A-9
    eps = ((F32(1) + k["cfk"]) * eps_c - k["cfk"] * eps_u).astype(F32)
    x0 = np.clip((k["sqrt_recip"] * x - k["sqrt_recipm1"] * eps).astype(F32), -1.0, 1.0).astype(F32)
    mean = (k["coef1"] * x0 + k["coef2"] * x).astype(F32)
    if k["last"]:
        return mean
    return (mean.astype(np.float64) + np.exp(0.5 * logvar.astype(np.float64)) * noise.astype(np.float64)).astype(F32)


# ----------------------------------------------------------------------------- vocoder stage
class VocoderOracle:
    """vocoder_graph (main.cpp:4068-4483) + vocoder() (main.cpp:6044-6127)."""

    def __init__(self, weights: dict):
        self.w = weights

    def run(self, mel_norm: np.ndarray, noise_flat: np.ndarray, return_intermediates: bool = False):
        w = self.w
        S = mel_norm.shape[1]
        N0 = S + 10
        MX, MN = F32(2.3143386840820312), F32(-11.512925148010254)
        mel = (((mel_norm.astype(F32) + F32(1)) / F32(2)) * (MX - MN) + MN).astype(F32)  # main.cpp:5575-5584
        c0 = np.concatenate([mel, np.full((100, 10), F32(-11.5129), dtype=F32)], 1)  # main.cpp:6051-6054
        z = noise_flat.reshape(64, N0).astype(F32)
        zp = np.concatenate([z[:, 3:0:-1], z, z[:, -2:-5:-1]], 1)  # reflect pad 3 (ggml.c:13993-14028)
        x = conv16(zp, w["conv_pre.weight"], w["conv_pre.bias"])
        inter = {"pre": x.copy()}
        strides, pads, hops = (8, 8, 4), (4, 4, 2), (8, 64, 256)
        for i in range(3):
            p = f"res_stack.{i}."
            s, cp, hop = strides[i], pads[i], hops[i]
            wt = w[p + "convt_pre.1.weight"]  # [in][out][K], F32 (ggml.c:14955-15052)
            K = wt.shape[2]
            xl = lrelu(x)
            L = xl.shape[1]
            full = np.zeros((32, (L - 1) * s + K), dtype=F32)
            for k in range(K):
                full[:, k:k + (L - 1) * s + 1:s] += (wt[:, :, k].T @ xl).astype(F32)
            x = (full[:, cp:full.shape[1] - cp] + w[p + "convt_pre.1.bias"][:, None]).astype(F32)
            inter[f"convt{i}"] = x.copy()
            kp = p + "kernel_predictor."
            c = lrelu(conv16(c0, w[kp + "input_conv.0.weight"], w[kp + "input_conv.0.bias"], pad=2))
            for r in range(3):
                o = lrelu(conv16(c, w[kp + f"residual_convs.{r}.1.weight"], w[kp + f"residual_convs.{r}.1.bias"], pad=1))
                o = lrelu(conv16(o, w[kp + f"residual_convs.{r}.3.weight"], w[kp + f"residual_convs.{r}.3.bias"], pad=1))
                c = (c + o).astype(F32)
            Kt = conv16(c, w[kp + "kernel_conv.weight"], w[kp + "kernel_conv.bias"], pad=1)  # [24576][N0]
            Bt = conv16(c, w[kp + "bias_conv.weight"], w[kp + "bias_conv.bias"], pad=1)      # [256][N0]
            inter[f"kt{i}"] = Kt
            for l in range(4):
                d = (1, 3, 9, 27)[l]
                y = lrelu(conv16(lrelu(x), w[p + f"conv_blocks.{l}.1.weight"], w[p + f"conv_blocks.{l}.1.bias"],
                                 pad=d, dil=d))
                T = y.shape[1]
                yp = np.zeros((32, T + 2), dtype=F32)
                yp[:, 1:T + 1] = y
                kl = Kt[l * 6144:(l + 1) * 6144].reshape(32, 64, 3, N0)  # [ic][oc][k][frame]
                win = np.stack([yp[:, k:k + T] for k in range(3)], 0).reshape(3, 32, N0, hop)  # [k][ic][frame][s]
                o = np.einsum("kilf,iokl->olf", win, kl).astype(F32)  # sum over taps and in-channels
                o = o + Bt[l * 64:(l + 1) * 64][:, :, None]
                o = o.reshape(64, T)
                gate = (F32(1) / (F32(1) + np.exp(-o[:32]).astype(F32))) * np.tanh(o[32:]).astype(F32)
                x = (x + gate).astype(F32)
                inter[f"x{i}_{l}"] = x.copy()
        audio = conv16(lrelu(x), w["conv_post.1.weight"].reshape(1, 32, 7), w["conv_post.1.bias"])
        audio = audio.reshape(-1).astype(F32)
        return (audio, inter) if return_intermediates else audio
