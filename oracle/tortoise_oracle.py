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
