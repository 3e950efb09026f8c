"""ctypes binding of the host-side C-ABI (include/tortoise_host.h).

`HostLib()` loads libtortoise_host.so (pure C++, works without a GPU); pass
`full=True` to bind libtortoise_b200.so instead, which additionally exports the stage
drivers tts_host_autoregressive / tts_host_diffusion / tts_host_vocoder."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "libtortoise_host.so")
FULL_LIB_PATH = os.path.join(_HERE, "libtortoise_b200.so")


class AROptions(C.Structure):
    _fields_ = [("max_steps", C.c_int32), ("forced_codes", C.c_int32), ("per_candidate_stop", C.c_int32),
                ("full_logits", C.c_int32), ("skip_latents", C.c_int32), ("reserved", C.c_int32 * 3)]


class HostLib:
    def __init__(self, full: bool = False):
        path = FULL_LIB_PATH if full else HOST_LIB_PATH
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not built (run __graft_entry__.build())")
        lib = C.CDLL(path)
        P = C.POINTER
        vp, i32, f32p, i32p = C.c_void_p, C.c_int32, P(C.c_float), P(C.c_int32)
        lib.tts_rng_create.argtypes = [C.c_uint32]
        lib.tts_rng_create.restype = vp
        lib.tts_rng_seed.argtypes = [vp, C.c_uint32]
        lib.tts_rng_free.argtypes = [vp]
        lib.tts_rng_uniform.argtypes = [vp]
        lib.tts_rng_uniform.restype = C.c_float
        lib.tts_rng_normal.argtypes = [vp, f32p, C.c_int64]
        lib.tts_host_tokenize.argtypes = [C.c_char_p, C.c_char_p, i32p, i32]
        lib.tts_host_vocab_size.argtypes = [C.c_char_p]
        lib.tts_host_normalize_text.argtypes = [C.c_char_p, C.c_char_p, i32]
        lib.tts_host_split_text.argtypes = [C.c_char_p, i32, i32p, i32]
        lib.tts_host_sample.argtypes = [vp, f32p, i32p, i32, i32, i32p, f32p]
        lib.tts_host_sample_reference_order.argtypes = [vp, f32p, i32p, i32, i32, i32p]
        lib.tts_host_sample_sparse.argtypes = [vp, f32p, i32p, i32, i32p, i32, i32p, f32p]
        lib.tts_host_apply_padding.argtypes = [i32p, i32, i32p]
        lib.tts_host_trim_count.argtypes = [i32p]
        lib.tts_host_write_wav.argtypes = [C.c_char_p, f32p, C.c_int64, i32]
        lib.tts_host_timestep_map.argtypes = [i32, i32p]
        lib.tts_host_timestep_embedding.argtypes = [i32, f32p]
        lib.tts_host_relative_position_buckets.argtypes = [i32, i32p]
        lib.tts_host_ddpm_schedule.argtypes = [i32, f32p]
        if full:
            lib.tts_host_autoregressive.argtypes = [vp, vp, i32p, i32, f32p, i32, P(AROptions), i32p, f32p, i32p,
                                                    f32p, i32p]
            lib.tts_host_diffusion.argtypes = [vp, vp, f32p, i32, i32, f32p, i32p]
            lib.tts_host_latents.argtypes = [vp, i32p, i32, f32p, i32p, f32p, i32p]
            lib.tts_host_autoregressive_multi.argtypes = [vp, P(vp), i32, P(i32p), i32p, f32p, i32p, i32, i32p, i32p, i32p]
            lib.tts_host_diffusion_batch.argtypes = [vp, P(vp), i32, P(f32p), i32p, i32, P(f32p), i32p]
            lib.tts_host_vocoder.argtypes = [vp, vp, f32p, i32, f32p]
        self.lib = lib
        self.full = full

    # ---- rng
    def rng(self, seed: int):
        return Rng(self, seed)

    # ---- tokenizer
    def tokenize(self, tokenizer_json: str, message: str) -> list[int]:
        out = (C.c_int32 * 2048)()
        n = self.lib.tts_host_tokenize(os.fsencode(tokenizer_json), message.encode(), out, 2048)
        if n < 0:
            raise RuntimeError(f"tokenize failed: {n}")
        return list(out[:n])

    def normalize_text(self, text: str) -> str:
        buf = C.create_string_buffer(16 * len(text.encode()) + 64)
        n = self.lib.tts_host_normalize_text(text.encode(), buf, len(buf))
        if n < 0:
            raise RuntimeError(f"normalize_text failed: {n}")
        return buf.value.decode()

    def split_text(self, text: str, max_chars: int) -> list[str]:
        raw = text.encode()
        spans = (C.c_int32 * (2 * (len(raw) + 1)))()
        n = self.lib.tts_host_split_text(raw, max_chars, spans, len(raw) + 1)
        if n < 0:
            raise RuntimeError(f"split_text failed: {n}")
        return [raw[spans[2 * i]:spans[2 * i + 1]].decode() for i in range(n)]

    def vocab_size(self, tokenizer_json: str) -> int:
        return self.lib.tts_host_vocab_size(os.fsencode(tokenizer_json))

    # ---- sampling
    def sample(self, rng, logits, prev, literal=False, want_logprob=False):
        logits = np.ascontiguousarray(logits, dtype=np.float32)
        B = logits.shape[0]
        prev = np.ascontiguousarray(prev, dtype=np.int32).reshape(B, -1)
        out = np.empty(B, dtype=np.int32)
        lp = np.empty(B, dtype=np.float32)
        f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        if literal:
            rc = self.lib.tts_host_sample_reference_order(rng.h, logits.ctypes.data_as(f32p), prev.ctypes.data_as(i32p),
                                                          prev.shape[1], B, out.ctypes.data_as(i32p))
        else:
            rc = self.lib.tts_host_sample(rng.h, logits.ctypes.data_as(f32p), prev.ctypes.data_as(i32p), prev.shape[1],
                                          B, out.ctypes.data_as(i32p), lp.ctypes.data_as(f32p))
        if rc != 0:
            raise RuntimeError(f"sample failed: {rc}")
        return (out, lp) if want_logprob else out

    def sample_sparse(self, rng, vals, idx, prev):
        """one candidate from (value, index) pairs; returns (sample, logprob) or None when the full row is needed"""
        vals = np.ascontiguousarray(vals, dtype=np.float32)
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        prev = np.ascontiguousarray(prev, dtype=np.int32)
        out, lp = C.c_int32(), C.c_float()
        rc = self.lib.tts_host_sample_sparse(rng.h, vals.ctypes.data_as(C.POINTER(C.c_float)),
                                             idx.ctypes.data_as(C.POINTER(C.c_int32)), len(vals),
                                             prev.ctypes.data_as(C.POINTER(C.c_int32)), len(prev), C.byref(out), C.byref(lp))
        if rc < 0:
            raise ValueError("bad argument")
        return None if rc == 1 else (out.value, lp.value)

    def apply_padding(self, seq):
        seq = np.ascontiguousarray(seq, dtype=np.int32)
        out = np.empty(502, dtype=np.int32)
        rc = self.lib.tts_host_apply_padding(seq.ctypes.data_as(C.POINTER(C.c_int32)), len(seq),
                                             out.ctypes.data_as(C.POINTER(C.c_int32)))
        if rc != 0:
            raise RuntimeError(f"apply_padding failed: {rc}")
        return out

    def trim_count(self, codes500):
        codes500 = np.ascontiguousarray(codes500, dtype=np.int32)
        assert codes500.size == 500
        return self.lib.tts_host_trim_count(codes500.ctypes.data_as(C.POINTER(C.c_int32)))

    def write_wav(self, path, data, rate=24000):
        data = np.ascontiguousarray(data, dtype=np.float32)
        rc = self.lib.tts_host_write_wav(os.fsencode(path), data.ctypes.data_as(C.POINTER(C.c_float)), data.size, rate)
        if rc != 0:
            raise RuntimeError("write_wav failed")

    def timestep_map(self, n):
        out = np.empty(n, dtype=np.int32)
        self.lib.tts_host_timestep_map(n, out.ctypes.data_as(C.POINTER(C.c_int32)))
        return out

    def timestep_embedding(self, t):
        out = np.empty(1024, dtype=np.float32)
        self.lib.tts_host_timestep_embedding(t, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    def relative_position_buckets(self, n):
        out = np.empty((n, n), dtype=np.int32)
        self.lib.tts_host_relative_position_buckets(n, out.ctypes.data_as(C.POINTER(C.c_int32)))
        return out

    def ddpm_schedule(self, n):
        out = np.empty((n, 9), dtype=np.float32)
        self.lib.tts_host_ddpm_schedule(n, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    # ---- stage drivers (full library only)
    def autoregressive(self, engine, rng, tokens, voice, B, max_steps=0, forced_codes=0, per_candidate_stop=False,
                       full_logits=False, skip_latents=False):
        assert self.full
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        voice = np.ascontiguousarray(voice, dtype=np.float32)
        opt = AROptions(max_steps=max_steps, forced_codes=forced_codes,
                        per_candidate_stop=1 if per_candidate_stop else 0, full_logits=1 if full_logits else 0,
                        skip_latents=1 if skip_latents else 0)
        codes = np.empty((B, 500), dtype=np.int32)
        lat = np.empty((1 if skip_latents else B, 500, 1024), dtype=np.float32)
        nlat = np.empty(B, dtype=np.int32)
        score = np.empty(B, dtype=np.float32)
        steps = C.c_int32()
        f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        rc = self.lib.tts_host_autoregressive(engine.h, rng.h, tokens.ctypes.data_as(i32p), len(tokens),
                                              voice.ctypes.data_as(f32p), B, C.byref(opt), codes.ctypes.data_as(i32p),
                                              lat.ctypes.data_as(f32p), nlat.ctypes.data_as(i32p),
                                              score.ctypes.data_as(f32p), C.byref(steps))
        if rc != 0:
            raise RuntimeError(f"tts_host_autoregressive failed ({rc}): {engine.lib.tts_last_error(engine.h).decode()}")
        return codes, lat, nlat, score, steps.value

    def autoregressive_multi(self, engine, rngs, tokens_list, voice, forced_codes=None, max_steps=0):
        """utterance-batched decode loop: U prompts on one batched decode launch per step.
        Returns (codes [U][500], n_latents [U], steps [U]); latents per utterance via latents()."""
        assert self.full
        U = len(tokens_list)
        arrs = [np.ascontiguousarray(t, dtype=np.int32) for t in tokens_list]
        i32p, f32p = C.POINTER(C.c_int32), C.POINTER(C.c_float)
        tok_p = (i32p * U)(*[a.ctypes.data_as(i32p) for a in arrs])
        T = np.array([len(a) for a in arrs], dtype=np.int32)
        rng_p = (C.c_void_p * U)(*[r.h for r in rngs])
        voice = np.ascontiguousarray(voice, dtype=np.float32)
        forced = None if forced_codes is None else np.ascontiguousarray(forced_codes, dtype=np.int32)
        codes = np.empty((U, 500), dtype=np.int32)
        nlat = np.empty(U, dtype=np.int32)
        steps = np.empty(U, dtype=np.int32)
        rc = self.lib.tts_host_autoregressive_multi(engine.h, rng_p, U, tok_p, T.ctypes.data_as(i32p), voice.ctypes.data_as(f32p),
                                                    forced.ctypes.data_as(i32p) if forced is not None else None, max_steps,
                                                    codes.ctypes.data_as(i32p), nlat.ctypes.data_as(i32p), steps.ctypes.data_as(i32p))
        if rc != 0:
            raise RuntimeError(f"tts_host_autoregressive_multi failed ({rc}): {engine.lib.tts_last_error(engine.h).decode()}")
        return codes, nlat, steps

    def latents(self, engine, tokens, voice, codes500):
        """latent pass + trim for one candidate: returns [n][1024]"""
        assert self.full
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        voice = np.ascontiguousarray(voice, dtype=np.float32)
        codes500 = np.ascontiguousarray(codes500, dtype=np.int32)
        lat = np.empty((500, 1024), dtype=np.float32)
        n = C.c_int32()
        f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        rc = self.lib.tts_host_latents(engine.h, tokens.ctypes.data_as(i32p), len(tokens), voice.ctypes.data_as(f32p),
                                       codes500.ctypes.data_as(i32p), lat.ctypes.data_as(f32p), C.byref(n))
        if rc != 0:
            raise RuntimeError(f"tts_host_latents failed ({rc}): {engine.lib.tts_last_error(engine.h).decode()}")
        return lat[:n.value]

    def diffusion(self, engine, rng, latents, n_steps=80):
        assert self.full
        latents = np.ascontiguousarray(latents, dtype=np.float32)
        L = latents.shape[0]
        S = L * 4 * 24000 // 22050
        mel = np.empty((100, S), dtype=np.float32)
        s_out = C.c_int32()
        rc = self.lib.tts_host_diffusion(engine.h, rng.h, latents.ctypes.data_as(C.POINTER(C.c_float)), L, n_steps,
                                         mel.ctypes.data_as(C.POINTER(C.c_float)), C.byref(s_out))
        if rc != 0:
            raise RuntimeError(f"tts_host_diffusion failed ({rc}): {engine.lib.tts_last_error(engine.h).decode()}")
        assert s_out.value == S
        return mel

    def diffusion_batch(self, engine, rngs, latents_list, n_steps=80):
        """U utterances on one launch set (utterance batching); returns a list of mels [100][S_u]"""
        assert self.full
        U = len(latents_list)
        lats = [np.ascontiguousarray(l, dtype=np.float32) for l in latents_list]
        L = np.array([l.shape[0] for l in lats], dtype=np.int32)
        mels = [np.empty((100, int(l) * 4 * 24000 // 22050), dtype=np.float32) for l in L]
        f32p = C.POINTER(C.c_float)
        lat_p = (f32p * U)(*[l.ctypes.data_as(f32p) for l in lats])
        mel_p = (f32p * U)(*[m.ctypes.data_as(f32p) for m in mels])
        rng_p = (C.c_void_p * U)(*[r.h for r in rngs])
        s_out = np.empty(U, dtype=np.int32)
        rc = self.lib.tts_host_diffusion_batch(engine.h, rng_p, U, lat_p, L.ctypes.data_as(C.POINTER(C.c_int32)), n_steps, mel_p,
                                               s_out.ctypes.data_as(C.POINTER(C.c_int32)))
        if rc != 0:
            raise RuntimeError(f"tts_host_diffusion_batch failed ({rc}): {engine.lib.tts_last_error(engine.h).decode()}")
        return mels

    def vocoder(self, engine, rng, mel):
        assert self.full
        mel = np.ascontiguousarray(mel, dtype=np.float32)
        S = mel.shape[1]
        audio = np.empty((S + 10) * 256 - 6, dtype=np.float32)
        rc = self.lib.tts_host_vocoder(engine.h, rng.h, mel.ctypes.data_as(C.POINTER(C.c_float)), S,
                                       audio.ctypes.data_as(C.POINTER(C.c_float)))
        if rc != 0:
            raise RuntimeError(f"tts_host_vocoder failed ({rc}): {engine.lib.tts_last_error(engine.h).decode()}")
        return audio


class Rng:
    def __init__(self, hostlib: HostLib, seed: int):
        self.hl = hostlib
        self.h = C.c_void_p(hostlib.lib.tts_rng_create(seed))

    def seed(self, s):
        self.hl.lib.tts_rng_seed(self.h, s)

    def uniform(self):
        return self.hl.lib.tts_rng_uniform(self.h)

    def normal(self, n):
        out = np.empty(n, dtype=np.float32)
        self.hl.lib.tts_rng_normal(self.h, out.ctypes.data_as(C.POINTER(C.c_float)), n)
        return out

    def __del__(self):
        try:
            self.hl.lib.tts_rng_free(self.h)
        except Exception:
            pass
