"""tortoise.cpp_b200 -- thin ctypes binding over libtortoise_b200.so (the C-ABI declared in
include/tortoise_b200.h and include/tortoise_host.h).

This Python layer is test/bench plumbing only: the product is the shared library plus the
`tortoise` executable (host C++ above the C-ABI).  There is NO fallback: if the library is
missing or no sm_100 GPU is present, construction raises.

The directory name contains a dot, so import it through `import_pkg()` in /_pkg.py
(importlib by path) rather than a plain `import`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtortoise_b200.so")
INCLUDE_DIR = os.path.join(os.path.dirname(_HERE), "include")

MEL_VOCAB = 8194
MEL_START = 8192
MEL_STOP = 8193
DTYPE_F32 = 0
DTYPE_F16 = 1


class TTSConfig(C.Structure):
    _fields_ = [
        ("device", C.c_int32),
        ("dtype", C.c_int32),
        ("max_batch", C.c_int32),
        ("max_positions", C.c_int32),
        ("parity_quirks", C.c_int32),
        ("reserved", C.c_int32 * 3),
    ]


class TTSError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libtortoise_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load_library() -> C.CDLL:
    """dlopen the in-tree library; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU / PyTorch fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    P = C.POINTER
    i32, f32p, i32p = C.c_int32, P(C.c_float), P(C.c_int32)
    vp = C.c_void_p
    lib.tts_init.argtypes = [P(TTSConfig), P(vp)]
    lib.tts_free.argtypes = [vp]
    lib.tts_free.restype = None
    lib.tts_last_error.argtypes = [vp]
    lib.tts_last_error.restype = C.c_char_p
    lib.tts_version.restype = C.c_int
    for n in ("tts_load_ar", "tts_load_diffusion", "tts_load_vocoder"):
        getattr(lib, n).argtypes = [vp, C.c_char_p]
    lib.tts_ar_prefill.argtypes = [vp, i32p, i32, f32p, i32, f32p]
    lib.tts_ar_step.argtypes = [vp, i32p, i32, f32p]
    lib.tts_ar_step_dev.argtypes = [vp, i32p, i32, P(vp)]
    lib.tts_ar_step_topk.argtypes = [vp, i32p, i32, f32p, i32p, i32p]
    lib.tts_ar_logits.argtypes = [vp, f32p]
    lib.tts_ar_latents.argtypes = [vp, i32p, i32, f32p, i32p, i32, i32, f32p]
    lib.tts_diffusion_eps.argtypes = [vp, f32p, i32, f32p, i32, i32, i32, f32p]
    lib.tts_diffusion_sample.argtypes = [vp, f32p, i32, i32, i32, f32p, f32p]
    lib.tts_vocoder.argtypes = [vp, f32p, i32, f32p, f32p]
    lib.tts_sync.argtypes = [vp]
    lib.tts_launch_count.argtypes = [vp]
    lib.tts_launch_count.restype = C.c_int64
    lib.tts_last_stage_ms.argtypes = [vp]
    lib.tts_last_stage_ms.restype = C.c_float
    lib.tts_device_ms_total.argtypes = [vp]
    lib.tts_device_ms_total.restype = C.c_double
    lib.tts_nccl_unique_id.argtypes = [C.c_char_p]
    lib.tts_group_init_rank.argtypes = [vp, i32, i32, C.c_char_p, P(vp)]
    lib.tts_group_init_local.argtypes = [P(vp), i32, P(vp)]
    lib.tts_gather_select.argtypes = [vp, f32p, i32p, i32, i32p, f32p, i32p]
    lib.tts_group_last_error.argtypes = [vp]
    lib.tts_group_last_error.restype = C.c_char_p
    lib.tts_group_free.argtypes = [vp]
    lib.tts_group_free.restype = None
    lib.tts_bench_decode_step.argtypes = [vp, i32, P(C.c_float), P(C.c_double)]
    lib.tts_bench_gemv.argtypes = [vp, i32, i32, i32, P(C.c_float), P(C.c_double)]
    lib.tts_bench_conv3.argtypes = [vp, i32, i32, P(C.c_float), P(C.c_double)]
    _lib = lib
    return lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int32))


class Engine:
    """One context == one GPU.  Method names mirror the C-ABI entry points."""

    def __init__(self, device: int = 0, dtype: int = DTYPE_F32, max_batch: int = 4, max_positions: int = 404,
                 parity_quirks: bool = True):
        self.lib = load_library()
        cfg = TTSConfig(device=device, dtype=dtype, max_batch=max_batch, max_positions=max_positions,
                        parity_quirks=1 if parity_quirks else 0)
        h = C.c_void_p()
        rc = self.lib.tts_init(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise TTSError(rc, self.lib.tts_last_error(None).decode())
        self.h = h
        self.B = 0
        self.max_batch = max_batch

    def _chk(self, rc):
        if rc != 0:
            raise TTSError(rc, self.lib.tts_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.tts_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- loaders
    def load_ar(self, path):
        self._chk(self.lib.tts_load_ar(self.h, os.fsencode(path)))

    def load_diffusion(self, path):
        self._chk(self.lib.tts_load_diffusion(self.h, os.fsencode(path)))

    def load_vocoder(self, path):
        self._chk(self.lib.tts_load_vocoder(self.h, os.fsencode(path)))

    # ---- AR
    def ar_prefill(self, text, voice, B):
        text, tp = _i32(text)
        voice, vp = _f32(voice)
        out = np.empty((B, MEL_VOCAB), dtype=np.float32)
        self._chk(self.lib.tts_ar_prefill(self.h, tp, len(text), vp, B, out.ctypes.data_as(C.POINTER(C.c_float))))
        self.B = B
        return out

    def ar_prefill_multi(self, texts, voice):
        """U different prompts, one per candidate slot (utterance-batched decode); returns first-step logits [U][8194]"""
        U = len(texts)
        arrs = [np.ascontiguousarray(t, dtype=np.int32) for t in texts]
        i32p = C.POINTER(C.c_int32)
        ptrs = (i32p * U)(*[a.ctypes.data_as(i32p) for a in arrs])
        T = np.array([len(a) for a in arrs], dtype=np.int32)
        voice, vp = _f32(voice)
        out = np.empty((U, MEL_VOCAB), dtype=np.float32)
        self.lib.tts_ar_prefill_multi.argtypes = [C.c_void_p, C.c_int32, C.POINTER(i32p), i32p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        self._chk(self.lib.tts_ar_prefill_multi(self.h, U, ptrs, T.ctypes.data_as(i32p), vp, out.ctypes.data_as(C.POINTER(C.c_float))))
        self.B = U
        return out

    def ar_step(self, tokens, pos_id):
        tokens, tp = _i32(tokens)
        out = np.empty((self.B, MEL_VOCAB), dtype=np.float32)
        self._chk(self.lib.tts_ar_step(self.h, tp, pos_id, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def ar_step_topk(self, tokens, pos_id):
        """(values [B][64], indices [B][64]) of the device-side top-k pre-selection; raises if a row overflowed"""
        tokens, tp = _i32(tokens)
        vals = np.empty((self.B, 64), dtype=np.float32)
        idx = np.empty((self.B, 64), dtype=np.int32)
        flags = np.empty(self.B, dtype=np.int32)
        self._chk(self.lib.tts_ar_step_topk(self.h, tp, pos_id, vals.ctypes.data_as(C.POINTER(C.c_float)),
                                            idx.ctypes.data_as(C.POINTER(C.c_int32)), flags.ctypes.data_as(C.POINTER(C.c_int32))))
        self.topk_flags = flags
        return vals, idx

    def ar_last_logits(self):
        out = np.empty((self.B, MEL_VOCAB), dtype=np.float32)
        self._chk(self.lib.tts_ar_logits(self.h, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def ar_step_dev(self, tokens, pos_id):
        tokens, tp = _i32(tokens)
        dev = C.c_void_p()
        self._chk(self.lib.tts_ar_step_dev(self.h, tp, pos_id, C.byref(dev)))
        return dev.value

    def ar_latents(self, text, voice, codes, n_keep=500):
        text, tp = _i32(text)
        voice, vp = _f32(voice)
        codes, cp = _i32(codes)
        B = codes.shape[0]
        assert codes.shape[1] == 502
        out = np.empty((B, 500, 1024), dtype=np.float32)
        self._chk(self.lib.tts_ar_latents(self.h, tp, len(text), vp, cp, B, n_keep,
                                          out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    # ---- diffusion / vocoder
    def diffusion_eps(self, latents, x, timestep, conditioning_free):
        latents, lp = _f32(latents)
        x, xp = _f32(x)
        L, S = latents.shape[0], x.shape[1]
        out = np.empty((200, S), dtype=np.float32)
        self._chk(self.lib.tts_diffusion_eps(self.h, lp, L, xp, S, timestep, 1 if conditioning_free else 0,
                                             out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def diffusion_sample(self, latents, S, n_steps, noise):
        latents, lp = _f32(latents)
        noise, np_ = _f32(noise)
        assert noise.size == (n_steps + 1) * 100 * S
        out = np.empty((100, S), dtype=np.float32)
        self._chk(self.lib.tts_diffusion_sample(self.h, lp, latents.shape[0], S, n_steps, np_,
                                                out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def vocoder(self, mel, noise):
        mel, mp = _f32(mel)
        noise, np_ = _f32(noise)
        S = mel.shape[1]
        assert noise.size == (S + 10) * 64
        out = np.empty((S + 10) * 256 - 6, dtype=np.float32)
        self._chk(self.lib.tts_vocoder(self.h, mp, S, np_, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    # ---- introspection
    def sync(self):
        self._chk(self.lib.tts_sync(self.h))

    @property
    def launch_count(self):
        return int(self.lib.tts_launch_count(self.h))

    @property
    def last_stage_ms(self):
        return float(self.lib.tts_last_stage_ms(self.h))

    @property
    def device_ms_total(self):
        return float(self.lib.tts_device_ms_total(self.h))

    def bench_decode_step(self, iters):
        ms, by = C.c_float(), C.c_double()
        self._chk(self.lib.tts_bench_decode_step(self.h, iters, C.byref(ms), C.byref(by)))
        return ms.value, by.value

    def bench_conv3(self, S, iters):
        ms, fl = C.c_float(), C.c_double()
        self._chk(self.lib.tts_bench_conv3(self.h, S, iters, C.byref(ms), C.byref(fl)))
        return ms.value, fl.value

    def bench_gemv(self, op, B, iters):
        ms, by = C.c_float(), C.c_double()
        self._chk(self.lib.tts_bench_gemv(self.h, op, B, iters, C.byref(ms), C.byref(by)))
        return ms.value, by.value


def nccl_unique_id() -> bytes:
    """128 opaque bytes for tts_group_init_rank (rank 0 creates them, the launcher distributes them)"""
    buf = C.create_string_buffer(128)
    rc = load_library().tts_nccl_unique_id(buf)
    if rc != 0:
        raise TTSError(rc, "tts_nccl_unique_id failed (libnccl.so.2 missing?)")
    return buf.raw


class Group:
    """Candidate gather / selection over NCCL (include/tortoise_b200.h, csrc/dist.cu)."""

    def __init__(self, engines, rank=None, world=None, unique_id=None):
        self.lib = load_library()
        h = C.c_void_p()
        if rank is None:  # one process, several GPUs
            arr = (C.c_void_p * len(engines))(*[e.h for e in engines])
            rc = self.lib.tts_group_init_local(arr, len(engines), C.byref(h))
            self.n_local, self.world = len(engines), len(engines)
        else:  # one process per GPU
            rc = self.lib.tts_group_init_rank(engines[0].h, rank, world, unique_id, C.byref(h))
            self.n_local, self.world = 1, world
        if rc != 0:
            raise TTSError(rc, "NCCL group initialisation failed")
        self.h = h

    def gather_select(self, scores, lens):
        """scores / lens: [n_local][per] -> (winner global index, scores_all [world][per], lens_all)"""
        scores = np.ascontiguousarray(scores, dtype=np.float32).reshape(self.n_local, -1)
        lens = np.ascontiguousarray(lens, dtype=np.int32).reshape(self.n_local, -1)
        per = scores.shape[1]
        win = C.c_int32()
        sa = np.empty((self.world, per), dtype=np.float32)
        la = np.empty((self.world, per), dtype=np.int32)
        rc = self.lib.tts_gather_select(self.h, scores.ctypes.data_as(C.POINTER(C.c_float)),
                                        lens.ctypes.data_as(C.POINTER(C.c_int32)), per, C.byref(win),
                                        sa.ctypes.data_as(C.POINTER(C.c_float)), la.ctypes.data_as(C.POINTER(C.c_int32)))
        if rc != 0:
            raise TTSError(rc, self.lib.tts_group_last_error(self.h).decode())
        return win.value, sa, la

    def close(self):
        if getattr(self, "h", None):
            self.lib.tts_group_free(self.h)
            self.h = None
