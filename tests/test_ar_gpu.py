"""GPU parity tests of the AR stage, through the C-ABI (ctypes), against
 (a) golden tensors produced by the UNMODIFIED reference (oracle/_ref) on the same
     synthetic weights / prompt / voice / seed, and
 (b) the numpy oracle on the same inputs.

Tolerances.  The reference's AR graph rounds q/k/v and the GELU through fp16 (SURVEY A-2,
A-3), which makes logits chaotic at the 1e-3 level: perturbing the voice latent by 1e-7
relative moves the numpy oracle's own logits by 8.5e-4 max-abs (measured while pinning
the oracle).  We therefore accept max-abs 2.5e-3 on logits (|logit| <= ~3) and the
reference's own bar of 1e-2 on latents (main.cpp:6183-6209).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LOGIT_TOL = 2.5e-3
LATENT_TOL = 1e-2


def _codes_to_steps(codes500):
    """Sampled tokens of a candidate = codes up to and including the first 8193."""
    out = []
    for c in codes500:
        out.append(int(c))
        if c == 8193:
            break
    return out


def test_prefill_and_teacher_forced_decode_b1(engine_f32, golden, voice):
    g = golden("ar_b1.npz")
    text = g["tokens"]
    lg = engine_f32.ar_prefill(text, voice, 1)
    assert np.abs(lg[0] - g["logits_0"]).max() < LOGIT_TOL
    toks = _codes_to_steps(g["codes500"])
    steps = {int(k.split("_")[1]) for k in g.files if k.startswith("logits_")}
    assert steps == set(range(len(toks)))  # every decode step of the reference run is pinned
    for i, t in enumerate(toks):
        lg = engine_f32.ar_step([t], i + 2)  # reference: fixed_position = i + 2 (main.cpp:5227)
        if (i + 1) in steps:
            err = np.abs(lg[0] - g[f"logits_{i + 1}"]).max()
            assert err < LOGIT_TOL, f"step {i + 1}: {err}"


def test_prefill_matches_numpy_oracle_b1(engine_f32, golden, voice, model_dir):
    import _pkg
    import tortoise_oracle as O
    sw = _pkg.import_sub("synth_weights")
    W = sw.read_container(os.path.join(model_dir, "ggml-model.bin"))
    g = golden("ar_b1.npz")
    text = g["tokens"]
    ar = O.AROracle(W)
    ref = ar.prefill(text, voice, 1)
    lg = engine_f32.ar_prefill(text, voice, 1)
    assert np.abs(lg - ref).max() < LOGIT_TOL
    tok = [int(g["codes500"][0])]
    ref = ar.step(np.array(tok), 2)
    lg = engine_f32.ar_step(tok, 2)
    assert np.abs(lg - ref).max() < LOGIT_TOL


def test_teacher_forced_decode_b4(engine_f32, golden, voice):
    g = golden("ar_b4.npz")
    text = g["tokens"]
    lg = engine_f32.ar_prefill(text, voice, 4)
    assert lg.shape == (4, 8194)
    assert np.abs(lg - g["logits_0"]).max() < LOGIT_TOL
    # all four candidates see identical prefill inputs -> identical rows
    assert np.abs(lg - lg[0:1]).max() == 0.0
    # what the reference fed back each step: the candidate's sampled token; a candidate that
    # already stopped keeps being fed its own samples (main.cpp:5208-5217) -- with the
    # absorbing stop token of the synthetic weights that is 8193; we only compare live ones.
    seqs = [_codes_to_steps(g[f"codes500_{b}"]) for b in range(4)]
    n_steps = max(len(s) for s in seqs)
    for i in range(n_steps):
        fed = [seqs[b][i] if i < len(seqs[b]) else 8193 for b in range(4)]
        live = [b for b in range(4) if i < len(seqs[b])]
        lg = engine_f32.ar_step(fed, i + 2)
        key = f"logits_{i + 1}"
        if key in g.files:
            err = np.abs(lg[live] - g[key][live]).max()
            assert err < LOGIT_TOL, f"step {i + 1}: {err}"


def test_latents_b1_with_reference_position_quirk(engine_f32, golden, voice):
    from tortoise_oracle import apply_padding, trim_count
    g = golden("ar_b1.npz")
    codes500 = [int(c) for c in g["codes500"]]
    n_keep = trim_count(codes500)
    assert n_keep * 1024 == g["trimmed_latents"].size
    codes502 = np.array([[8192] + codes500 + [8193]], dtype=np.int32)
    lat = engine_f32.ar_latents(g["tokens"], voice, codes502, n_keep=n_keep)
    ref = g["trimmed_latents"].reshape(n_keep, 1024)
    assert np.abs(lat[0, :n_keep] - ref).max() < LATENT_TOL
    assert np.all(lat[0, n_keep:] == 0)


def test_latents_b4(engine_f32, golden, voice):
    from tortoise_oracle import trim_count
    g = golden("ar_b4.npz")
    codes502 = np.stack([np.concatenate([[8192], g[f"codes500_{b}"], [8193]]) for b in range(4)]).astype(np.int32)
    keeps = [trim_count([int(c) for c in g[f"codes500_{b}"]]) for b in range(4)]
    lat = engine_f32.ar_latents(g["tokens"], voice, codes502, n_keep=max(keeps))
    for b in range(4):
        ref = g[f"trimmed_latents_{b}"].reshape(keeps[b], 1024)
        err = np.abs(lat[b, :keeps[b]] - ref).max()
        assert err < LATENT_TOL, f"candidate {b}: {err}"


def test_truncated_latent_pass_is_exact(engine_f32, golden, voice):
    """n_keep truncation is exact by causality: same rows from a longer pass."""
    g = golden("ar_b1.npz")
    codes502 = np.array([[8192] + [int(c) for c in g["codes500"]] + [8193]], dtype=np.int32)
    a = engine_f32.ar_latents(g["tokens"], voice, codes502, n_keep=20)
    b = engine_f32.ar_latents(g["tokens"], voice, codes502, n_keep=40)
    assert np.array_equal(a[0, :20], b[0, :20])


def test_fp16_weight_mode_tracks_oracle(pkg, golden, voice, model_dir):
    """fast mode: f16 weight streaming, f32 activations/accumulation."""
    import _pkg
    import tortoise_oracle as O
    sw = _pkg.import_sub("synth_weights")
    W = sw.read_container(os.path.join(model_dir, "ggml-model.bin"))
    g = golden("ar_b1.npz")
    eng = pkg.Engine(device=0, dtype=pkg.DTYPE_F16, max_batch=2, max_positions=128)
    try:
        eng.load_ar(os.path.join(model_dir, "ggml-model.bin"))
        ar = O.AROracle(W, weight_dtype="f16")
        ref = ar.prefill(g["tokens"], voice, 2)
        lg = eng.ar_prefill(g["tokens"], voice, 2)
        assert np.abs(lg - ref).max() < LOGIT_TOL
        toks = [int(g["codes500"][0]), int(g["codes500"][1])]
        ref = ar.step(np.array(toks), 2)
        lg = eng.ar_step(toks, 2)
        assert np.abs(lg - ref).max() < LOGIT_TOL
    finally:
        eng.close()


@pytest.mark.parametrize("B", [1, 2, 4, 6])
def test_fp16_decode_tensor_core_path_vs_cuda_core_path_and_oracle(pkg, golden, voice, model_dir, B):
    """f16 weights: the tensor-core persistent step (ar_mega3.cuh, split-f16 activations) against the
    CUDA-core one (ar_mega2.cuh, TTS_MEGA_V2=1) over a teacher-forced run, and against the oracle.
    B = 6 runs as two launches per step (4 + 2 candidates)."""
    import _pkg
    import tortoise_oracle as O
    sw = _pkg.import_sub("synth_weights")
    W = sw.read_container(os.path.join(model_dir, "ggml-model.bin"))
    g = golden("ar_b1.npz")
    codes = [int(x) for x in g["codes500"][:12]]
    outs = {}
    for path in ("v3", "v2"):
        if path == "v2":
            os.environ["TTS_MEGA_V2"] = "1"
        try:
            eng = pkg.Engine(device=0, dtype=pkg.DTYPE_F16, max_batch=max(B, 4), max_positions=128)
        finally:
            os.environ.pop("TTS_MEGA_V2", None)
        try:
            eng.load_ar(os.path.join(model_dir, "ggml-model.bin"))
            eng.ar_prefill(g["tokens"], voice, B)
            lgs = []
            for i, c in enumerate(codes):
                toks = [(c + 7 * b) % 8192 for b in range(B)]
                lgs.append(eng.ar_step(toks, i + 2).copy())
            outs[path] = np.stack(lgs)
        finally:
            eng.close()
    assert np.isfinite(outs["v3"]).all()
    assert np.abs(outs["v3"] - outs["v2"]).max() < LOGIT_TOL
    ar = O.AROracle(W, weight_dtype="f16")
    ar.prefill(g["tokens"], voice, B)
    for i in range(3):
        toks = np.array([(codes[i] + 7 * b) % 8192 for b in range(B)])
        ref = ar.step(toks, i + 2)
        assert np.abs(outs["v3"][i] - ref).max() < LOGIT_TOL


def test_fp16_decode_long_context_walks_key_tiles(pkg, golden, voice, model_dir):
    """> 128 and > 256 cached positions: the attention item of the persistent step walks several
    128-key tiles with running (max, sum, acc).  Checked against the per-op path (TTS_NO_MEGA=1:
    wsgemv_kernel + ar_attn_decode_kernel, an independent attention implementation on the same f16
    weights) over a 290-step teacher-forced run; the numpy oracle needs 2.7 s per step."""
    g = golden("ar_b1.npz")
    n_steps = 290
    rs = np.random.RandomState(5)
    toks = rs.randint(0, 8192, size=n_steps)
    outs = {}
    for path in ("mega", "per_op"):
        if path == "per_op":
            os.environ["TTS_NO_MEGA"] = "1"
        try:
            eng = pkg.Engine(device=0, dtype=pkg.DTYPE_F16, max_batch=1, max_positions=404)
        finally:
            os.environ.pop("TTS_NO_MEGA", None)
        try:
            eng.load_ar(os.path.join(model_dir, "ggml-model.bin"))
            eng.ar_prefill(g["tokens"], voice, 1)
            lgs = []
            for i in range(n_steps):
                lgs.append(eng.ar_step([int(toks[i])], 2 + i).copy())
            outs[path] = np.stack(lgs)
        finally:
            eng.close()
    assert np.isfinite(outs["mega"]).all()
    err = np.abs(outs["mega"] - outs["per_op"]).reshape(n_steps, -1).max(axis=1)
    # 18 prefill positions + step + 1 keys: steps 108..112 straddle 128 keys, 236..240 straddle 256
    assert err.max() < LOGIT_TOL, (int(err.argmax()), float(err.max()))


def test_limits_are_errors_not_aborts(engine_f32, golden, voice, pkg):
    g = golden("ar_b1.npz")
    with pytest.raises(pkg.TTSError):
        engine_f32.ar_prefill(g["tokens"], voice, 5)  # > max_batch
    with pytest.raises(pkg.TTSError):
        engine_f32.ar_prefill(np.array([255, 999, 0]), voice, 1)  # token out of range
    engine_f32.ar_prefill(g["tokens"], voice, 1)
    with pytest.raises(pkg.TTSError):
        engine_f32.ar_step([9000], 2)
