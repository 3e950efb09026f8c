"""Pins the numpy oracle (oracle/tortoise_oracle.py) against tensors produced by the
UNMODIFIED reference (tests/golden/, see make_golden.py).  CPU only.

The oracle is the checker used by the GPU tests at other sizes/seeds; these tests are what
makes its parity claim "pinned" rather than self-referential."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, nmse


@pytest.fixture(scope="module")
def weights(model_dir):
    import _pkg
    sw = _pkg.import_sub("synth_weights")
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = sw.read_container(os.path.join(model_dir, name))
        return cache[name]
    return get


def test_ar_oracle_prefill_and_steps(weights, voice):
    import tortoise_oracle as O
    g = np.load(os.path.join(GOLDEN, "ar_b1.npz"))
    ar = O.AROracle(weights("ggml-model.bin"))
    lg = ar.prefill(g["tokens"], voice, 1)
    assert np.abs(lg[0] - g["logits_0"]).max() < 2.5e-3
    for i in range(2):
        lg = ar.step(np.array([g["codes500"][i]]), i + 2)
        assert np.abs(lg[0] - g[f"logits_{i + 1}"]).max() < 2.5e-3


def test_ar_oracle_latents_b1_quirk(weights, voice):
    import tortoise_oracle as O
    g = np.load(os.path.join(GOLDEN, "ar_b1.npz"))
    codes500 = [int(c) for c in g["codes500"]]
    n = O.trim_count(codes500)
    codes502 = np.array([[8192] + codes500 + [8193]])
    lat = O.AROracle(weights("ggml-model.bin")).latents(g["tokens"], voice, codes502, n_keep=n)
    assert np.abs(lat[0] - g["trimmed_latents"].reshape(n, 1024)).max() < 1e-2


def test_diffusion_oracle_single_passes(weights):
    import tortoise_oracle as O
    g = np.load(os.path.join(GOLDEN, "diffusion.npz"))
    d = O.DiffusionOracle(weights("ggml-diffusion-model.bin"))
    lat = g["latents"].reshape(-1, 1024)
    sched = O.ddpm_schedule(80)
    for k in (0, 1):
        out = d.eps(lat, g[f"x_{k}"], sched[k // 2]["timestep"], conditioning_free=bool(k % 2))
        assert np.abs(out - g[f"out_{k}"]).max() < 1e-2
        assert nmse(out, g[f"out_{k}"]) < 1e-5


def test_ddpm_step_oracle_reproduces_reference_trajectory():
    """x_2 (what the reference fed to its 3rd graph run) from x_0, the two model outputs of
    step 0 and the seed-0 noise stream: pins schedule, CFG blend, variance swap, clamp."""
    import _pkg
    import tortoise_oracle as O
    g = np.load(os.path.join(GOLDEN, "diffusion.npz"))
    S = g["x_0"].shape[1]
    rng = _pkg.import_sub("host").HostLib().rng(0)
    noise = rng.normal(2 * 100 * S).reshape(2, 100, S)
    assert np.array_equal(noise[0], g["x_0"])
    sched = O.ddpm_schedule(80)
    x1 = O.ddpm_step(g["x_0"], g["out_0"], g["out_1"], noise[1], sched[0])
    assert np.abs(x1 - g["x_2"]).max() < 2e-6
    # the C++ host schedule (product code) equals the oracle's
    hs = _pkg.import_sub("host").HostLib().ddpm_schedule(80)
    for i, k in enumerate(sched):
        row = [k["cfk"], k["sqrt_recip"], k["sqrt_recipm1"], k["coef1"], k["coef2"], k["min_log"], k["max_log"]]
        assert np.allclose(hs[i, :7], np.array(row, dtype=np.float32), rtol=1e-6, atol=0), i
        assert int(hs[i, 8]) == k["timestep"]


def test_200_step_schedule_against_patched_reference():
    """--steps 200 (BASELINE configs[3]): first and last sampling step of the "patched reference"
    (oracle/patch_steps.py: the 80 / 79 literals become a run-time step count) reproduced from its own
    model outputs -- pins the 200-entry timestep map, the respaced betas and both ends of the schedule
    for the oracle AND the C++ host code."""
    import _pkg
    import tortoise_oracle as O
    g = np.load(os.path.join(GOLDEN, "diffusion200.npz"))
    S = g["x_0"].shape[1]
    hl = _pkg.import_sub("host").HostLib()
    noise = hl.rng(0).normal(2 * 100 * S).reshape(2, 100, S)
    assert np.array_equal(noise[0], g["x_0"])
    sched = O.ddpm_schedule(200)
    x1 = O.ddpm_step(g["x_0"], g["out_0"], g["out_1"], noise[1], sched[0])
    assert np.abs(x1 - g["x_2"]).max() < 2e-6
    assert sched[199]["last"] and sched[199]["timestep"] == 0
    mel = O.ddpm_step(g["x_398"], g["out_398"], g["out_399"], np.zeros_like(g["x_0"]), sched[199])
    assert np.abs(mel - g["mel"]).max() < 2e-6
    hs = hl.ddpm_schedule(200)
    tmap = hl.timestep_map(200)
    for i, k in enumerate(sched):
        row = [k["cfk"], k["sqrt_recip"], k["sqrt_recipm1"], k["coef1"], k["coef2"], k["min_log"], k["max_log"]]
        assert np.allclose(hs[i, :7], np.array(row, dtype=np.float32), rtol=1e-6, atol=0), i
        assert int(hs[i, 8]) == k["timestep"] == int(tmap[199 - i])
    assert int(g["t_0"]) == int(tmap[199]) and int(g["t_398"]) == int(tmap[0])
    # mid-trajectory sampler steps: the C++ host schedule (the PRODUCT code) reproduces the patched
    # reference's x bit for bit; the numpy oracle to float rounding
    names = ["cfk", "sqrt_recip", "sqrt_recipm1", "coef1", "coef2", "min_log", "max_log"]
    for k in (67, 100, 150):
        noise = hl.rng(0).normal((k + 2) * 100 * S).reshape(k + 2, 100, S)[k + 1]
        args = (g[f"x_{2 * k}"], g[f"out_{2 * k}"], g[f"out_{2 * k + 1}"], noise)
        assert np.abs(O.ddpm_step(*args, sched[k]) - g[f"x_{2 * k + 2}"]).max() < 2e-6
        hk = dict(sched[k])
        hk.update({n: np.float32(hs[k, j]) for j, n in enumerate(names)})
        assert np.array_equal(O.ddpm_step(*args, hk), g[f"x_{2 * k + 2}"]), k


def test_vocoder_oracle(weights):
    import tortoise_oracle as O
    g = np.load(os.path.join(GOLDEN, "vocoder.npz"))
    audio = O.VocoderOracle(weights("ggml-vocoder-model.bin")).run(g["mel"], g["noise"])
    assert audio.shape == g["audio"].shape
    assert np.abs(audio - g["audio"]).max() < 1e-2
    assert nmse(audio, g["audio"]) < 1e-6
