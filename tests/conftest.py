import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

GOLDEN = os.path.join(ROOT, "tests", "golden")
# synthetic weights are generated once per machine into this cache (2.3 GB, ~30 s)
MODEL_DIR = os.environ.get("TTS_MODEL_DIR", "/tmp/tortoise_b200_models")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def pkg():
    import _pkg
    return _pkg.import_pkg()


@pytest.fixture(scope="session")
def model_dir():
    """Seeded synthetic weight files + the reference's tokenizer / voice fixtures."""
    import _pkg
    sw = _pkg.import_sub("synth_weights")
    digests = sw.generate(MODEL_DIR)
    import json
    import shutil
    with open(os.path.join(GOLDEN, "weights_digest.json")) as f:
        want = json.load(f)
    assert digests == want, "synthetic weights differ from the ones the golden fixtures were made with"
    for f in ("tokenizer.json", "mol.bin"):
        shutil.copyfile(os.path.join(GOLDEN, "models", f), os.path.join(MODEL_DIR, f))
    return MODEL_DIR


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def engine_f32(pkg, model_dir):
    eng = pkg.Engine(device=0, dtype=pkg.DTYPE_F32, max_batch=4, max_positions=404, parity_quirks=True)
    eng.load_ar(os.path.join(model_dir, "ggml-model.bin"))
    eng.load_diffusion(os.path.join(model_dir, "ggml-diffusion-model.bin"))
    eng.load_vocoder(os.path.join(model_dir, "ggml-vocoder-model.bin"))
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def voice():
    return np.fromfile(os.path.join(GOLDEN, "models", "mol.bin"), dtype=np.float32)


@pytest.fixture(scope="session")
def hostlib_full():
    import _pkg
    return _pkg.import_sub("host").HostLib(full=True)


def nmse(a, b):
    """normalised mean squared error, the metric of ggml's own cross-backend test
    (ggml/tests/test-backend-ops.cpp:193-206)."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-30))
