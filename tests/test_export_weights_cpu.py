"""tools/export_weights.py: PyTorch-side state dicts -> the reference's container files.
The real checkpoints are not available offline, so the test builds state dicts with the PyTorch-side
names / shapes (gpt.* aliases, trailing 1x1-conv dims, weight_norm pairs under "model_g", extra tensors
the file must not hold) from the loader-produced manifest with every dimension shrunk, exports them,
and reads the files back with the container parser the engine's tests use."""
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402

SHRINK = {1024: 16, 2048: 32, 3072: 48, 4096: 64, 8194: 34, 24576: 96, 608: 12, 404: 10, 256: 8, 200: 6, 100: 5,
          64: 4, 32: 4}


def _load_tool():
    spec = importlib.util.spec_from_file_location("export_weights", os.path.join(ROOT, "tools", "export_weights.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.fixture()
def tool(monkeypatch):
    m = _load_tool()
    full = m.manifest()
    small = {f: [{"name": e["name"], "ne": [SHRINK.get(d, d) for d in e["ne"]]} for e in es] for f, es in full.items()}
    # (the one shape the squeeze rule must not be fooled by: diffusion_conditioning_latent is [1, 2048])
    monkeypatch.setattr(m, "manifest", lambda: small)
    return m, small


def _torch_side(small, rs):
    """Invert the exporter's mapping: what the three PyTorch checkpoints would hold."""
    ar, expect_ar = {}, {}
    for e in small["ggml-model.bin"]:
        a = rs.standard_normal(tuple(reversed(e["ne"]))).astype(np.float32)
        expect_ar[e["name"]] = a
        n = e["name"]
        if n.startswith("inference_model.transformer."):
            n = "gpt." + n[len("inference_model.transformer."):]
        elif n.startswith("inference_model.lm_head.0."):
            n = "final_norm." + n[len("inference_model.lm_head.0."):]
        elif n.startswith("inference_model.lm_head.1."):
            n = "mel_head." + n[len("inference_model.lm_head.1."):]
        ar[n] = a
    ar["conditioning_encoder.init.weight"] = np.zeros((4, 4, 1), np.float32)  # not part of the file
    ar["text_head.weight"] = np.zeros((8, 16), np.float32)

    diff, expect_diff, latent = {}, {}, None
    for e in small["ggml-diffusion-model.bin"]:
        a = rs.standard_normal(tuple(reversed(e["ne"]))).astype(np.float32)
        expect_diff[e["name"]] = a
        n = e["name"]
        if n == "diffusion_conditioning_latent":
            latent = a.reshape(-1)
            continue
        if n == "unconditioned_embedding":
            diff[n] = a.reshape(1, -1, 1)
        elif a.ndim == 2 and (n.endswith("qkv.weight") or n.endswith("proj_out.weight") or n.endswith("in_layers.2.weight")
                              or n == "integrating_conv.weight"):
            diff[n] = a[:, :, None]  # Conv1d(kernel_size=1)
        else:
            diff[n] = a
    voc, expect_voc = {}, {}
    for e in small["ggml-vocoder-model.bin"]:
        shape = tuple(reversed(e["ne"]))
        n = e["name"]
        if n == "conv_post.1.weight":
            shape = (1,) + shape
        if n.endswith(".weight"):
            v = rs.standard_normal(shape).astype(np.float32)
            g = rs.uniform(0.5, 2.0, size=(shape[0],) + (1,) * (len(shape) - 1)).astype(np.float32)
            norm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=tuple(range(1, v.ndim)), keepdims=True))
            expect_voc[n] = (v * (g / norm)).astype(np.float32).reshape(tuple(reversed(e["ne"])))
            base = n[: -len(".weight")]
            if "conv_blocks" in n:  # the newer parametrization spelling
                voc[base + ".parametrizations.weight.original0"] = g
                voc[base + ".parametrizations.weight.original1"] = v
            else:
                voc[base + ".weight_g"] = g
                voc[base + ".weight_v"] = v
        else:
            a = rs.standard_normal(shape).astype(np.float32)
            expect_voc[n] = a
            voc[n] = a
    return ar, expect_ar, diff, expect_diff, latent, {"model_g": voc}, expect_voc


def test_export_round_trip(tool, tmp_path):
    m, small = tool
    sw = _pkg.import_sub("synth_weights")
    rs = np.random.default_rng(0)
    ar, e_ar, diff, e_diff, latent, voc, e_voc = _torch_side(small, rs)
    w = m.export(str(tmp_path), ar, diff, voc, cond_latent=latent)
    assert set(w) == set(small)
    for fname, expect in (("ggml-model.bin", e_ar), ("ggml-diffusion-model.bin", e_diff), ("ggml-vocoder-model.bin", e_voc)):
        got = sw.read_container(w[fname])
        assert list(got) == [e["name"] for e in small[fname]]  # manifest order, nothing extra
        for e in small[fname]:
            a = got[e["name"]]
            assert a.shape == tuple(reversed(e["ne"])), e["name"]
            np.testing.assert_allclose(a, expect[e["name"]], rtol=1e-6, atol=1e-7, err_msg=e["name"])


def test_export_errors_are_loud(tool, tmp_path):
    m, small = tool
    rs = np.random.default_rng(1)
    ar, _, diff, _, latent, voc, _ = _torch_side(small, rs)
    with pytest.raises(KeyError, match="cond-latent"):
        m.export(str(tmp_path), None, diff, None, cond_latent=None)        # latent is not in the checkpoint
    del ar["mel_head.weight"]
    with pytest.raises(KeyError, match="lm_head.1.weight"):
        m.export(str(tmp_path), ar, None, None)
    bad = dict(diff)
    k = "inp_block.weight"
    bad[k] = bad[k][:, :, :2]
    with pytest.raises(ValueError, match="inp_block.weight"):
        m.export(str(tmp_path), None, bad, None, cond_latent=latent)
    assert not os.listdir(tmp_path)                                        # nothing half-written


def test_full_manifest_names_cover_what_the_engine_loads():
    """the real (unshrunk) manifest is what the exporter validates against."""
    m = _load_tool()
    man = m.manifest()
    assert len(man["ggml-model.bin"]) == 370 and len(man["ggml-diffusion-model.bin"]) == 297
    assert len(man["ggml-vocoder-model.bin"]) == 88
    names = {e["name"] for e in man["ggml-model.bin"]}
    assert "inference_model.transformer.h.29.mlp.c_proj.weight" in names and "inference_model.lm_head.1.bias" in names
