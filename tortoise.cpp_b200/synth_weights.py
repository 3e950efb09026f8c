"""Seeded synthetic weight files in the reference's container format.

The three HuggingFace weight files the reference downloads (README.md:34) are not
available offline, so every parity / bench run uses these deterministic stand-ins.  The
container format is the one parsed by the reference loaders (main.cpp:494-500 magic,
main.cpp:811-888 records):

    uint32 magic 0x67676d6c, then until EOF:
    int32 n_dims, int32 name_len, int32 ttype(0 = f32), int32 ne[n_dims] (ggml order,
    ne[0] fastest), char name[name_len], f32 data

Tensor names and shapes come from ``weights_manifest.json``, which was produced by asking
the reference loaders themselves (``oracle/_ref/ref_harness shapes``).

The same generator runs here (to feed the CPU reference that produces tests/golden) and
on the GPU box (to feed the CUDA engine), so both sides see bit-identical files.
"""
from __future__ import annotations

import hashlib
import json
import os
import struct

import numpy as np

MAGIC = 0x67676D6C
SEED = 1234
_HERE = os.path.dirname(os.path.abspath(__file__))
FILES = ("ggml-model.bin", "ggml-diffusion-model.bin", "ggml-vocoder-model.bin")

STOP_TOKEN = 8193


def manifest() -> dict:
    with open(os.path.join(_HERE, "weights_manifest.json")) as f:
        return json.load(f)


def _fan_in(name: str, ne: list[int]) -> int:
    """ne is ggml order (fastest first)."""
    if len(ne) == 1:
        return 1
    if "convt_pre" in name:  # ConvTranspose1d [K, out, in] in ne order -> fan-in = in*K/stride ~ in*2
        return ne[2] * 2
    if name.endswith("c_attn.weight") or name.endswith("c_proj.weight") or name.endswith("c_fc.weight"):
        return ne[1]  # GPT-2 Conv1D stores [in][out]: ne = [out, in]
    if len(ne) == 2:
        return ne[0]  # [out][in]: ne = [in, out]
    return ne[0] * ne[1]  # conv [K, IC, OC]


def _make_tensor(rng: np.random.Generator, fname: str, name: str, ne: list[int]) -> np.ndarray:
    n = int(np.prod(ne))
    shape = tuple(reversed(ne))  # row-major numpy shape

    def normal(std):
        return (rng.standard_normal(n, dtype=np.float32) * np.float32(std)).reshape(shape)

    is_norm = (
        ".ln_1." in name or ".ln_2." in name or ".ln_f." in name or "lm_head.0." in name
        or ".norm." in name or "code_norm." in name or "in_layers.0." in name
        or "out_layers.0." in name or name.startswith("out.0.")
    )
    if is_norm:
        if name.endswith("weight"):
            return (np.float32(1.0) + normal(0.1)).astype(np.float32)
        return normal(0.05)
    if "relative_attention_bias" in name:
        return normal(0.1)
    if name in ("diffusion_conditioning_latent",):
        return normal(0.3)
    if name in ("unconditioned_embedding",):
        return normal(1.0)
    if name.endswith("bias"):
        return normal(0.01)
    if fname == "ggml-model.bin":
        if "embedding" in name:
            return normal(0.02 if "pos" in name else 0.05)
        return normal(0.02)
    # diffusion / vocoder matrices: roughly variance preserving
    fi = _fan_in(name, ne)
    gain = 0.8
    # vocoder kernel predictor: its input is the denormalised mel (|values| up to 11.5) and its
    # outputs are the LVC kernels that multiply 96 taps; keep the sigmoid/tanh gate
    # pre-activations O(1) so the generator is well conditioned (with larger gains the
    # network amplifies 1e-6 input perturbations to 1e-3 output NMSE -- useless for parity).
    if "input_conv.0.weight" in name:
        gain = 0.15
    if "kernel_conv.weight" in name:
        gain = 0.08
    if "bias_conv.weight" in name:
        gain = 0.3
    return normal(gain / np.sqrt(fi))


def _engineer_ar(tensors: dict[str, np.ndarray], rng: np.random.Generator) -> None:
    """Make the stop token reachable and absorbing (SURVEY.md App. D hazards 1-2) so the
    reference's own stop rule (all candidates emit 8193 in the same step,
    main.cpp:5206-5222) terminates on random weights."""
    signs = np.where(rng.random(1024) < 0.5, -1.0, 1.0).astype(np.float32)
    tensors["mel_embedding.weight"][STOP_TOKEN, :] = 2.0 * signs
    tensors["inference_model.lm_head.1.weight"][STOP_TOKEN, :] = 0.03 * signs
    tensors["inference_model.lm_head.1.bias"][STOP_TOKEN] = 2.2


def generate(model_dir: str, force: bool = False, verbose: bool = False) -> dict:
    """Write the three files into model_dir (skipped when already present and complete).
    Returns {filename: sha256-of-first-MiB} for cross-machine consistency checks."""
    os.makedirs(model_dir, exist_ok=True)
    man = manifest()
    digests = {}
    for fi, fname in enumerate(FILES):
        path = os.path.join(model_dir, fname)
        entries = man[fname]
        expect = 4 + sum(12 + 4 * len(e["ne"]) + len(e["name"]) + 4 * int(np.prod(e["ne"])) for e in entries)
        if not force and os.path.exists(path) and os.path.getsize(path) == expect:
            with open(path, "rb") as f:
                digests[fname] = hashlib.sha256(f.read(1 << 20)).hexdigest()
            continue
        rng = np.random.default_rng(SEED + fi)
        tensors = {e["name"]: _make_tensor(rng, fname, e["name"], e["ne"]) for e in entries}
        if fname == "ggml-model.bin":
            _engineer_ar(tensors, rng)
        tmp = path + ".tmp"
        with open(tmp, "wb") as f:
            f.write(struct.pack("<I", MAGIC))
            for e in entries:
                name = e["name"].encode()
                ne = e["ne"]
                f.write(struct.pack("<iii", len(ne), len(name), 0))
                f.write(struct.pack("<%di" % len(ne), *ne))
                f.write(name)
                f.write(np.ascontiguousarray(tensors[e["name"]], dtype=np.float32).tobytes())
        os.replace(tmp, path)
        assert os.path.getsize(path) == expect
        with open(path, "rb") as f:
            digests[fname] = hashlib.sha256(f.read(1 << 20)).hexdigest()
        if verbose:
            print(f"wrote {path} ({expect} bytes)")
        del tensors
    return digests


def read_container(path: str) -> dict[str, np.ndarray]:
    """Parse a container file into {name: row-major ndarray (reversed ne)}."""
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    (magic,) = struct.unpack_from("<I", data, 0)
    if magic != MAGIC:
        raise ValueError("bad magic")
    off = 4
    while off < len(data):
        n_dims, name_len, ttype = struct.unpack_from("<iii", data, off)
        off += 12
        ne = struct.unpack_from("<%di" % n_dims, data, off)
        off += 4 * n_dims
        name = data[off:off + name_len].decode()
        off += name_len
        n = int(np.prod(ne))
        out[name] = np.frombuffer(data, dtype=np.float32, count=n, offset=off).reshape(tuple(reversed(ne)))
        off += 4 * n
    return out


if __name__ == "__main__":
    import sys
    d = generate(sys.argv[1], force="--force" in sys.argv, verbose=True)
    print(json.dumps(d, indent=1))
