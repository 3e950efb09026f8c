#!/usr/bin/env python
"""export_weights.py -- tortoise-tts PyTorch checkpoints -> the three container files the reference
(and this engine) load: ggml-model.bin, ggml-diffusion-model.bin, ggml-vocoder-model.bin.

The reference promises this converter but does not ship it (README.md:34,77); SURVEY.md §8(f) #1.
The container format is App. B of SURVEY.md (main.cpp:494-500 magic, 811-888 records); the tensor
names and shapes every file must hold are `tortoise.cpp_b200/weights_manifest.json`, which was
produced by the reference loaders themselves (`oracle/_ref/ref_harness shapes`), so an export that
validates against it is loadable by `autoregressive_model_load` / `diffusion_model_load` /
`vocoder_model_load` (main.cpp:482, 931, 1665) and by `tts_load_*`.

What has to happen on the way (all checked by tests/test_export_weights_cpu.py on synthetic
state dicts; the real checkpoints are not available offline):

  autoregressive.pth (UnifiedVoice):
    gpt.*                    -> inference_model.transformer.*     (GPT2InferenceModel aliases the trunk)
    final_norm.*             -> inference_model.lm_head.0.*
    mel_head.*               -> inference_model.lm_head.1.*
    text_embedding / mel_embedding / *_pos_embedding.emb          unchanged
    (GPT-2 Conv1D weights stay [in][out]: the reference transposes them in the graph, main.cpp:2769)
    conditioning_encoder.*, text_head.*, attention masks: not part of the file (voices are pre-baked
    1024-float latents, main.cpp:5179)
  diffusion_decoder.pth (DiffusionTts):
    1x1 Conv1d weights [O, I, 1] are squeezed to [O, I] (the loader compares ne[0], ne[1],
    main.cpp:1585-1592); unconditioned_embedding [1, 1024, 1] -> [1024];
    diffusion_conditioning_latent [1, 2048] is NOT in the checkpoint (it is computed per voice by
    get_conditioning): pass it with --cond-latent (.npy / .pth / raw f32)
  vocoder.pth (UnivNet generator, optionally under the key "model_g"):
    weight_norm is folded: w = v * g / ||v|| (norm over all dims but 0), both the legacy
    weight_g / weight_v and the parametrizations.weight.original0 / original1 spellings;
    conv_post.1.weight [1, 32, 7] -> [32, 7]

usage: export_weights.py --out models/ [--autoregressive a.pth] [--diffusion d.pth --cond-latent c.npy]
                         [--vocoder v.pth]
"""
from __future__ import annotations

import argparse
import json
import os
import struct
import sys

import numpy as np

MAGIC = 0x67676D6C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MANIFEST = os.path.join(ROOT, "tortoise.cpp_b200", "weights_manifest.json")


def manifest() -> dict:
    with open(MANIFEST) as f:
        return json.load(f)


def _np(t) -> np.ndarray:
    """torch tensor / ndarray -> float32 ndarray (no torch import unless needed)."""
    if isinstance(t, np.ndarray):
        return np.ascontiguousarray(t, dtype=np.float32)
    return np.ascontiguousarray(t.detach().cpu().float().numpy())


def fold_weight_norm(sd: dict) -> dict:
    """Replace (X.weight_g, X.weight_v) or (X.parametrizations.weight.original0/1) by X.weight."""
    out = {}
    done = set()
    for k in sd:
        for g_suffix, v_suffix in ((".weight_g", ".weight_v"),
                                   (".parametrizations.weight.original0", ".parametrizations.weight.original1")):
            if k.endswith(g_suffix):
                base = k[: -len(g_suffix)]
                g, v = _np(sd[k]), _np(sd[base + v_suffix])
                axes = tuple(range(1, v.ndim))
                norm = np.sqrt((v.astype(np.float64) ** 2).sum(axis=axes, keepdims=True))
                out[base + ".weight"] = (v * (g.reshape(norm.shape) / norm)).astype(np.float32)
                done.update((k, base + v_suffix))
    for k in sd:
        if k not in done:
            out[k] = _np(sd[k])
    return out


def map_autoregressive(sd: dict) -> dict:
    out = {}
    for k, v in sd.items():
        if k.startswith("inference_model."):
            out[k] = _np(v)
        elif k.startswith("gpt."):
            out.setdefault("inference_model.transformer." + k[len("gpt."):], _np(v))
        elif k.startswith("final_norm."):
            out.setdefault("inference_model.lm_head.0." + k[len("final_norm."):], _np(v))
        elif k.startswith("mel_head."):
            out.setdefault("inference_model.lm_head.1." + k[len("mel_head."):], _np(v))
        else:
            out[k] = _np(v)
    return out


def map_diffusion(sd: dict, cond_latent) -> dict:
    out = {k: _np(v) for k, v in sd.items()}
    if cond_latent is not None:
        out["diffusion_conditioning_latent"] = _np(cond_latent).reshape(1, -1)  # [1, 2048]; checked against the manifest
    return out


def map_vocoder(sd: dict) -> dict:
    if "model_g" in sd and isinstance(sd["model_g"], dict):
        sd = sd["model_g"]
    return fold_weight_norm(sd)


def fit(name: str, a: np.ndarray, ne: list[int]) -> np.ndarray:
    """Bring a PyTorch-order tensor to the manifest's shape (reversed ne): only size-1 dims may be
    dropped (1x1 conv kernels, the leading 1 of conv_post / unconditioned_embedding)."""
    want = tuple(reversed(ne))
    if a.shape == want:
        return a
    squeezed = tuple(d for d in a.shape if d != 1)
    if squeezed == tuple(d for d in want if d != 1) and int(np.prod(a.shape)) == int(np.prod(want)):
        return a.reshape(want)
    raise ValueError(f"{name}: checkpoint shape {a.shape} does not fit the loader's {want}")


def write_container(path: str, entries: list[dict], tensors: dict, verbose: bool = False) -> list[str]:
    """entries: manifest order.  Returns the names that were missing (the loaders leave missing tensors
    uninitialised, main.cpp:811-888, so a partial file is an error here)."""
    missing = [e["name"] for e in entries if e["name"] not in tensors]
    if missing:
        return missing
    fitted = [fit(e["name"], tensors[e["name"]], e["ne"]) for e in entries]  # every shape checked before a byte is written
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(struct.pack("<I", MAGIC))
        for e, a in zip(entries, fitted):
            name, ne = e["name"], e["ne"]
            nb = name.encode()
            f.write(struct.pack("<iii", len(ne), len(nb), 0))
            f.write(struct.pack("<%di" % len(ne), *ne))
            f.write(nb)
            f.write(np.ascontiguousarray(a, dtype="<f4").tobytes())
            if verbose:
                print(f"  {name} {list(a.shape)}")
    os.replace(tmp, path)
    return []


def load_checkpoint(path: str) -> dict:
    if path.endswith(".npz"):
        return dict(np.load(path))
    import torch
    sd = torch.load(path, map_location="cpu", weights_only=True)
    if isinstance(sd, dict) and "state_dict" in sd and isinstance(sd["state_dict"], dict):
        sd = sd["state_dict"]
    return sd


def load_latent(path: str):
    if path.endswith(".npy"):
        return np.load(path)
    if path.endswith(".pth") or path.endswith(".pt"):
        import torch
        t = torch.load(path, map_location="cpu", weights_only=True)
        if isinstance(t, (tuple, list)):  # tortoise voice files hold (autoregressive latent, diffusion latent)
            t = t[-1]
        return _np(t)
    return np.fromfile(path, dtype=np.float32)


def export(out_dir: str, autoregressive=None, diffusion=None, vocoder=None, cond_latent=None, verbose=False) -> dict:
    """Each argument is a state dict (or None to skip that file).  Returns {filename: path}."""
    man = manifest()
    os.makedirs(out_dir, exist_ok=True)
    jobs = (("ggml-model.bin", autoregressive, map_autoregressive),
            ("ggml-diffusion-model.bin", diffusion, lambda sd: map_diffusion(sd, cond_latent)),
            ("ggml-vocoder-model.bin", vocoder, map_vocoder))
    written = {}
    for fname, sd, mapper in jobs:
        if sd is None:
            continue
        tensors = mapper(sd)
        path = os.path.join(out_dir, fname)
        missing = write_container(path, man[fname], tensors, verbose)
        if missing:
            hint = " (pass --cond-latent)" if "diffusion_conditioning_latent" in missing else ""
            raise KeyError(f"{fname}: {len(missing)} tensors the loader needs are not in the checkpoint{hint}: "
                           + ", ".join(missing[:6]) + (" ..." if len(missing) > 6 else ""))
        unused = sorted(set(tensors) - {e["name"] for e in man[fname]})
        if verbose and unused:
            print(f"{fname}: {len(unused)} checkpoint tensors are not part of the file, e.g. {unused[:4]}")
        written[fname] = path
    return written


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--out", required=True)
    ap.add_argument("--autoregressive")
    ap.add_argument("--diffusion")
    ap.add_argument("--cond-latent", help="[2048] diffusion conditioning latent of the voice (.npy / .pth / raw f32)")
    ap.add_argument("--vocoder")
    ap.add_argument("-v", "--verbose", action="store_true")
    a = ap.parse_args()
    if not (a.autoregressive or a.diffusion or a.vocoder):
        ap.error("nothing to export")
    w = export(a.out,
               load_checkpoint(a.autoregressive) if a.autoregressive else None,
               load_checkpoint(a.diffusion) if a.diffusion else None,
               load_checkpoint(a.vocoder) if a.vocoder else None,
               load_latent(a.cond_latent) if a.cond_latent else None, a.verbose)
    for k, v in w.items():
        print(f"wrote {v} ({os.path.getsize(v) / 1e6:.1f} MB)")


if __name__ == "__main__":
    sys.exit(main())
