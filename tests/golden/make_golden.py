"""Regenerates tests/golden/* from runs of the UNMODIFIED reference (oracle/_ref/ref_harness,
built from /root/reference by `make -C oracle`) on the seeded synthetic weights.

Only runnable in the build container (needs /root/reference); the fixtures it writes are
committed so the GPU box never needs the reference.  Steps (about 10 minutes of CPU):

    make -C oracle
    python tests/golden/make_golden.py --work /tmp/w            # runs everything missing
    python tests/golden/make_golden.py --work /tmp/w --assemble # only re-packs the .npz

Reference invocations (cwd = <work>/build, models in <work>/models, exactly like the
reference's own `cd build && ./tortoise` layout, main.cpp:5078/5625/6046/6551):
    ref_harness full "this is a test message." ../models/mol.bin 0 <work>/out_full
    ref_harness ar   "this is a test message." ../models/mol.bin 4 0 <work>/out_ar4
    ref_harness diff <work>/out_full/trimmed_latents_0.f32 0 <work>/out_diff
    ref_harness voc  <work>/out_diff/mel.f32 0 <work>/out_voc
    ref_harness hostfn <work>/out_host
    ref_harness sample ... (see below)
    ref_harness tokenize "<sentence>"   for every sentence of CORPUS
Round-2 additions (`--long`; ~35 minutes of CPU, only these two fixtures are rewritten):
    ref_harness ar "this is a test message." ../models/mol.bin 1 0 <work>/out_arlong -1 272
        forced-length run of the UNMODIFIED reference: the harness replaces the stop token's logit by
        -1e30 in transit for the first 272 read-backs (after dumping the true row), so the KV cache
        grows past 128 and 256 keys -> ar_long.npz (all fed tokens, logits of ~35 pinned steps)
    HX_STEPS=200 ref_harness_steps diff <latents of ar_b1.npz> 0 <work>/out_diff200
        "patched reference" (oracle/patch_steps.py: 80 / 79 literals + table -> run-time step count)
        -> diffusion200.npz (final mel, 4 teacher-forced passes)
"""
import argparse
import glob
import json
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLD = os.path.join(ROOT, "tests", "golden")
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
PROMPT = "this is a test message."

CORPUS = [
    "this is a test message.",
    "based... dr freeman?",
    "hello world",
    "the quick brown fox jumps over the lazy dog.",
    "congratulations! autoregressive model complete!",
    "it's a truth universally acknowledged, that a single man in possession of a good fortune, must be in want of a wife.",
    "what's up? nothing much - you?",
    "a",
    "zzz qqq xxx",
    "tortoise text to speech, on blackwell!",
    "means are important, as ends. crisis makes it tempting to ignore the wise restraints that make men free.",
    "we're they'll i'd you've",
    "UPPER case and digits 123 are skipped with a warning",
]


def run(cmd, cwd):
    print("+", " ".join(cmd))
    subprocess.check_call(cmd, cwd=cwd)


def f32(p):
    return np.fromfile(p, dtype=np.float32)


def i32(p):
    return np.fromfile(p, dtype=np.int32)


def make_long(W, build, assemble_only):
    steps_bin = os.path.join(ROOT, "oracle", "_ref", "ref_harness_steps")
    lat_path = W + "/lat_b1.f32"
    np.load(os.path.join(GOLD, "ar_b1.npz"))["trimmed_latents"].astype(np.float32).tofile(lat_path)
    if not assemble_only:
        if not os.path.exists(W + "/out_arlong/codes_0.i32"):
            run([HARNESS, "ar", PROMPT, "../models/mol.bin", "1", "0", W + "/out_arlong", "-1", "272"], build)
        if not os.path.exists(W + "/out_diff200/mel.f32"):
            env = dict(os.environ, HX_STEPS="200")
            print("+ HX_STEPS=200", steps_bin, "diff ...")
            subprocess.check_call([steps_bin, "diff", lat_path, "0", W + "/out_diff200"], cwd=build, env=env)
    # ---- forced-length AR run
    codes = i32(W + "/out_arlong/codes_0.i32")
    fed = []
    for c in codes:
        fed.append(int(c))
        if c == 8193:
            break
    logit_files = sorted(p for p in glob.glob(W + "/out_arlong/ar_get*_c*.f32") if os.path.getsize(p) == 8194 * 4)
    assert len(logit_files) == len(fed), (len(logit_files), len(fed))  # row k produced sample k
    T = len(i32(W + "/out_arlong/tokens.i32"))
    d = {"tokens": i32(W + "/out_arlong/tokens.i32"), "fed_tokens": np.array(fed, np.int32)}
    # pinned rows: n_keys of row k = T + 2 + k.  Around the 128- and 256-key tile boundaries every
    # step, plus a sparse sweep (a row is 32 KB)
    b128, b256 = 128 - T - 2, 256 - T - 2
    pinned = {0, 1, 2, 40, 80, 100, 140, 180, 220, len(fed) - 2, len(fed) - 1}
    pinned |= set(range(b128 - 3, b128 + 4)) | set(range(b256 - 3, b256 + 4))
    for k in sorted(x for x in pinned if 0 <= x < len(logit_files)):
        d[f"logits_{k}"] = f32(logit_files[k])
    np.savez(os.path.join(GOLD, "ar_long.npz"), **d)
    print("ar_long.npz:", len(fed), "fed tokens,", sum(k.startswith("logits_") for k in d), "pinned rows")
    # ---- 200 sampling steps, patched reference
    S = f32(W + "/out_diff200/mel.f32").size // 100
    outs = sorted(glob.glob(W + "/out_diff200/diff_get*_c*.f32"))
    xs = sorted(glob.glob(W + "/out_diff200/diff_set_noise_tensor_c*.f32"))
    assert len(outs) == 400 and len(xs) == 400, (len(outs), len(xs))
    from tortoise_oracle import ddpm_schedule  # noqa: F401  (timestep map cross-check below)
    frac, cur, tmap = 3999.0 / 199.0, 0.0, []
    for _ in range(200):
        tmap.append(int(round(cur)))
        cur += frac
    d = {"latents": f32(lat_path), "mel": f32(W + "/out_diff200/mel.f32").reshape(100, S)}
    for k in (0, 1, 398, 399):
        d[f"x_{k}"] = f32(xs[k]).reshape(100, S)
        d[f"out_{k}"] = f32(outs[k]).reshape(200, S)
        d[f"t_{k}"] = np.array(tmap[199 - k // 2])
    d["x_2"] = f32(xs[2]).reshape(100, S)  # x after the first sampling step (trajectory check on the CPU)
    for k in (67, 100, 150):  # mid-trajectory sampler steps: x before, both model outputs, x after
        d[f"x_{2 * k}"] = f32(xs[2 * k]).reshape(100, S)
        d[f"out_{2 * k}"] = f32(outs[2 * k]).reshape(200, S)
        d[f"out_{2 * k + 1}"] = f32(outs[2 * k + 1]).reshape(200, S)
        d[f"x_{2 * k + 2}"] = f32(xs[2 * k + 2]).reshape(100, S)
    np.savez(os.path.join(GOLD, "diffusion200.npz"), **d)
    print("diffusion200.npz: S =", S)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--work", default="/tmp/w")
    ap.add_argument("--assemble", action="store_true")
    ap.add_argument("--long", action="store_true", help="only the round-2 fixtures (ar_long.npz, diffusion200.npz)")
    a = ap.parse_args()
    W = a.work
    build = os.path.join(W, "build")
    models = os.path.join(W, "models")
    os.makedirs(build, exist_ok=True)
    import _pkg
    sw = _pkg.import_sub("synth_weights")
    digests = sw.generate(models, verbose=True)
    for f in ("tokenizer.json", "mol.bin"):
        if not os.path.exists(os.path.join(models, f)):
            shutil.copyfile(os.path.join("/root/reference/models", f), os.path.join(models, f))
    if a.long:
        make_long(W, build, a.assemble)
        return
    if not a.assemble:
        if not os.path.exists(W + "/out_full/audio.f32"):
            run([HARNESS, "full", PROMPT, "../models/mol.bin", "0", W + "/out_full"], build)
        if not os.path.exists(W + "/out_ar4/codes_3.i32"):
            run([HARNESS, "ar", PROMPT, "../models/mol.bin", "4", "0", W + "/out_ar4"], build)
        if not os.path.exists(W + "/out_diff/mel.f32"):
            run([HARNESS, "diff", W + "/out_full/trimmed_latents_0.f32", "0", W + "/out_diff"], build)
        if not os.path.exists(W + "/out_voc/audio.f32"):
            run([HARNESS, "voc", W + "/out_diff/mel.f32", "0", W + "/out_voc"], build)
        run([HARNESS, "hostfn", W + "/out_host"], build)

    os.makedirs(os.path.join(GOLD, "models"), exist_ok=True)
    for f in ("tokenizer.json", "mol.bin"):
        shutil.copyfile(os.path.join(models, f), os.path.join(GOLD, "models", f))
    with open(os.path.join(GOLD, "weights_digest.json"), "w") as f:
        json.dump(digests, f, indent=1)

    # ---- tokenizer corpus
    tok = {}
    for s in CORPUS:
        out = subprocess.run([HARNESS, "tokenize", s], cwd=build, capture_output=True, text=True, check=True).stdout
        line = [l for l in out.splitlines() if l and l[0].isdigit()][-1]
        tok[s] = [int(x) for x in line.split(",")]
    with open(os.path.join(GOLD, "tokenizer_corpus.json"), "w") as f:
        json.dump(tok, f, indent=1)

    # ---- AR, B = 1 (from the full run: seed 0)
    d = {"tokens": i32(W + "/out_full/tokens.i32"), "codes500": i32(W + "/out_full/codes_0.i32"),
         "trimmed_latents": f32(W + "/out_full/trimmed_latents_0.f32")}
    logit_files = sorted(glob.glob(W + "/out_full/ar_get*_c*.f32"))
    logit_files = [p for p in logit_files if os.path.getsize(p) == 8194 * 4]
    keep = range(len(logit_files))  # every step: the free-running comparison needs them all
    for k in keep:
        d[f"logits_{k}"] = f32(logit_files[k])
    d["n_logit_steps"] = np.array(len(logit_files))
    np.savez(os.path.join(GOLD, "ar_b1.npz"), **d)

    # ---- AR, B = 4
    d = {"tokens": i32(W + "/out_ar4/tokens.i32")}
    for b in range(4):
        d[f"codes500_{b}"] = i32(W + f"/out_ar4/codes_{b}.i32")
        d[f"trimmed_latents_{b}"] = f32(W + f"/out_ar4/trimmed_latents_{b}.f32")
    logit_files = sorted(glob.glob(W + "/out_ar4/ar_get*_c*.f32"))
    logit_files = [p for p in logit_files if os.path.getsize(p) == 4 * 8194 * 4]
    for k in sorted({0, 1, 5, len(logit_files) - 1}):
        d[f"logits_{k}"] = f32(logit_files[k]).reshape(4, 8194)
    np.savez(os.path.join(GOLD, "ar_b4.npz"), **d)

    # ---- sampler goldens through the reference's own process_logits_and_sample
    os.makedirs(W + "/out_sample", exist_ok=True)
    T = len(d["tokens"])
    lgA = f32(logit_files[0])
    prevA = np.array(([1] * (T + 1) + [8192]) * 4, np.int32)
    lgB = f32(logit_files[5])
    prevB = np.array([6031, 83, 8193, 1], np.int32)
    smp = {}
    for tag, lg, prev, seed in (("a", lgA, prevA, 0), ("b", lgB, prevB, 123)):
        lg.tofile(W + f"/out_sample/{tag}_logits.f32")
        prev.tofile(W + f"/out_sample/{tag}_prev.i32")
        run([HARNESS, "sample", W + f"/out_sample/{tag}_logits.f32", W + f"/out_sample/{tag}_prev.i32", "4",
             str(seed), "8", W + f"/out_sample/{tag}_out.i32"], build)
        smp[f"{tag}_logits"] = lg.reshape(4, 8194)
        smp[f"{tag}_prev"] = prev
        smp[f"{tag}_seed"] = np.array(seed)
        smp[f"{tag}_samples"] = i32(W + f"/out_sample/{tag}_out.i32").reshape(8, 4)
    np.savez(os.path.join(GOLD, "sampler.npz"), **smp)

    # ---- diffusion (standalone run, seed 0): teacher-forced passes + final mel
    S = f32(W + "/out_diff/mel.f32").size // 100
    d = {"latents": f32(W + "/out_full/trimmed_latents_0.f32"), "mel": f32(W + "/out_diff/mel.f32").reshape(100, S)}
    outs = sorted(glob.glob(W + "/out_diff/diff_get*_c*.f32"))
    xs = sorted(glob.glob(W + "/out_diff/diff_set_noise_tensor_c*.f32"))
    assert len(outs) == 160 and len(xs) == 160, (len(outs), len(xs))
    for k in (0, 1, 2, 3, 80, 81, 158, 159):
        d[f"x_{k}"] = f32(xs[k]).reshape(100, S)
        d[f"out_{k}"] = f32(outs[k]).reshape(200, S)
    np.savez(os.path.join(GOLD, "diffusion.npz"), **d)

    # ---- vocoder (standalone, seed 0)
    d = {"mel": f32(W + "/out_diff/mel.f32").reshape(100, S),
         "noise": f32(glob.glob(W + "/out_voc/voc_set_vocoder_noise_tensor_c*.f32")[0]),
         "audio": f32(W + "/out_voc/audio.f32")}
    np.savez(os.path.join(GOLD, "vocoder.npz"), **d)

    # ---- full pipeline, seed 0 (free running): final mel/audio of ./tortoise --seed 0
    d = {"mel": f32(W + "/out_full/mel.f32"), "audio": f32(W + "/out_full/audio.f32")}
    np.savez(os.path.join(GOLD, "full_seed0.npz"), **d)
    shutil.copyfile(W + "/out_full/output.wav", os.path.join(GOLD, "full_seed0_head.wav.tmp"))
    with open(os.path.join(GOLD, "full_seed0_head.wav.tmp"), "rb") as f:
        head = f.read(44)
    os.remove(os.path.join(GOLD, "full_seed0_head.wav.tmp"))
    with open(os.path.join(GOLD, "full_seed0_wav_header.bin"), "wb") as f:
        f.write(head)

    # ---- pure host functions
    H = W + "/out_host"
    d = {"timestep_values": i32(H + "/timestep_values.i32"),
         "timestep_embeddings": f32(H + "/timestep_embeddings.f32").reshape(-1, 1024),
         "apply_padding_a": i32(H + "/apply_padding_a.i32"), "apply_padding_b": i32(H + "/apply_padding_b.i32"),
         "denorm_mel": f32(H + "/denorm_mel.f32"), "normal_seed0_1000": f32(H + "/normal_seed0_1000.f32"),
         "uniform_seed7_1000": f32(H + "/uniform_seed7_1000.f32")}
    for n in (26, 113, 300):
        d[f"buckets_{n}"] = i32(H + f"/buckets_{n}.i32").reshape(n, n)
    with open(H + "/tiny.wav", "rb") as f:
        d["tiny_wav"] = np.frombuffer(f.read(), dtype=np.uint8)
    np.savez(os.path.join(GOLD, "hostfn.npz"), **d)
    tot = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(GOLD) for f in fs)
    print("golden bytes:", tot)


if __name__ == "__main__":
    main()
