"""CPU tests of the host-side C++ (libtortoise_host.so) against golden vectors produced by
the UNMODIFIED reference (tests/golden/make_golden.py).  All integer results are bit-exact."""
import json
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
TOKJSON = os.path.join(GOLDEN, "models", "tokenizer.json")


@pytest.fixture(scope="module")
def hl():
    import _pkg
    return _pkg.import_sub("host").HostLib()


def test_tokenizer_matches_reference_corpus(hl):
    with open(os.path.join(GOLDEN, "tokenizer_corpus.json")) as f:
        corpus = json.load(f)
    assert hl.vocab_size(TOKJSON) == 255  # SURVEY A-14: HF vocab - [STOP] + "version"
    for text, ids in corpus.items():
        assert hl.tokenize(TOKJSON, text) == ids, text


def test_rng_streams_match_libstdcxx_reference(hl):
    g = np.load(os.path.join(GOLDEN, "hostfn.npz"))
    r = hl.rng(0)
    assert np.array_equal(r.normal(1000), g["normal_seed0_1000"])
    r = hl.rng(7)
    u = np.array([r.uniform() for _ in range(1000)], dtype=np.float32)
    assert np.array_equal(u, g["uniform_seed7_1000"])
    # reseeding restarts the generator (main.cpp:6546)
    r.seed(7)
    assert r.uniform() == g["uniform_seed7_1000"][0]


@pytest.mark.parametrize("tag", ["a", "b"])
def test_sampler_bit_exact_vs_reference(hl, tag):
    g = np.load(os.path.join(GOLDEN, "sampler.npz"))
    logits, prev, want = g[f"{tag}_logits"], g[f"{tag}_prev"], g[f"{tag}_samples"]
    for literal in (False, True):
        r = hl.rng(int(g[f"{tag}_seed"]))
        got = np.stack([hl.sample(r, logits, prev.reshape(4, -1), literal=literal) for _ in range(want.shape[0])])
        assert np.array_equal(got, want), ("literal" if literal else "fast")


def test_sampler_fast_equals_literal_on_random_logits(hl):
    rs = np.random.RandomState(0)
    for trial in range(6):
        logits = (rs.randn(3, 8194) * (0.5 + trial)).astype(np.float32)
        if trial == 5:  # force ties among the survivors: fast path must defer to the literal one
            logits[:, 100] = logits[:, 200] = 30.0
        prev = rs.randint(0, 8194, size=(3, 2)).astype(np.int32)
        a = hl.sample(hl.rng(trial), logits, prev)
        b = hl.sample(hl.rng(trial), logits, prev, literal=True)
        assert np.array_equal(a, b)


def test_sampler_fast_path_threshold_edge_cases(hl):
    """the heap-based top-k threshold of the fast path against the literal algorithm where it could go
    wrong: values one ulp around the 50th largest, quotients that collide after the division by the
    temperature, many equal values at the threshold, penalised entries at the threshold, negatives."""
    rs = np.random.RandomState(7)
    cases = []
    for trial in range(40):
        lg = (rs.randn(2, 8194) * rs.choice([0.01, 1.0, 20.0])).astype(np.float32)
        if trial % 4 == 1:  # a ladder of adjacent floats straddling the 50th largest value
            top = np.sort(lg[0])[-50]
            ladder = top
            for k in range(12):
                lg[0, 1000 + k] = ladder
                ladder = np.nextafter(ladder, np.float32(-np.inf), dtype=np.float32)
        if trial % 4 == 2:  # a plateau of equal values exactly at the threshold (ties -> literal path)
            lg[1, 300:340] = np.sort(lg[1])[-50]
        if trial % 4 == 3:  # all negative, tiny spread
            lg = -np.abs(lg) - np.float32(5.0)
        prev = rs.randint(0, 8194, size=(2, 6)).astype(np.int32)
        if trial % 5 == 0:  # penalised entries among the largest
            prev[0, :3] = np.argsort(lg[0])[-3:]
        cases.append((lg, prev))
    for i, (lg, prev) in enumerate(cases):
        for seed in (0, 1):
            a = hl.sample(hl.rng(seed), lg, prev)
            b = hl.sample(hl.rng(seed), lg, prev, literal=True)
            assert np.array_equal(a, b), i


def test_padding_and_trim(hl):
    import tortoise_oracle as O
    g = np.load(os.path.join(GOLDEN, "hostfn.npz"))
    assert np.array_equal(hl.apply_padding([5, 6, 7, 8139, 83, 83, 8193]), g["apply_padding_a"])
    assert np.array_equal(hl.apply_padding([100, 200, 8139, 8139]), g["apply_padding_b"])
    assert list(g["apply_padding_a"]) == O.apply_padding([5, 6, 7, 8139, 83, 83, 8193])
    b1 = np.load(os.path.join(GOLDEN, "ar_b1.npz"))
    n = hl.trim_count(b1["codes500"])
    assert n * 1024 == b1["trimmed_latents"].size
    assert n == O.trim_count(list(b1["codes500"]))
    assert hl.trim_count(np.full(500, 83)) == 8
    assert hl.trim_count(np.arange(500) % 80) == 500
    with pytest.raises(RuntimeError):
        hl.apply_padding(np.zeros(501))


def test_timestep_map_embedding_buckets(hl):
    import tortoise_oracle as O
    literal = [0, 51, 101, 152, 202, 253, 304, 354, 405, 456, 506, 557, 607, 658, 709, 759, 810, 861, 911, 962, 1012,
               1063, 1114, 1164, 1215, 1266, 1316, 1367, 1417, 1468, 1519, 1569, 1620, 1670, 1721, 1772, 1822, 1873,
               1924, 1974, 2025, 2075, 2126, 2177, 2227, 2278, 2329, 2379, 2430, 2480, 2531, 2582, 2632, 2683, 2733,
               2784, 2835, 2885, 2936, 2987, 3037, 3088, 3138, 3189, 3240, 3290, 3341, 3392, 3442, 3493, 3543, 3594,
               3645, 3695, 3746, 3797, 3847, 3898, 3948, 3999]  # main.cpp:5641-5648
    assert list(hl.timestep_map(80)) == literal
    g = np.load(os.path.join(GOLDEN, "hostfn.npz"))
    for t, want in zip(g["timestep_values"], g["timestep_embeddings"]):
        assert np.array_equal(hl.timestep_embedding(int(t)), want), t
        assert np.array_equal(O.timestep_embedding(int(t)), want)
    for n in (26, 113, 300):
        assert np.array_equal(hl.relative_position_buckets(n), g[f"buckets_{n}"])
        assert np.array_equal(O.relative_position_buckets(n), g[f"buckets_{n}"])


def test_ddpm_schedule_sane(hl):
    s = hl.ddpm_schedule(80)
    assert s.shape == (80, 9)
    assert s[0, 8] == 3999 and s[-1, 8] == 0 and s[-1, 7] == 1 and s[:-1, 7].sum() == 0
    assert abs(s[0, 0] - 2.0 * (1 - 79 / 80)) < 1e-7 and s[-1, 0] == 2.0
    assert np.all(np.diff(s[:, 1]) < 0)  # sqrt(1/alpha_bar) shrinks towards t = 0
    assert np.all(s[:, 5] <= s[:, 6] + 1e-6)  # posterior log-variance <= log beta


def test_wav_header_and_payload_bytes(hl, tmp_path):
    g = np.load(os.path.join(GOLDEN, "hostfn.npz"))
    p = tmp_path / "t.wav"
    hl.write_wav(str(p), np.array([0.0, 0.5, -0.5, 1.0], dtype=np.float32))
    assert p.read_bytes() == g["tiny_wav"].tobytes()


def test_c_abi_exports_every_declared_symbol():
    """both libraries load and export what include/*.h declare (no compute calls)."""
    import ctypes
    import _pkg
    pkg = _pkg.import_pkg()
    decl = {}
    for hdr in ("tortoise_b200.h", "tortoise_host.h", "tortoise_b200_bench.h"):
        src = open(os.path.join(ROOT, "include", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        decl[hdr] = set(re.findall(r"\b(tts_[a-z0-9_]+)\s*\(", src))
    full = ctypes.CDLL(pkg.LIB_PATH)
    for name in sorted(decl["tortoise_b200.h"] | decl["tortoise_host.h"] | decl["tortoise_b200_bench.h"]):
        assert hasattr(full, name), f"libtortoise_b200.so lacks {name}"
    host = ctypes.CDLL(os.path.join(os.path.dirname(pkg.LIB_PATH), "libtortoise_host.so"))
    drivers = {"tts_host_autoregressive", "tts_host_diffusion", "tts_host_vocoder", "tts_host_latents",
               "tts_host_diffusion_batch", "tts_host_autoregressive_multi"}
    for name in sorted(decl["tortoise_host.h"] - drivers):
        assert hasattr(host, name), f"libtortoise_host.so lacks {name}"


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    import _pkg
    pkg = _pkg.import_pkg()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.TTSError) as e:
        pkg.Engine()
    assert e.value.code == -3  # TTS_ENODEV


def test_text_front_end_normalize_and_split(hl):
    """optional front-end: everything it emits is in the reference's alphabet."""
    n = hl.normalize_text
    assert n("this is a test message.") == "this is a test message."          # supported text is untouched
    assert n("Hello, World!") == "hello, world!"
    assert n("I have 2 cats & 21 dogs") == "i have two cats and twenty-one dogs"
    assert n("It costs $1500.") == "it costs one thousand five hundred dollars."
    assert n("pi is 3.14") == "pi is three point one four"
    assert n("100% sure") == "one hundred percent sure"
    assert n("1000000 and 1000001") == "one million and one million one"
    assert n("x=y+z @ home") == "x equals y plus z at home"
    assert n("café — naïve") == "caf na ve"                      # non-ASCII bytes -> space
    assert n("  a   b  ,c ") == "a b,c"
    assert n("12345678901234567") == " ".join(
        ["one", "two", "three", "four", "five", "six", "seven", "eight", "nine", "zero", "one", "two", "three", "four",
         "five", "six", "seven"])
    import re
    for s in ("Dr. Freeman? 42!", "A+B=C", "tel: 555-0199", "\t\ttabs\nnewlines"):
        assert re.fullmatch(r"[a-z .,!?'-]*", n(s)), n(s)
    text = ("this is the first sentence. here is a second one, with a comma! and a third? "
            "finally a very long tail without any punctuation that has to be cut at a space somewhere")
    for limit in (20, 40, 80, 1000):
        parts = hl.split_text(text, limit)
        assert all(0 < len(p) <= limit for p in parts)
        assert " ".join(parts).split() == text.split()                          # nothing lost, order kept
    assert hl.split_text(text, 40)[0] == "this is the first sentence."
    assert hl.split_text("short", 100) == ["short"]
    assert hl.split_text("", 100) == []
    assert hl.split_text("a" * 50, 16) == ["a" * 16, "a" * 16, "a" * 16, "aa"]


def test_sparse_sampler_matches_full_row_sampler_bit_exactly(hl):
    """tts_host_sample_sparse (fed with the 64 largest raw logits, as tts_ar_step_topk delivers them) draws
    the same tokens from the same generator state as the full-row sampler -- itself pinned against the
    reference's process_logits_and_sample above -- or declines without touching the generator."""
    g = np.load(os.path.join(GOLDEN, "sampler.npz"))
    rs = np.random.RandomState(3)
    n_sparse = n_decl = 0
    for tag in ("a", "b"):
        logits = g[f"{tag}_logits"]
        for trial in range(40):
            b = trial % 4
            row = logits[b].copy()
            if trial >= 8:  # perturbed copies: different winners, penalised tokens inside / outside the set
                row = (row + rs.normal(0, 0.3, size=row.shape)).astype(np.float32)
            order = np.argsort(-row, kind="stable")[:64]
            prev = np.array([int(order[rs.randint(0, 64)]), int(rs.randint(0, 8194))] if trial % 3 else [1] * 17 + [8192],
                            np.int32)
            seed = 100 + trial
            r_full, r_sp = hl.rng(seed), hl.rng(seed)
            want, want_lp = hl.sample(r_full, row[None], prev[None], want_logprob=True)
            perm = rs.permutation(64)  # the device emits the pairs unsorted
            got = hl.sample_sparse(r_sp, row[order][perm], order[perm], prev)
            if got is None:
                n_decl += 1
                assert r_sp.uniform() == hl.rng(seed).uniform()  # generator untouched
                continue
            n_sparse += 1
            assert got[0] == int(want[0]), (tag, trial)
            assert got[1] == float(want_lp[0]) or (np.isnan(got[1]) and np.isnan(want_lp[0]))
            assert r_sp.uniform() == r_full.uniform()  # same number of draws consumed
    assert n_sparse >= 60, (n_sparse, n_decl)


def test_sparse_sampler_declines_when_the_cut_is_too_close(hl):
    """near-uniform row: the 64th largest value sits inside the survivor window of the 50th -> decline"""
    row = np.zeros(8194, np.float32)
    row[:70] = 1.0
    order = np.argsort(-row, kind="stable")[:64]
    assert hl.sample_sparse(hl.rng(0), row[order], order, np.array([5], np.int32)) is None
