"""bench.py contract pieces that run without a GPU: the reference arm (the UNMODIFIED reference
CPU build from oracle/_ref timed on a bounded sample) prints ONE JSON line with the keys the
driver reads, and the B200 arm refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")


@pytest.mark.skipif(not os.path.exists(HARNESS), reason="oracle/_ref/ref_harness not built (needs /root/reference)")
def test_reference_arm_prints_one_json_line():
    # (--ref-sample bounded: the default reference arm runs ONE FULL utterance, ~5 minutes of CPU -- what the
    # driver measures; the bounded sample exercises the same plumbing in ~40 s)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--ref-sample", "bounded"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True
    assert d["metric"].startswith("real-time factor") and d["unit"] == "audio-s/wall-s"
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 4
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "configs[1]" in d["config"]["workload"]
    assert "BOUNDED" in d["cpu_baseline"]["sample"]


def test_workload_definitions_match_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    assert len(bench.PROMPT_C3) == 50 and len(bench.PROMPT_C4) == 200
    for name, cand, steps in (("C2", 1, 80), ("C3", 16, 80), ("C4", 8, 200), ("C5", 1, 80)):
        w = bench.workload(name, 8)
        assert w["cand"] == cand and w["steps"] == steps and not w["reduced"]
    assert bench.workload("C4", 8)["codes"] == 300 and bench.workload("C2", 1)["codes"] == 35 and bench.workload("C3", 1)["codes"] == 75
    ps = bench.c5_prompts()
    assert len(ps) == 256 and min(map(len, ps)) >= 20 and max(map(len, ps)) <= 300
    assert all(set(p) <= set("abcdefghijklmnopqrstuvwxyz .,!?'-") for p in ps + [bench.PROMPT_C3, bench.PROMPT_C4])
    cfg = bench.workload_config(bench.workload("C4", 8), 8)
    assert "configs[3]" in cfg["workload"] and cfg["global_candidates"] == 64


def test_b200_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)


def test_c5_sharding_partitions_the_utterances():
    """configs[4]: every utterance lands on exactly one rank and in exactly one batch, batches hold at most U
    utterances of neighbouring lengths, and the prompt set is the seeded one (20..300 characters)"""
    sys.path.insert(0, ROOT)
    import bench
    prompts = bench.c5_prompts()
    assert len(prompts) == 256 and all(20 <= len(p) <= 300 for p in prompts)
    assert prompts == bench.c5_prompts()  # seeded
    for world in (1, 2, 3, 8):
        seen = []
        for rank in range(world):
            batches = bench.c5_batches(prompts, rank, world, 8)
            assert all(1 <= len(b) <= 8 for b in batches)
            assert all(u % world == rank for b in batches for u in b)
            lens = [len(prompts[u]) for b in batches for u in b]
            assert lens == sorted(lens)
            seen += [u for b in batches for u in b]
        assert sorted(seen) == list(range(256))


def test_clock_sampler_summary_without_nvidia_smi():
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler([0, 1])
    assert s.summary()["reasons"] == ["nvidia-smi unavailable"]
    s.samples = [(1965.0, 1965.0), (1900.0, 1965.0), (1965.0, 1965.0)]
    s.reasons = {"sw_power_cap"}
    out = s.summary()
    assert out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"]
    assert out["gpus_sampled"] == [0, 1] and out["samples"] == 3
