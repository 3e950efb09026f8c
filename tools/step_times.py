# decode-step device time per candidate count (dev tool; bench.py is the contract):
# `iters` back-to-back steps between two CUDA events (tts_bench_decode_step), f16 weights
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg
pkg = _pkg.import_pkg(); sw = _pkg.import_sub("synth_weights")
md = os.environ.get("TTS_MODEL_DIR", "/tmp/tortoise_b200_models"); sw.generate(md)
voice = np.fromfile("tests/golden/models/mol.bin", np.float32)
g = np.load("tests/golden/ar_b1.npz")
for B in [int(x) for x in os.environ.get("BS", "1,2,4,8,16").split(",")]:
    eng = pkg.Engine(dtype=pkg.DTYPE_F16, max_batch=B, max_positions=404)
    eng.load_ar(md + "/ggml-model.bin")
    eng.ar_prefill(g["tokens"], voice, B)
    ms, by = eng.bench_decode_step(int(os.environ.get("NSTEP", "100")))
    print(f"f16 B={B:2d}: {ms*1e3:7.1f} us/step  {B/ms*1e3:8.0f} tok/s  {by/ms/1e6:7.0f} GB/s algorithmic", flush=True)
    eng.close()
