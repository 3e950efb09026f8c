# quick AR timing probe (dev tool; bench.py is the contract)
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg
pkg = _pkg.import_pkg(); sw = _pkg.import_sub("synth_weights")
md = os.environ.get("TTS_MODEL_DIR", "/tmp/tortoise_b200_models"); sw.generate(md)
voice = np.fromfile("tests/golden/models/mol.bin", np.float32)
g = np.load("tests/golden/ar_b1.npz")
for dtype, name in ((pkg.DTYPE_F32, "f32"), (pkg.DTYPE_F16, "f16")):
    for B in (1, 2, 4):
        eng = pkg.Engine(dtype=dtype, max_batch=4, max_positions=404)
        eng.load_ar(md + "/ggml-model.bin")
        for op in range(4):
            ms, by = eng.bench_gemv(op, B, 120)
            print(f"{name} B={B} gemv op{op}: {ms*1e3:.2f} us/launch, {by/ms/1e6:.0f} GB/s", flush=True)
        eng.ar_prefill(g["tokens"], voice, B)
        print(f"{name} B={B} prefill {eng.last_stage_ms:.3f} ms")
        for i in range(5): eng.ar_step([100]*B, i+2)
        t=[]
        for i in range(40):
            eng.ar_step([100+i]*B, i+7); t.append(eng.last_stage_ms)
        print(f"{name} B={B} decode step median {np.median(t)*1e3:.1f} us  -> {B/np.median(t)*1e3:.0f} tok/s", flush=True)
        eng.close()
