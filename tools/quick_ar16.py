# quick AR decode timing, f16 only (dev tool; bench.py is the contract)
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg
pkg = _pkg.import_pkg(); sw = _pkg.import_sub("synth_weights")
md = os.environ.get("TTS_MODEL_DIR", "/tmp/tortoise_b200_models"); sw.generate(md)
voice = np.fromfile("tests/golden/models/mol.bin", np.float32)
g = np.load("tests/golden/ar_b1.npz")
for B in ((1,) if os.environ.get('B_ONLY') else (1, 2, 4, 8, 16)):
    eng = pkg.Engine(dtype=pkg.DTYPE_F16, max_batch=max(B, 4), max_positions=404)
    eng.load_ar(md + "/ggml-model.bin")
    eng.ar_prefill(g["tokens"], voice, B)
    for i in range(5): eng.ar_step([100]*B, i+2)
    t=[]
    for i in range(int(os.environ.get('NSTEP','60'))):
        eng.ar_step([100+i]*B, i+7); t.append(eng.last_stage_ms)
    if os.environ.get('PER_STEP'): print(' '.join(f"{x*1e3:.0f}" for x in t))
    print(f"f16 B={B} decode step median {np.median(t)*1e3:.1f} us min {np.min(t)*1e3:.1f} -> {B/np.median(t)*1e3:.0f} tok/s", flush=True)
    eng.close()
