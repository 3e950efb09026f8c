# tiny f16 decode workload for compute-sanitizer racecheck (dev tool):
#   compute-sanitizer --tool racecheck --kernel-regex kns=mega3 python tools/racecheck_ar.py
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg
pkg = _pkg.import_pkg(); sw = _pkg.import_sub("synth_weights")
md = os.environ.get("TTS_MODEL_DIR", "/tmp/tortoise_b200_models"); sw.generate(md)
voice = np.fromfile("tests/golden/models/mol.bin", np.float32)
g = np.load("tests/golden/ar_b1.npz")
B = int(os.environ.get("B", "2"))
eng = pkg.Engine(dtype=pkg.DTYPE_F16, max_batch=4, max_positions=404)
eng.load_ar(md + "/ggml-model.bin")
eng.ar_prefill(g["tokens"], voice, B)
for i in range(2): eng.ar_step([100 + i] * B, i + 2)
print("done", flush=True)
eng.close()
