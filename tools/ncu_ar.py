# short AR workload for ncu launch lists (dev tool)
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg
pkg = _pkg.import_pkg(); sw = _pkg.import_sub("synth_weights")
md = os.environ.get("TTS_MODEL_DIR", "/tmp/tortoise_b200_models"); sw.generate(md)
voice = np.fromfile("tests/golden/models/mol.bin", np.float32)
g = np.load("tests/golden/ar_b1.npz")
dt = pkg.DTYPE_F16 if os.environ.get("DT", "f16") == "f16" else pkg.DTYPE_F32
eng = pkg.Engine(dtype=dt, max_batch=max(4, int(os.environ.get("B", "1"))), max_positions=404)
eng.load_ar(md + "/ggml-model.bin")
B = int(os.environ.get("B", "1"))
eng.ar_prefill(g["tokens"], voice, B)
for i in range(4): eng.ar_step([100+i]*B, i+2)
eng.close()
