# one vocoder forward at the bench shape (S = 191 mel frames -> 201 frames with padding) for ncu (dev tool)
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg
pkg = _pkg.import_pkg(); sw = _pkg.import_sub("synth_weights")
md = os.environ.get("TTS_MODEL_DIR", "/tmp/tortoise_b200_models"); sw.generate(md)
eng = pkg.Engine(dtype=pkg.DTYPE_F16, max_batch=1, max_positions=64)
eng.load_vocoder(md + "/ggml-vocoder-model.bin")
S = int(os.environ.get("S", "191"))
rs = np.random.RandomState(0)
mel = rs.uniform(-1, 1, size=(100, S)).astype(np.float32)
noise = rs.randn((S + 10) * 64).astype(np.float32)
for _ in range(int(os.environ.get("REPS", "2"))):
    audio = eng.vocoder(mel, noise)
    print("ms", eng.last_stage_ms, "launches", eng.launch_count)
eng.close()
