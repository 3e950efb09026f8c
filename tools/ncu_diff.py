# small diffusion workload for ncu (dev tool): L=44, S=191 (the bench shape), 2 sampling steps
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg
pkg = _pkg.import_pkg(); sw = _pkg.import_sub("synth_weights")
md = os.environ.get("TTS_MODEL_DIR", "/tmp/tortoise_b200_models"); sw.generate(md)
eng = pkg.Engine(dtype=pkg.DTYPE_F16, max_batch=1, max_positions=64)
eng.load_diffusion(md + "/ggml-diffusion-model.bin")
L, S, n = 44, 191, int(os.environ.get("NSTEPS", "2"))
rs = np.random.RandomState(0)
lat = rs.randn(L, 1024).astype(np.float32)
noise = rs.randn((n + 1) * 100 * S).astype(np.float32)
mel = eng.diffusion_sample(lat, S, n, noise)
print("ms", eng.last_stage_ms)
mel = eng.diffusion_sample(lat, S, n, noise)
print("ms", eng.last_stage_ms)
eng.close()
