# in-kernel %globaltimer timeline of the tcgen05 GEMM launches of one denoiser pass (dev tool)
import os, sys, numpy as np
os.environ["TTS_TC5_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg
pkg = _pkg.import_pkg(); sw = _pkg.import_sub("synth_weights")
md = os.environ.get("TTS_MODEL_DIR", "/tmp/tortoise_b200_models"); sw.generate(md)
eng = pkg.Engine(dtype=pkg.DTYPE_F16, max_batch=1, max_positions=64)
eng.load_diffusion(md + "/ggml-diffusion-model.bin")
L, S = 44, 191
rs = np.random.RandomState(0)
lat = rs.randn(L, 1024).astype(np.float32)
x = rs.randn(100, S).astype(np.float32)
for i in range(2):
    print("=== pass", i, file=sys.stderr, flush=True)
    eng.diffusion_eps(lat, x, 500, 0)
eng.close()
