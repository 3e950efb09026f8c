# dev tool: per-phase clock trace of the tcgen05 GEMM (TTS_TC5_TRACE=1) on the denoiser's 3-tap convolution
import os, sys
os.environ["TTS_TC5_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg
pkg = _pkg.import_pkg(); sw = _pkg.import_sub("synth_weights")
md = os.environ.get("TTS_MODEL_DIR", "/tmp/tortoise_b200_models"); sw.generate(md)
eng = pkg.Engine(dtype=pkg.DTYPE_F16, max_batch=1, max_positions=64)
eng.load_diffusion(md + "/ggml-diffusion-model.bin")
for S in [int(x) for x in os.environ.get("SS", "191,1306").split(",")]:
    ms, fl = eng.bench_conv3(S, 3)
    print("S", S, "ms", ms, "TFLOP/s", fl / ms / 1e9, flush=True)
eng.close()
