# diffusion sampling: device time per utterance vs utterances per batch (dev tool; bench.py --config C5 is the contract)
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg
pkg = _pkg.import_pkg(); sw = _pkg.import_sub("synth_weights"); hostmod = _pkg.import_sub("host")
md = os.environ.get("TTS_MODEL_DIR", "/tmp/tortoise_b200_models"); sw.generate(md)
hl = hostmod.HostLib(full=True)
eng = pkg.Engine(dtype=pkg.DTYPE_F16, max_batch=1, max_positions=64)
eng.load_diffusion(md + "/ggml-diffusion-model.bin")
rs = np.random.RandomState(0)
L = int(os.environ.get("L", "44"))  # 44 latents -> S = 191 (the bench shape)
steps = int(os.environ.get("STEPS", "20"))
for U in [int(x) for x in os.environ.get("US", "1,2,4,8,16").split(",")]:
    lats = [(0.5 * rs.randn(L, 1024)).astype(np.float32) for _ in range(U)]
    for rep in range(2):
        rngs = [hl.rng(100 + u) for u in range(U)]
        t0 = time.perf_counter()
        mels = hl.diffusion_batch(eng, rngs, lats, steps) if U > 1 else [hl.diffusion(eng, rngs[0], lats[0], steps)]
        wall = time.perf_counter() - t0
    ms = eng.last_stage_ms
    S = mels[0].shape[1]
    flop = 2 * steps * U * (2.0 * S * 124.7e6 + 13 * 4.0 * S * S * 1024)
    print(f"U={U:2d} S={S}: {ms:8.2f} ms device for {steps} steps = {ms/steps*1e3:7.1f} us/step, {ms/steps/U*1e3:7.1f} us/step/utterance, "
          f"{flop/ms/1e9:6.1f} TFLOP/s, wall {wall*1e3:.1f} ms", flush=True)
eng.close()
