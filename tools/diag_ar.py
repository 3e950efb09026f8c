import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg
pkg = _pkg.import_pkg(); sw = _pkg.import_sub("synth_weights"); hl=_pkg.import_sub("host").HostLib()
md = os.environ.get("TTS_MODEL_DIR", "/tmp/tortoise_b200_models"); sw.generate(md)
voice = np.fromfile("tests/golden/models/mol.bin", np.float32)
g = np.load("tests/golden/ar_b1.npz")
eng = pkg.Engine(dtype=pkg.DTYPE_F32, max_batch=4, max_positions=404)
eng.load_ar(md + "/ggml-model.bin")
codes=[int(c) for c in g["codes500"][:18]]
lg = eng.ar_prefill(g["tokens"], voice, 1)
r1=hl.rng(0); r2=hl.rng(0)
prev=np.array([[1]*17+[8192]])
for i in range(18):
    ref=g[f"logits_{i}"][None]
    a=hl.sample(r1, lg, prev)[0]; b=hl.sample(r2, ref, prev)[0]
    d=np.abs(lg[0]-ref[0])
    print(i, 'maxerr %.2e meanerr %.2e'%(d.max(), d.mean()), 'mine',a,'ref',b, 'OK' if a==b else 'FLIP')
    prev=np.array([[codes[i]]])
    lg=eng.ar_step([codes[i]], i+2)
