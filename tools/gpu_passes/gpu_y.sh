#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -p no:cacheprovider -k "two_gpus or gather_select" 2>&1 | tail -2
python - <<'PY' > gpurun_out/y_cli_gpus2.txt 2>&1
import os, subprocess, sys, tempfile
sys.path.insert(0, os.getcwd())
import _pkg
pkg = _pkg.import_pkg(); sw = _pkg.import_sub("synth_weights")
md = os.environ.get("TTS_MODEL_DIR", "/tmp/tortoise_b200_models"); sw.generate(md)
exe = os.path.join(os.path.dirname(pkg.LIB_PATH), "tortoise")
tmp = tempfile.mkdtemp(); work = os.path.join(tmp, "build"); os.mkdir(work); os.symlink(md, os.path.join(tmp, "models"))
r = subprocess.run([exe, "--seed", "0", "--gpus", "2", "--candidates", "16", "--dtype", "f16", "--bench-json", "x"], cwd=work, capture_output=True, text=True, timeout=300)
print("rc", r.returncode); print(r.stdout[-1200:]); print(r.stderr[-500:])
PY
tail -6 gpurun_out/y_cli_gpus2.txt | cut -c1-400
