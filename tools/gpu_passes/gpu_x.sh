#!/bin/bash
# 2 GPUs: the CLI with --gpus 2 (NCCL communicators built in the background; stage times in the JSON) and bench.py --gpus 2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -p no:cacheprovider -k "two_gpus or gather_select" 2>&1 | tail -2
python - <<'PY' > gpurun_out/x_cli_gpus2.txt 2>&1
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
tail -6 gpurun_out/x_cli_gpus2.txt | cut -c1-400
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/x_bench_n2.json 2> gpurun_out/x_bench_n2.err
echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/x_bench_n2.json') if l.startswith('{')][-1])
print('N2 value', d['value'], 'e2e', d['e2e']['value'], d['stage_ms'], 'tok/s', d.get('ar_mel_tokens_per_s'))
PY
