#!/bin/bash
# round-2 GPU pass T (8 GPUs): configs[3] = 64 candidates on 8 GPUs under torchrun, then the C++ CLI with --gpus 8
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/t_bench_n8.json 2> gpurun_out/t_bench_n8.err
echo "bench n8 rc=$?"; tail -3 gpurun_out/t_bench_n8.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/t_bench_n8.json') if l.startswith('{')][-1])
print('N8 value', d['value'], 'e2e', d['e2e']['value'], d['stage_ms'], 'tok/s', d.get('ar_mel_tokens_per_s'), d['config']['global_candidates'], d['clocks'])
PY
python - <<'PY' > gpurun_out/t_cli_gpus8.txt 2>&1
import os, subprocess, sys, json, tempfile
sys.path.insert(0, os.getcwd())
import _pkg
pkg = _pkg.import_pkg(); sw = _pkg.import_sub("synth_weights")
md = os.environ.get("TTS_MODEL_DIR", "/tmp/tortoise_b200_models"); sw.generate(md)
exe = os.path.join(os.path.dirname(pkg.LIB_PATH), "tortoise")
tmp = tempfile.mkdtemp(); work = os.path.join(tmp, "build"); os.mkdir(work); os.symlink(md, os.path.join(tmp, "models"))
r = subprocess.run([exe, "--seed", "0", "--gpus", "8", "--candidates", "64", "--dtype", "f16", "--bench-json", "x"], cwd=work, capture_output=True, text=True, timeout=500)
print("rc", r.returncode); print(r.stdout[-1500:]); print(r.stderr[-800:])
PY
tail -12 gpurun_out/t_cli_gpus8.txt
