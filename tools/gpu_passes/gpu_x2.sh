#!/bin/bash
# 2 GPUs: bench.py --gpus 2 with the single looping nvidia-smi sampler on rank 0
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/x2_bench_n2.json 2> gpurun_out/x2_bench_n2.err
echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/x2_bench_n2.json') if l.startswith('{')][-1])
print('N2 value', d['value'], 'e2e', d['e2e']['value'], d['stage_ms'], 'tok/s', d.get('ar_mel_tokens_per_s'), d['clocks'])
PY
