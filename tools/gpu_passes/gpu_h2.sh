#!/bin/bash
# round-2 GPU pass H2 (2 GPUs): C-ABI NCCL group, CLI --gpus 2, bench N = 2 (configs[3])
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L
timeout -s KILL 600 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -p no:cacheprovider -s -k "gather_select or two_gpus or cli" > gpurun_out/h_pytest.log 2>&1; tail -8 gpurun_out/h_pytest.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench_n2.json 2> gpurun_out/h_bench_n2.err
echo "bench n2 rc=$?"; tail -5 gpurun_out/h_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/h_bench_n2.json') if l.startswith('{')][-1])
print('N2 value', d['value'], 'e2e', d['e2e']['value'], d['stage_ms'], 'tok/s', d.get('ar_mel_tokens_per_s'), d['config']['workload'][:60])
PY
