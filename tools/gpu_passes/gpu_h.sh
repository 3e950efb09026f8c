#!/bin/bash
# round-2 GPU pass H (2 GPUs): bench N = 1 (configs[1] + configs[2]), C-ABI NCCL group, CLI --gpus 2, bench N = 2 (configs[3])
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -s KILL 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/h_bench_n1.json 2> gpurun_out/h_bench_n1.err
echo "bench n1 rc=$?"; tail -3 gpurun_out/h_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/h_bench_n1.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], d['stage_ms'], 'launches', d['gpu_launches'])
print('roofline', d['roofline']['frac'], d['roofline']['us_per_launch'], {k:(v['us_per_step'], v['tok_s']) for k,v in d['roofline']['decode_step_batched'].items()})
print('tensor', d['roofline_tensor']['achieved'], d['roofline_tensor']['frac'], d['roofline_tensor']['us_per_launch'])
print('c3', json.dumps(d.get('c3'))[:900])
PY
timeout -s KILL 600 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -p no:cacheprovider -s -k "gather_select or two_gpus or cli" > gpurun_out/h_pytest.log 2>&1; tail -6 gpurun_out/h_pytest.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/h_bench_n2.json 2> gpurun_out/h_bench_n2.err
echo "bench n2 rc=$?"; tail -5 gpurun_out/h_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/h_bench_n2.json') if l.startswith('{')][-1])
print('N2 value', d['value'], 'e2e', d['e2e']['value'], d['stage_ms'], 'tok/s', d['ar_mel_tokens_per_s'], d['config']['workload'][:60])
PY
