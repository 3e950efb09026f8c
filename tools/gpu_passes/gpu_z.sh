#!/bin/bash
# last pass of the round: full GPU suite, smoke(), bench.py on the final code
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02s_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02s_pytest_gpu.log
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 900 python bench.py > gpurun_out/r02s_bench_n1.json 2> gpurun_out/z_bench.err; tail -1 gpurun_out/z_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02s_bench_n1.json') if l.startswith('{')][-1])
print('value', d['value'], 'e2e', d['e2e']['value'], d['stage_ms'], 'launches', d['gpu_launches'])
print('roofline', d['roofline']['frac'], d['roofline']['us_per_launch'], 'tensor', d['roofline_tensor']['achieved'], d['roofline_tensor']['frac'])
print('c3', d['c3']['rtf_e2e'], d['c3']['ar_mel_tokens_per_s'], 'cpu', d['cpu_baseline']['value'])
PY
