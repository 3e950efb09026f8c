#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
( time timeout -s KILL 900 python bench.py > gpurun_out/s_bench_n1.json 2> gpurun_out/s_bench_n1.err ) 2>&1 | grep real
tail -2 gpurun_out/s_bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/s_bench_n1.json') if l.startswith('{')][-1])
print('value', d['value'], 'e2e', d['e2e']['value'], d['stage_ms'], 'launches', d['gpu_launches'], 'ms/step', d['ms_per_step'])
print('roofline', d['roofline']['frac'], d['roofline']['us_per_launch'], {k:(round(v['us_per_step']), round(v['tok_s'])) for k,v in d['roofline']['decode_step_batched'].items()})
print('tensor', d['roofline_tensor']['achieved'], d['roofline_tensor']['frac'], d['roofline_tensor']['us_per_launch'])
print('cpu', d.get('cpu_baseline'))
print('c3', {k:v for k,v in d.get('c3',{}).items() if k!='config'})
print('clocks', d['clocks'])
PY
( time timeout -s KILL 900 python bench.py --config C5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s_bench_c5.json 2> gpurun_out/s_bench_c5.err ) 2>&1 | grep real
tail -2 gpurun_out/s_bench_c5.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/s_bench_c5.json') if l.startswith('{')][-1])
print('C5 value', d['value'], 'e2e', d['e2e']['value'], d['stage_ms'], d.get('stage_rtf'), 'tok/s', d.get('ar_mel_tokens_per_s'))
PY
