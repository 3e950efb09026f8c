#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_f16_gpu.py -m gpu -q -p no:cacheprovider -s -k "utterance_batched or batched_decode_vs_oracle" > gpurun_out/u_pytest.log 2>&1; grep -E "utterance|passed|failed|Error|error" gpurun_out/u_pytest.log | tail -14
( time timeout -s KILL 600 python bench.py --config C5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/u_bench_c5.json 2> gpurun_out/u_bench_c5.err ) 2>&1 | grep real
tail -2 gpurun_out/u_bench_c5.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/u_bench_c5.json') if l.startswith('{')][-1])
print('C5 value', d['value'], 'e2e', d['e2e']['value'], d['stage_ms'], d.get('stage_rtf'), 'tok/s', d.get('ar_mel_tokens_per_s'))
PY
