#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TTS_DSTEP_TRACE=1 timeout -s KILL 240 python -m pytest tests/test_diffusion_gpu.py -m gpu -q -p no:cacheprovider -s -k "stage_driver" > gpurun_out/g_trace.log 2>&1
grep "dstep" gpurun_out/g_trace.log | tail -130 > gpurun_out/g_dstep_trace.txt
tail -1 gpurun_out/g_dstep_trace.txt; tail -2 gpurun_out/g_trace.log
timeout -s KILL 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/g_pytest.log 2>&1; tail -3 gpurun_out/g_pytest.log
timeout -s KILL 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
python -c "
import json; d=json.load(open('gpurun_out/g_bench.json')); print(d['value'], d['stage_ms'])"
TTS_NO_DSTEP=1 timeout -s KILL 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/g_bench_old.json 2> gpurun_out/g_bench_old.err
python -c "
import json; d=json.load(open('gpurun_out/g_bench_old.json')); print('per-op graph path', d['value'], d['stage_ms'])"
