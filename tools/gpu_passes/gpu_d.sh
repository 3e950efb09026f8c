#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for bn in 32 64; do
TTS_DSTEP_BN=$bn TTS_DSTEP_TRACE=1 timeout -s KILL 240 python -m pytest tests/test_diffusion_gpu.py -m gpu -q -p no:cacheprovider -s -k "stage_driver" > gpurun_out/d_trace_$bn.log 2>&1
grep "dstep" gpurun_out/d_trace_$bn.log | tail -130 > gpurun_out/d_dstep_trace_bn$bn.txt
tail -1 gpurun_out/d_dstep_trace_bn$bn.txt; tail -2 gpurun_out/d_trace_$bn.log
done
timeout -s KILL 240 python -m pytest tests/test_diffusion_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/d_pytest_diff.log 2>&1; tail -2 gpurun_out/d_pytest_diff.log
for bn in 32 64; do
TTS_DSTEP_BN=$bn timeout -s KILL 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/d_bench_$bn.json 2> gpurun_out/d_bench_$bn.err
python -c "
import json; d=json.load(open('gpurun_out/d_bench_$bn.json')); print('bn $bn', d['value'], d['stage_ms'])"
done
BS=8,16 timeout -s KILL 200 python tools/step_times.py > gpurun_out/d_step_times.txt 2>&1
cat gpurun_out/d_step_times.txt
