#!/bin/bash
# round-2 GPU pass B: diffusion sampling step as one persistent kernel -- parity, then timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -s KILL 240 python -m pytest tests/test_diffusion_gpu.py -m gpu -q -p no:cacheprovider -s -x > gpurun_out/b_pytest_diff.log 2>&1
echo "diff rc=$?" >> gpurun_out/b_pytest_diff.log
tail -15 gpurun_out/b_pytest_diff.log
timeout -s KILL 240 python -m pytest tests/test_pipeline_gpu.py tests/test_f16_gpu.py -m gpu -q -p no:cacheprovider -s -k "seed_matched or 200_steps" > gpurun_out/b_pytest_pipe.log 2>&1
echo "pipe rc=$?" >> gpurun_out/b_pytest_pipe.log
tail -8 gpurun_out/b_pytest_pipe.log
timeout -s KILL 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/b_bench.json')); print(d['value'], d['e2e'], d['stage_ms'])"
TTS_NO_DSTEP=1 timeout -s KILL 400 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/b_bench_old.json 2> gpurun_out/b_bench_old.err
python -c "
import json; d=json.load(open('gpurun_out/b_bench_old.json')); print('old path', d['value'], d['stage_ms'])"
