#!/bin/bash
# round-2 GPU pass A: full GPU test suite (risky new kernels in their own, time-boxed invocation), decode-step times
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
timeout -s KILL 1500 python -m pytest tests -m gpu -q -k "not batched and not topk and not more_than_four" -p no:cacheprovider -s > gpurun_out/a_pytest_main.log 2>&1
echo "main rc=$?" >> gpurun_out/a_pytest_main.log
timeout -s KILL 400 python -m pytest tests/test_f16_gpu.py -m gpu -q -k "batched or topk or more_than_four" -p no:cacheprovider -s > gpurun_out/a_pytest_mega4.log 2>&1
echo "mega4 rc=$?" >> gpurun_out/a_pytest_mega4.log
timeout -s KILL 300 python tools/step_times.py > gpurun_out/a_step_times.txt 2>&1
tail -5 gpurun_out/a_pytest_main.log; tail -5 gpurun_out/a_pytest_mega4.log; cat gpurun_out/a_step_times.txt
