#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_diffusion_gpu.py -m gpu -q -p no:cacheprovider -s > gpurun_out/i_pytest_diff.log 2>&1; tail -8 gpurun_out/i_pytest_diff.log
timeout -s KILL 300 python tools/diff_batch_times.py > gpurun_out/i_diff_batch_times.txt 2>&1; cat gpurun_out/i_diff_batch_times.txt
