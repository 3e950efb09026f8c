#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=300 US=1 STEPS=10 timeout -s KILL 300 python tools/diff_batch_times.py 2>&1 | tail -1
US=1,8 timeout -s KILL 300 python tools/diff_batch_times.py 2>&1 | tail -2
timeout -s KILL 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r_pytest_gpu.log 2>&1; tail -5 gpurun_out/r_pytest_gpu.log
