#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_diffusion_gpu.py -m gpu -q -p no:cacheprovider -s > gpurun_out/p_pytest_diff.log 2>&1; grep -E "80-step|batched|utterance|passed|failed|Error" gpurun_out/p_pytest_diff.log | tail -12
L=300 US=1,4 STEPS=10 timeout -s KILL 300 python tools/diff_batch_times.py 2>&1 | tail -3
L=120 US=1,4 STEPS=10 timeout -s KILL 300 python tools/diff_batch_times.py 2>&1 | tail -3
US=1,8 timeout -s KILL 300 python tools/diff_batch_times.py 2>&1 | tail -3
SS=1306 timeout -s KILL 200 python tools/tc5_trace.py 2>&1 | grep -E "^S|csz" | tail -3
