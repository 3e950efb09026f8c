#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -s KILL 500 python -m pytest tests/test_diffusion_gpu.py tests/test_f16_gpu.py -m gpu -q -p no:cacheprovider -s > gpurun_out/m_pytest_diff.log 2>&1; grep -E "max-abs|nmse|passed|failed|Error" gpurun_out/m_pytest_diff.log | tail -30
US=1,8 timeout -s KILL 300 python tools/diff_batch_times.py > gpurun_out/m_diff_times_S191.txt 2>&1; cat gpurun_out/m_diff_times_S191.txt
L=300 US=1,4 STEPS=10 timeout -s KILL 300 python tools/diff_batch_times.py > gpurun_out/m_diff_times_S1306.txt 2>&1; cat gpurun_out/m_diff_times_S1306.txt
