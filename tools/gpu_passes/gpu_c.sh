#!/bin/bash
# round-2 GPU pass C: per-op timeline of the diffusion step kernel + decode-step times after the attention change
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TTS_DSTEP_TRACE=1 timeout -s KILL 240 python -m pytest tests/test_diffusion_gpu.py -m gpu -q -p no:cacheprovider -s -k "stage_driver" > gpurun_out/c_trace.log 2>&1
grep "dstep" gpurun_out/c_trace.log | tail -130 > gpurun_out/c_dstep_trace.txt
tail -3 gpurun_out/c_dstep_trace.txt
timeout -s KILL 300 python -m pytest tests/test_f16_gpu.py -m gpu -q -k "batched or topk" -p no:cacheprovider -s > gpurun_out/c_pytest_mega4.log 2>&1
tail -4 gpurun_out/c_pytest_mega4.log
BS=8,16 timeout -s KILL 200 python tools/step_times.py > gpurun_out/c_step_times.txt 2>&1
cat gpurun_out/c_step_times.txt
