#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TTS_DSTEP_TRACE=1 timeout -s KILL 240 python -m pytest tests/test_diffusion_gpu.py -m gpu -q -p no:cacheprovider -s -k "stage_driver" > gpurun_out/f_trace.log 2>&1
grep "dstep" gpurun_out/f_trace.log | tail -130 > gpurun_out/f_dstep_trace.txt
head -30 gpurun_out/f_dstep_trace.txt; tail -2 gpurun_out/f_trace.log
