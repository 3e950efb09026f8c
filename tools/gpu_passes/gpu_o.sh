#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=300 US=1 STEPS=2 timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,lts__t_bytes.sum,dram__bytes_read.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_diffusion_step_launches_S1306.csv python tools/diff_batch_times.py > gpurun_out/o.log 2>&1
echo "rc=$?"; wc -l gpurun_out/r02_diffusion_step_launches_S1306.csv
