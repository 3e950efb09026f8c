#!/bin/bash
# diffusion step: per-kernel launch list (time + DRAM bytes) of 2 sampling steps at S = 191, and --set full of tc5v2 + the attention kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
NSTEPS=2 timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_diffusion_step_launches.csv python tools/ncu_diff.py > gpurun_out/n_diff_list.log 2>&1
echo "list rc=$?"; wc -l gpurun_out/r02_diffusion_step_launches.csv
NSTEPS=2 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:tc5v2 -s 60 -c 3 -o gpurun_out/r02_tc5v2 python tools/ncu_diff.py > gpurun_out/n_tc5v2.log 2>&1
echo "tc5v2 rc=$?"
ncu -i gpurun_out/r02_tc5v2.ncu-rep --page raw --csv > gpurun_out/r02_tc5v2_ncu_full_raw.csv 2>/dev/null
NSTEPS=2 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 10 -c 1 -o gpurun_out/r02_attn_tc python tools/ncu_diff.py > gpurun_out/n_attn.log 2>&1
ncu -i gpurun_out/r02_attn_tc.ncu-rep --page raw --csv > gpurun_out/r02_attn_tc_ncu_full_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
