#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
L=300 US=1 STEPS=1 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 8 -c 1 -o gpurun_out/r02_attn_tc_S1306 python tools/diff_batch_times.py > gpurun_out/q_attn.log 2>&1
echo "rc=$?"
ncu -i gpurun_out/r02_attn_tc_S1306.ncu-rep --page raw --csv > gpurun_out/r02_attn_tc_S1306_ncu_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_attn_tc_S1306.ncu-rep --page details 2>/dev/null | grep -E "Duration|Theoretical Occupancy|Achieved Occupancy|Registers Per|Waves Per SM|Issue Slots Busy|Executed Ipc|No Eligible|Eligible Warps|L2 Hit|Block Limit|Shared Memory Config|Dynamic Shared" | head -30
ncu -i gpurun_out/r02_attn_tc_S1306.ncu-rep --page details 2>/dev/null | grep -A12 "Warp State Statistics" | head -30
