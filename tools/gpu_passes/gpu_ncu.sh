#!/bin/bash
# round-2 ncu evidence (1 GPU): vocoder per-kernel time + DRAM bytes, batched decode kernel (--set full), launch list of a reduced bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python tools/ncu_voc.py > gpurun_out/n_voc_plain.txt 2>&1
NL=$(grep launches gpurun_out/n_voc_plain.txt | head -1 | awk '{print $4}')
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -s $NL -c 400 --csv --log-file gpurun_out/r02_vocoder_launches.csv python tools/ncu_voc.py > gpurun_out/n_voc_ncu.log 2>&1
echo "voc rc=$? first-pass launches=$NL"; wc -l gpurun_out/r02_vocoder_launches.csv
B=16 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:mega4 -s 2 -c 2 -o gpurun_out/r02_mega4_b16 python tools/ncu_ar.py > gpurun_out/n_mega4.log 2>&1
echo "mega4 rc=$?"; ls -la gpurun_out/r02_mega4_b16.ncu-rep
ncu -i gpurun_out/r02_mega4_b16.ncu-rep --page raw --csv > gpurun_out/r02_mega4_b16_ncu_full_raw.csv 2>/dev/null
NSTEPS=2 timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:tc5v2 -s 60 -c 3 -o gpurun_out/r02_tc5v2 python tools/ncu_diff.py > gpurun_out/n_tc5v2.log 2>&1
echo "tc5v2 rc=$?"
ncu -i gpurun_out/r02_tc5v2.ncu-rep --page raw --csv > gpurun_out/r02_tc5v2_ncu_full_raw.csv 2>/dev/null
TTS_BENCH_CODES=4 TTS_BENCH_DIFF_STEPS=2 timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches_bench_reduced.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extra > gpurun_out/n_bench_under_ncu.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r02_launches_bench_reduced.csv
