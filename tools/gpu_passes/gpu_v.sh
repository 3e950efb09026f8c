#!/bin/bash
# round-2 final evidence pass (1 GPU): ncu --set full of the three hot kernels, launch lists, GPU test suite, bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
B=1 timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:mega3 -s 2 -c 1 -o gpurun_out/r02q_mega3 python tools/ncu_ar.py > gpurun_out/v_mega3.log 2>&1; echo "mega3 rc=$?"
ncu -i gpurun_out/r02q_mega3.ncu-rep --page raw --csv > gpurun_out/r02q_mega3_ncu_full_raw.csv 2>/dev/null
NSTEPS=2 timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:tc5v2 -s 60 -c 4 -o gpurun_out/r02q_tc5v2 python tools/ncu_diff.py > gpurun_out/v_tc5v2.log 2>&1; echo "tc5v2 rc=$?"
ncu -i gpurun_out/r02q_tc5v2.ncu-rep --page raw --csv > gpurun_out/r02q_tc5v2_ncu_full_raw.csv 2>/dev/null
NSTEPS=2 timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 10 -c 1 -o gpurun_out/r02q_attn_tc python tools/ncu_diff.py > gpurun_out/v_attn.log 2>&1; echo "attn rc=$?"
ncu -i gpurun_out/r02q_attn_tc.ncu-rep --page raw --csv > gpurun_out/r02q_attn_tc_S191_ncu_full_raw.csv 2>/dev/null
NSTEPS=2 timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02q_diffusion_step_launches.csv python tools/ncu_diff.py > gpurun_out/v_list.log 2>&1; echo "list rc=$?"
TTS_BENCH_CODES=4 TTS_BENCH_DIFF_STEPS=2 timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02q_launches_bench_reduced.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extra > gpurun_out/v_bench_under_ncu.log 2>&1; echo "bench list rc=$?"
rm -f gpurun_out/*.ncu-rep
timeout -s KILL 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r02q_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02q_pytest_gpu.log
timeout -s KILL 900 python bench.py > gpurun_out/r02q_bench_n1.json 2> gpurun_out/v_bench.err; tail -1 gpurun_out/v_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02q_bench_n1.json') if l.startswith('{')][-1])
print('value', d['value'], 'e2e', d['e2e']['value'], d['stage_ms'], 'launches', d['gpu_launches'])
print('roofline', d['roofline']['frac'], d['roofline']['us_per_launch'], d['roofline']['traffic'], 'tensor', d['roofline_tensor']['achieved'], d['roofline_tensor']['frac'])
print('c3', d['c3']['rtf_e2e'], d['c3']['ar_mel_tokens_per_s'])
PY
