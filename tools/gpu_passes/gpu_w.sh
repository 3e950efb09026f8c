#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
SS=191 timeout -s KILL 200 python tools/tc5_trace.py 2>&1 | grep -E "^S|csz" | tail -3
US=1,8 timeout -s KILL 300 python tools/diff_batch_times.py 2>&1 | tail -2
L=300 US=1 STEPS=10 timeout -s KILL 300 python tools/diff_batch_times.py 2>&1 | tail -1
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout -s KILL 300 python -m pytest tests/test_diffusion_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2
