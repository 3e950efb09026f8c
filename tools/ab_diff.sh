#!/bin/bash
# A/B of the diffusion GEMM generations on the GPU box: parity tests, then the bench per variant.
timeout 300 python -m pytest tests/test_diffusion_gpu.py tests/test_ar_gpu.py tests/test_vocoder_gpu.py -m gpu -x -q 2>&1 | tail -6
for v in 0 1; do
  echo "== TTS_TC5_V1=$v"
  TTS_TC5_V1=$v timeout 150 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],d['stage_ms'])"
done
