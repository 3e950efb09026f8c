import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _pkg
pkg = _pkg.import_pkg()
eng = pkg.Engine(dtype=pkg.DTYPE_F16, max_batch=1, max_positions=64)
per = 8 << 20
ms, by = eng.bench_stream(1, 16384, 1, per); print(f"LDG.128 plain        : {by/ms/1e6:8.0f} GB/s")
for stage in (4096, 8192, 16384, 32768, 65536):
    for stages in (2, 3, 4, 6, 8, 12):
        if stage * stages > 200 * 1024: continue
        ms, by = eng.bench_stream(0, stage, stages, per)
        print(f"bulk ring {stage:6d} B x {stages:2d}: {by/ms/1e6:8.0f} GB/s", flush=True)
eng.close()
