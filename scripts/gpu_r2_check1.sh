#!/bin/bash
# round 2: parity on the GPU + C3-shape build + the default bench after the kernel changes
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_1.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu_1.log
python scripts/cf_scale.py 4000000 4000000 > gpurun_out/r2_c3_after1.txt 2>&1; tail -12 gpurun_out/r2_c3_after1.txt
python bench.py --steps 30 --warmup 3 > gpurun_out/r2_bench_full_1.json 2> gpurun_out/r2_bench_full_1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_full_1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_full_1.json'))
print({k:d[k] for k in ('value','ms_per_step','get_mops','checks','upsert_rounds','gpu_launches','table')})
print('step_ms',d['step_ms']); print('kern',d['step_upsert_kernel_ms']); print(d['host_phase_ms_per_step']); print(d['roofline']); print(d.get('e2e')); print(d.get('reads'))
PY
