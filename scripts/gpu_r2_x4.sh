#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_x4.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu_x4.log
python bench.py --workload c3 --no-cpu > gpurun_out/r2_bench_c3_x4.json 2> gpurun_out/r2_bench_c3_x4.err; echo "c3 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_c3_x4.json'))
print({k:d.get(k) for k in ('metric','value','ms_per_step','get_mops')}, d['parity']['mismatches'], d['checks'], d['table'])
print(' step_ms',d.get('step_ms')); print(' phases',d.get('host_phase_ms_per_step'))
PY
