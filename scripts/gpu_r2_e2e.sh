#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest_gpu_final.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1_steps20.json 2> gpurun_out/r2_bench_n1_steps20.err; echo "c2/20 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1_steps20.json'))
print({k:d.get(k) for k in ('value','ms_per_step','get_mops')}, d['parity']['mismatches'])
e=d['e2e']; print({k:e.get(k) for k in ('value','get_mops','ms_per_step','pcie_frac')}, e['h2d_ceiling'], e['step_ms'][:6])
PY
