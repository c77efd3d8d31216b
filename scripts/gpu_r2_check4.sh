#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu_2.log
timeout 900 python bench.py --workload c4 > gpurun_out/r2_bench_c4_2.json 2> gpurun_out/r2_bench_c4_2.err; echo "c4 rc=$?"; tail -5 gpurun_out/r2_bench_c4_2.err
python bench.py --steps 8 --no-e2e --no-cpu --no-probes > gpurun_out/r2_bench_c2_3.json 2> gpurun_out/r2_bench_c2_3.err; echo "c2 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_c4_2.json'))
print({k:d.get(k) for k in ('metric','value','ms_per_step','parity','checks','gpu_launches','rowlen_mops','build','step_ms')})
print(d['roofline']); print(d['e2e'])
d=json.load(open('gpurun_out/r2_bench_c2_3.json'))
print({k:d.get(k) for k in ('value','get_mops','reads','parity')})
PY
