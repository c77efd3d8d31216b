#!/bin/bash
# the complete config-3 line of the final build (parity, e2e, CPU side by side) + the GPU tests of the arena variant
python bench.py --workload c3 > gpurun_out/r2h_bench_c3_n1.json 2> gpurun_out/r2h_bench_c3_n1.err; echo "c3 rc=$?"
timeout 45 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "arena" > gpurun_out/r2h_pytest_arena.log 2>&1; echo "pytest arena rc=$?"; tail -2 gpurun_out/r2h_pytest_arena.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2h_bench_c3_n1.json'))
print({k:d.get(k) for k in ('metric','value','ms_per_step','steps','get_mops')}, d['parity']['mismatches'], d['checks'])
print(' step_ms',d.get('step_ms')); print(' e2e',d['e2e']['value'],d['e2e']['get_mops'],d['e2e']['h2d_ceiling']['frac']); print(' cpu', d['cpu_baseline'])
PY
