#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_tiles.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu_tiles.log
for t in 1 0; do
  SMATRIX_MIGRATE_TILES=$t python bench.py --workload c3 --no-cpu --no-e2e --no-probes > gpurun_out/r2_bench_c3_tiles$t.json 2> gpurun_out/r2_bench_c3_tiles$t.err; echo "c3 tiles=$t rc=$?"
  python - $t <<'PY'
import json,sys
d=json.load(open('gpurun_out/r2_bench_c3_tiles%s.json'%sys.argv[1]))
print('tiles',sys.argv[1],{k:d.get(k) for k in ('value','ms_per_step')}, d['parity']['mismatches'], d['checks']['value_sum_ok'])
print(' step_ms',d.get('step_ms')); print(' phases',d.get('host_phase_ms_per_step'))
PY
done
python bench.py --workload c4 --no-cpu --no-e2e > gpurun_out/r2_bench_c4_tiles.json 2> gpurun_out/r2_bench_c4_tiles.err; echo "c4 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_c4_tiles.json'))
print({k:d.get(k) for k in ('value','ms_per_step','build')}, d['parity']['mismatches'], d['checks'])
PY
