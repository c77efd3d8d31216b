#!/bin/bash
# last validation of the final commit on one GPU
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu_final.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1_steps20.json 2> gpurun_out/r2_bench_n1_steps20.err; echo "c2/20 rc=$?"
python bench.py --workload c3 > gpurun_out/r2_bench_c3_n1.json 2> gpurun_out/r2_bench_c3_n1.err; echo "c3 rc=$?"
python - <<'PY'
import json
for f in ('r2_bench_n1_steps20','r2_bench_c3_n1'):
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f, {k:d.get(k) for k in ('metric','value','ms_per_step','steps','get_mops')}, d['parity']['mismatches'], d['checks'])
    print(' step_ms',d.get('step_ms')); print(' phases',d.get('host_phase_ms_per_step')); print(' e2e',d['e2e']['value'],d['e2e']['get_mops'],d['e2e']['h2d_ceiling']['frac'])
PY
