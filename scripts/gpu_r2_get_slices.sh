#!/bin/bash
# point reads in directory-slice order + 256 slices for write chunks without column 0: GPU parity, the c2 line with the
# A/B variants (set_get_slices values) and the batch-size sweep, the c2 build with SMATRIX_WIDE_SLICES=1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sliced or column0 or device_pointer" > gpurun_out/r2g_pytest_sliced.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2g_pytest_sliced.log
SMX_BENCH_GET_VARIANTS=2,6,10,34,42 SMX_BENCH_GET_BATCHES=1048576,4194304,8388608,16777216,33554432 python bench.py --steps 20 --warmup 3 > gpurun_out/r2g_bench_n1_steps20.json 2> gpurun_out/r2g_bench_n1_steps20.err; echo "c2/20 rc=$?"; tail -3 gpurun_out/r2g_bench_n1_steps20.err
SMATRIX_WIDE_SLICES=1 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-parity --no-probes > gpurun_out/r2g_bench_n1_wide.json 2> gpurun_out/r2g_bench_n1_wide.err; echo "c2/20 wide rc=$?"; tail -3 gpurun_out/r2g_bench_n1_wide.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2g_bench_n1_steps20.json'))
print({k:d.get(k) for k in ('metric','value','ms_per_step','steps','get_mops','get_ms')}, d['parity']['mismatches'], d['checks'])
print(' step_ms', d['step_ms']); print(' kern', d['step_upsert_kernel_ms']); print(' phases', d['host_phase_ms_per_step'])
g=d['roofline']['get']; print(' get', {k:g[k] for k in g if k not in ('variants','batch_sweep_mops','kernel')})
for k,v in g.get('variants',{}).items(): print('   variant',k,v)
print('   sweep', g.get('batch_sweep_mops'))
print(' e2e',d['e2e']['value'],d['e2e']['get_mops'],d['e2e']['h2d_ceiling']['frac'])
w=json.load(open('gpurun_out/r2g_bench_n1_wide.json'))
print('WIDE', {k:w.get(k) for k in ('value','ms_per_step','get_mops')}, w['checks'])
print(' step_ms', w['step_ms']); print(' kern', w['step_upsert_kernel_ms']); print(' phases', w['host_phase_ms_per_step'])
PY
