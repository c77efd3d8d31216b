#!/bin/bash
# point reads in directory-slice order: GPU parity of the three modes, the c2 line (carries the input-order A side),
# and the DRAM traffic of the read kernels on the full-scale table
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sliced or device_pointer" > gpurun_out/r2g_pytest_sliced.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2g_pytest_sliced.log
SMX_BENCH_GET_VARIANTS=6,10,14,18,26 python bench.py --steps 20 --warmup 3 > gpurun_out/r2g_bench_n1_steps20.json 2> gpurun_out/r2g_bench_n1_steps20.err; echo "c2/20 rc=$?"; tail -3 gpurun_out/r2g_bench_n1_steps20.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2g_bench_n1_steps20.json'))
print({k:d.get(k) for k in ('metric','value','ms_per_step','steps','get_mops','get_ms')}, d['parity']['mismatches'], d['checks'])
print(' get', d['roofline']['get']); print(' rs', d['roofline'].get('random_sector'))
print(' e2e',d['e2e']['value'],d['e2e']['get_mops'],d['e2e']['h2d_ceiling']['frac'])
PY
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,launch__grid_size
timeout 200 ncu --metrics $M --clock-control none -k regex:"k_get|k_gather|k_parts_prefix|k_partition_scatter|k_row_counts|k_rowlen" --launch-skip 30 --launch-count 56 --csv --log-file gpurun_out/r2g_dram_traffic_c2_reads.csv \
  python bench.py --steps 4 --warmup 1 --no-e2e --no-cpu --no-probes --no-parity > /dev/null 2> gpurun_out/r2g_traffic_c2_reads.err; echo "c2 read traffic rc=$?"
grep -c k_get gpurun_out/r2g_dram_traffic_c2_reads.csv
