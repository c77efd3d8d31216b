#!/bin/bash
# final build: DRAM traffic of the dominant kernels (c2 writes / reads, c3 writes), c4 line again (e2e fixed)
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,launch__grid_size
timeout 1500 ncu --metrics $M --clock-control none -k regex:"k_upsert|k_partition_scatter|k_partition_count" --launch-skip 75 --launch-count 12 --csv --log-file gpurun_out/r2f_dram_traffic_c2_writes.csv \
  python bench.py --steps 4 --warmup 1 --no-e2e --no-cpu --no-probes --no-parity > /dev/null 2> gpurun_out/r2f_traffic_c2_writes.err; echo "c2 write traffic rc=$?"
timeout 1500 ncu --metrics $M --clock-control none -k regex:"k_get|k_getrow|k_row_counts|k_rowlen" --csv --log-file gpurun_out/r2f_dram_traffic_c2_reads.csv \
  python bench.py --steps 4 --warmup 1 --no-e2e --no-cpu --no-probes --no-parity > /dev/null 2> gpurun_out/r2f_traffic_c2_reads.err; echo "c2 read traffic rc=$?"
timeout 1500 ncu --metrics $M --clock-control none -k regex:"k_upsert|k_migrate|k_partition_scatter" --launch-skip 120 --launch-count 40 --csv --log-file gpurun_out/r2f_dram_traffic_c3_writes.csv \
  python bench.py --workload c3 --steps 4 --warmup 1 --no-e2e --no-cpu --no-probes --no-parity > /dev/null 2> gpurun_out/r2f_traffic_c3_writes.err; echo "c3 write traffic rc=$?"
python bench.py --workload c4 > gpurun_out/r2_bench_c4_n1.json 2> gpurun_out/r2_bench_c4_n1.err; echo "c4 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_c4_n1.json'))
print({k:d.get(k) for k in ('metric','value','ms_per_step','rowlen_mops','build')}); print(d['e2e']); print(d['parity']['mismatches'], d['checks'])
PY
