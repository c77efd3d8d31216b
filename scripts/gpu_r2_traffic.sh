#!/bin/bash
# DRAM traffic + L2 hit rate of the dominant kernels at FULL scale (c2: k_upsert / k_get / getrow; c4: getrow kernels),
# launch lists (gpu__time_duration) of c2 / c4 steps, and one --set full capture each of k_get and k_getrow_chunks.
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,launch__grid_size
timeout 1500 ncu --metrics $M --clock-control none -k regex:"k_upsert|k_partition_scatter|k_partition_count" --launch-skip 150 --launch-count 24 --csv --log-file gpurun_out/r2_dram_traffic_c2_writes.csv \
  python bench.py --steps 4 --warmup 1 --no-e2e --no-cpu --no-probes --no-parity > gpurun_out/r2_traffic_c2.json 2> gpurun_out/r2_traffic_c2.err; echo "c2 write traffic rc=$?"
timeout 1500 ncu --metrics $M --clock-control none -k regex:"k_get|k_getrow|k_row_counts|k_rowlen" --csv --log-file gpurun_out/r2_dram_traffic_c2_reads.csv \
  python bench.py --steps 4 --warmup 1 --no-e2e --no-cpu --no-probes --no-parity > /dev/null 2> gpurun_out/r2_traffic_c2_reads.err; echo "c2 read traffic rc=$?"
timeout 900 ncu --metrics $M --clock-control none -k regex:"k_getrow|k_row_counts|k_rowlen|k_scan" --launch-skip 30 --launch-count 40 --csv --log-file gpurun_out/r2_dram_traffic_c4.csv \
  python bench.py --workload c4 --steps 4 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/r2_traffic_c4.json 2> gpurun_out/r2_traffic_c4.err; echo "c4 traffic rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_get$|k_get\(" --launch-skip 3 --launch-count 1 -o gpurun_out/r2_ncu_full_k_get \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-probes --no-parity > /dev/null 2> gpurun_out/r2_ncu_full_k_get.err; echo "k_get full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_upsert" --launch-skip 16 --launch-count 1 -o gpurun_out/r2_ncu_full_k_upsert_quarter \
  python bench.py --scale 0.25 --steps 2 --warmup 1 --no-e2e --no-cpu --no-probes --no-parity > /dev/null 2> gpurun_out/r2_ncu_full_k_upsert.err; echo "k_upsert full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_getrow_chunks" --launch-skip 4 --launch-count 1 -o gpurun_out/r2_ncu_full_k_getrow_chunks \
  python bench.py --workload c4 --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity > /dev/null 2> gpurun_out/r2_ncu_full_getrow.err; echo "getrow full rc=$?"
SMATRIX_CHUNK=67108864 python bench.py --steps 8 --no-e2e --no-cpu --no-probes --no-parity > gpurun_out/r2_chunk26.json 2> gpurun_out/r2_chunk26.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_chunk26.json'))
print('chunk 2^26: incr',round(d['value']),'kern/step',d['step_upsert_kernel_ms'][:5],'step',d['step_ms'][:6],d['host_phase_ms_per_step'])
PY
ls -la gpurun_out/*.ncu-rep
