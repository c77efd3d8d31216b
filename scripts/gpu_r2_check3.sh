#!/bin/bash
# c4 at full scale; c2 phase series; ncu launch list of a c3 build (quarter scale)
timeout 900 python bench.py --workload c4 > gpurun_out/r2_bench_c4_1.json 2> gpurun_out/r2_bench_c4_1.err; echo "c4 rc=$?"; tail -5 gpurun_out/r2_bench_c4_1.err
python bench.py --phase-series --no-e2e --no-cpu --no-probes --no-parity > gpurun_out/r2_bench_c2_phase.json 2> gpurun_out/r2_bench_c2_phase.err; grep "^step" gpurun_out/r2_bench_c2_phase.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,launch__grid_size \
  --clock-control none -c 260 --csv --log-file gpurun_out/r2_c3_ncu2.csv \
  python bench.py --workload c3 --scale 0.25 --steps 4 --warmup 1 --no-e2e --no-cpu --no-probes --no-parity > gpurun_out/r2_c3_ncu2.json 2> gpurun_out/r2_c3_ncu2.err; echo "ncu rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_c4_1.json'))
print({k:d.get(k) for k in ('metric','value','ms_per_step','parity','checks','gpu_launches','table','rowlen_mops','build','step_ms')})
print(d['roofline']); print(d['e2e']); print(d['cpu_baseline'])
PY
