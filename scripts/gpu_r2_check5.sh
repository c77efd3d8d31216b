#!/bin/bash
python -m pytest tests -m gpu -x -q -k "snapshot or getrow or java or quirks or big_row or cf" > gpurun_out/r2_pytest_gpu_3.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu_3.log
timeout 900 python bench.py --workload c4 > gpurun_out/r2_bench_c4_3.json 2> gpurun_out/r2_bench_c4_3.err; echo "c4 rc=$?"; tail -5 gpurun_out/r2_bench_c4_3.err
df -h /dev/shm /tmp | tail -2; free -g | head -2
timeout 600 python scripts/snapshot_scale.py 0.25 /dev/shm > gpurun_out/r2_snapshot_scale_025.json 2> gpurun_out/r2_snapshot_scale_025.err; echo "snap rc=$?"; cat gpurun_out/r2_snapshot_scale_025.json; tail -3 gpurun_out/r2_snapshot_scale_025.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_c4_3.json'))
print({k:d.get(k) for k in ('metric','value','ms_per_step','parity','checks','gpu_launches','rowlen_mops','build','step_ms')})
print(d['roofline']); print(d['e2e'])
PY
