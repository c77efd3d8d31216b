#!/bin/bash
# (a) how much does a coarser slice partition cost k_upsert?  (b) snapshot save/load at full config-2 scale
for pl in 7 5 4; do
  SMATRIX_PARTS_LOG2=$pl python bench.py --steps 8 --no-e2e --no-cpu --no-probes --no-parity > gpurun_out/r2_parts_$pl.json 2> gpurun_out/r2_parts_$pl.err
  python - $pl <<'PY'
import json,sys
d=json.load(open('gpurun_out/r2_parts_%s.json'%sys.argv[1]))
print('parts_log2',sys.argv[1],'incr',round(d['value']),'kern/step',d['step_upsert_kernel_ms'][:6],'step',d['step_ms'][:6],d['host_phase_ms_per_step'])
PY
done
timeout 900 python scripts/snapshot_scale.py 1.0 /dev/shm > gpurun_out/r2_snapshot_scale_full.json 2> gpurun_out/r2_snapshot_scale_full.err; echo "snap rc=$?"; cat gpurun_out/r2_snapshot_scale_full.json; tail -3 gpurun_out/r2_snapshot_scale_full.err
