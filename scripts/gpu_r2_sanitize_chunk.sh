#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_4.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_gpu_4.log
compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitize_small.py > gpurun_out/r2_sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2_sanitize_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/sanitize_small.py > gpurun_out/r2_sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2_sanitize_racecheck.log
for c in 33554432 67108864; do
  SMATRIX_CHUNK=$c python bench.py --no-e2e --no-cpu --no-probes --no-parity > gpurun_out/r2_chunk_full_$c.json 2> gpurun_out/r2_chunk_full_$c.err
  python - $c <<'PY'
import json,sys
d=json.load(open('gpurun_out/r2_chunk_full_%s.json'%sys.argv[1]))
print('chunk',sys.argv[1],'whole build',round(d['value']),'get',round(d['get_mops']),'step',d['step_ms'])
PY
done
