#!/bin/bash
# config 3 three times with the per-step phase series: is the first step (fresh table, everything new) stable now that a
# table with an arena takes its cudaMalloc'ed scratch at open?  (before: 13.4 / 13.8 / 34.5 / 165 ms on four runs)
for r in 1 2 3; do
  python bench.py --workload c3 --no-cpu --no-e2e --no-parity --phase-series > gpurun_out/r2h_bench_c3_n1_r$r.json 2> gpurun_out/r2h_bench_c3_n1_r$r.err; echo "c3 run $r rc=$?"
done
python - <<'PY'
import json
for r in (1,2,3):
    f='r2h_bench_c3_n1_r%d'%r
    d=json.load(open('gpurun_out/%s.json'%f))
    print(f, round(d['value']), round(d['get_mops']), d['step_ms'][:6], d['host_phase_ms_per_step'])
    print(open('gpurun_out/%s.err'%f).read().split('\n')[0])
PY
