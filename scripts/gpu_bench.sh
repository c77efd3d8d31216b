#!/bin/bash
# usage: gpu_bench.sh TAG [extra bench args]  — parity tests, then the full-scale bench
TAG=$1; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$TAG.log
tail -4 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
r=d['roofline']
print('incr Mops/s',round(d['value']),'ms/step',round(d['ms_per_step'],2),'get Mops/s',round(d['get_mops']),'rounds',d['upsert_rounds'],'launches',d['gpu_launches'])
print('kernel share',r['kernel_share_of_step'],'avg launch ms',r['avg_launch_ms'],'frac',r['frac'],'rand',r.get('random_sector',{}).get('incr_frac'))
print('table',d['table'],'nnz',d['nnz'])
print('e2e',d.get('e2e'))
print('cpu',d.get('cpu_baseline',{}).get('value'))
PY
