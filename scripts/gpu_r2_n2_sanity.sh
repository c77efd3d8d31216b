#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "2-c" > gpurun_out/r2_pytest_sharded_n2_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest_sharded_n2_final.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29951 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "c2 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n2.json'))
print({k:d.get(k) for k in ('value','ms_per_step','get_mops')}, d['parity']['mismatches'], d['parity']['ranks'], d['checks'], d['e2e']['value'], d['e2e']['h2d_ceiling']['frac'])
PY
