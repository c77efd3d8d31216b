#!/bin/bash
SMATRIX_DEBUG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29931 bench.py --gpus 2 --workload c5 > gpurun_out/r2_bench_c5_n2.json 2> gpurun_out/r2_bench_c5_n2.err; echo "rc=$?"
grep -n "smatrix\]\|libsmatrix error" gpurun_out/r2_bench_c5_n2.err | head -60
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2_bench_c5_n2.json')); print({k:d.get(k) for k in ('metric','value','ms_per_step','steps','get_mops','nnz')}, d['parity']['mismatches'], d['checks'], d['table']); print(d['e2e']['value'], d['e2e']['h2d_ceiling'])
except Exception as e: print('unreadable', e)
PY
