#!/bin/bash
N=$1
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29941 bench.py --gpus $N --workload c3 > gpurun_out/r2_bench_c3_n$N.json 2> gpurun_out/r2_bench_c3_n$N.err; echo "c3 rc=$?"
python - $N <<'PY'
import json,sys
d=json.load(open('gpurun_out/r2_bench_c3_n%s.json'%sys.argv[1]))
print({k:d.get(k) for k in ('metric','value','ms_per_step','steps','get_mops','nnz')}, d['parity']['mismatches'], d['parity']['ranks'], d['checks'])
print(d['step_ms']); print(d['roofline']['nvlink']); print(d['e2e']['value'], d['e2e']['h2d_ceiling'])
PY
