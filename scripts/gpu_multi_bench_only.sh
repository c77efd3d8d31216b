#!/bin/bash
N=$1
mkdir -p gpurun_out
SMX_ROUTE_DEBUG=${SMX_ROUTE_DEBUG:-1} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $?"
grep -E "libsmatrix error|Error" gpurun_out/bench_n$N.err | head -5
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_n$N.json") if l.startswith("{")][0])
print("n$N incr",round(d["value"]),"ms/step",round(d["ms_per_step"],2),"get",round(d["get_mops"]),"get_ms",round(d["get_ms"],1),"nnz",d["nnz"])
print("  steps",d["step_ms"])
PY
grep "\[route\]" gpurun_out/bench_n$N.err | tail -${TAILN:-3}
