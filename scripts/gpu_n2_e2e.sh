#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q -k "2" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench exit $?"
grep -E "libsmatrix error|Error|Traceback" gpurun_out/bench_n2.err | head -5
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_n2.json") if l.startswith("{")][0])
print("n2 incr",round(d["value"]),"get",round(d["get_mops"]),"e2e",d.get("e2e"))
PY
