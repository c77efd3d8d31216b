#!/bin/bash
# usage: gpu_multi.sh N  — NCCL sharded parity test + N-GPU bench (+ single-GPU sanity on the same box)
N=$1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/pytest_sharded_$N.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_sharded_$N.log
tail -3 gpurun_out/pytest_sharded_$N.log
timeout 600 python bench.py --no-e2e --no-cpu --no-probes > gpurun_out/bench_n1_same_box.json 2> gpurun_out/bench_n1_same_box.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $?"
grep -E "libsmatrix error|Error" gpurun_out/bench_n$N.err | head -5
python - <<PY
import json
for tag in ("n1_same_box","n$N"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/bench_{tag}.json") if l.startswith("{")][0])
        print(tag,"incr",round(d["value"]),"ms/step",round(d["ms_per_step"],2),"get",round(d["get_mops"]),"get_ms",round(d["get_ms"],1),"nnz",d["nnz"])
        print("  steps",d["step_ms"])
    except Exception as e: print(tag,"failed",e)
PY
