#!/bin/bash
# N=2: default fused route (with IPC warm-up) vs pipelined route (piece = one 2^25 chunk)
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --no-e2e --no-cpu --no-probes > gpurun_out/n2_$2.json 2> gpurun_out/n2_$2.err; }
run 29511 base
SMX_PIPELINE_MIN=67108864 SMX_PIPELINE_PIECE=33554432 run 29512 pipe
SMX_ROUTE_DEBUG=1 run 29513 dbg
python - <<'PY'
import json
for t in ("base","pipe","dbg"):
    try:
        d=json.load(open(f"gpurun_out/n2_{t}.json")); print(t, round(d["value"]), d["ms_per_step"], d.get("get_mops"), d["step_ms"])
    except Exception as e: print(t, "failed", e)
PY
grep "\[route\]" gpurun_out/n2_dbg.err | tail -4
