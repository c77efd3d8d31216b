#!/bin/bash
# both arms back to back on one box, plus the steady-state variant
mkdir -p gpurun_out
timeout 600 python bench.py --impl reference --steps 30 --warmup 5 > gpurun_out/r1_bench_reference.json 2> gpurun_out/r1_bench_reference.err
timeout 900 python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; echo "bench exit $?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r1_bench_steps10.json 2> gpurun_out/r1_bench_steps10.err; echo "bench10 exit $?"
python - <<'PY'
import json
for f in ("r1_bench_reference","r1_bench","r1_bench_steps10"):
    d=json.load(open(f"gpurun_out/{f}.json")); e=d.get("e2e",{})
    print(f, round(d["value"],2), d["unit"], "ms/step", round(d["ms_per_step"],2), "get", d.get("get_mops"), "e2e", e.get("value"), e.get("get_mops"), "roof", (d.get("roofline") or {}).get("frac"), ((d.get("roofline") or {}).get("random_sector") or {}).get("incr_frac"))
PY
