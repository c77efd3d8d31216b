#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python bench.py --no-e2e --no-cpu --no-probes --gets 67108864 "$@" > gpurun_out/sw_$tag.json 2> gpurun_out/sw_$tag.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/sw_$tag.json')); r=d['roofline']
    print('$tag: incr',round(d['value']),' ms/step',round(d['ms_per_step'],2),' rounds',d['upsert_rounds'],' kern_ms/step',round(r['avg_launch_ms']*r['launches']/d['steps'],2),' dir_cap',d['table']['dir_cap'],' phases',d.get('host_phase_ms_per_step'))
except Exception as e: print('$tag failed',e, open('gpurun_out/sw_$tag.err').read()[-300:])
PY
}
SMATRIX_PARTS_LOG2=8 run p8_c26
SMATRIX_PARTS_LOG2=7 run p7_c26
SMATRIX_PARTS_LOG2=6 run p6_c26
SMATRIX_PARTS_LOG2=5 run p5_c26
SMATRIX_PARTS_LOG2=8 SMATRIX_CHUNK=16777216 run p8_c24
SMATRIX_PARTS_LOG2=7 SMATRIX_CHUNK=16777216 run p7_c24
SMATRIX_PARTS_LOG2=5 SMATRIX_CHUNK=16777216 run p5_c24
SMATRIX_PARTITION_MIN=4000000000 SMATRIX_CHUNK=16777216 run nopart_c24
SMATRIX_PARTS_LOG2=8 SMATRIX_CHUNK=33554432 run p8_c25
