#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python bench.py --no-e2e --no-cpu --no-probes --gets 67108864 "$@" > gpurun_out/sw_$tag.json 2> gpurun_out/sw_$tag.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/sw_$tag.json')); r=d['roofline']
    print('$tag: incr',round(d['value']),' ms/step',round(d['ms_per_step'],2),' rounds',d['upsert_rounds'],' slabGB',round(d['table']['slab_bytes']/1e9,1),' phases',d.get('host_phase_ms_per_step'))
    print('   steps',d['step_ms'])
except Exception as e: print('$tag failed',e, open('gpurun_out/sw_$tag.err').read()[-300:])
PY
}
SMATRIX_CHUNK=16777216 run pro_c24
SMATRIX_CHUNK=33554432 run pro_c25
run pro_c26
SMATRIX_PROACTIVE=0 SMATRIX_CHUNK=16777216 run nopro_c24
