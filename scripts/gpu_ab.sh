#!/bin/bash
# A/B: partitioning on/off at small and full scale (no e2e, no cpu, no probes)
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python bench.py --no-e2e --no-cpu --no-probes "$@" > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/ab_$tag.json')); r=d['roofline']
    print('$tag: incr',round(d['value']),'Mops/s  ms/step',round(d['ms_per_step'],2),' get',round(d['get_mops']),' rounds',d['upsert_rounds'],' launches',d['gpu_launches'],' kshare',round(r['kernel_share_of_step'],3),' avg_launch_ms',round(r['avg_launch_ms'],3),' slabGB',round(d['table']['slab_bytes']/1e9,1),' grows',d['table']['row_grows'],' phases',d.get('host_phase_ms_per_step'))
except Exception as e: print('$tag failed',e, open('gpurun_out/ab_$tag.err').read()[-500:])
PY
}
SMALL="--steps 10 --warmup 3 --batch 16777216 --total-ops 167772160 --rows 1300000 --gets 33554432"
SMATRIX_PARTITION_MIN=4000000000 run small_nopart $SMALL
run small_part $SMALL
SMATRIX_PARTITION_MIN=4000000000 run full_nopart
