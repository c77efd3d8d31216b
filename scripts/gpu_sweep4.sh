#!/bin/bash
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --no-e2e --no-cpu --no-probes > gpurun_out/sw4_$tag.json 2>gpurun_out/sw4.err; python -c "
import json; d=json.load(open('gpurun_out/sw4_$tag.json')); print('$tag', round(d['value']), round(d['ms_per_step'],2), d['host_phase_ms_per_step'], round(d.get('get_mops',0))); print('   ', d['step_ms'][:10], round(sum(d['step_ms'][10:26])/16,3))"; }
run base A=1
run chunk26 SMATRIX_CHUNK=67108864
run parts8 SMATRIX_PARTS_LOG2=8
run chunk26parts8 SMATRIX_CHUNK=67108864 SMATRIX_PARTS_LOG2=8
run slice16 SMATRIX_SLICE_LOG2=16 SMATRIX_PARTS_LOG2=8
