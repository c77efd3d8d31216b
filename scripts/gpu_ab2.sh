#!/bin/bash
run() { tag=$1; for i in 1 2; do timeout 600 python bench.py --no-e2e --no-cpu --no-probes > gpurun_out/ab2_$tag.json 2>gpurun_out/ab2.err; python -c "
import json; d=json.load(open('gpurun_out/ab2_$tag.json')); print('$tag', round(d['value']), round(d['ms_per_step'],2), d['host_phase_ms_per_step']); print('   ', d['step_ms'][:10], sum(d['step_ms'][10:26])/16)"; done; }
run regs32
SMX_NVCC_EXTRA="-DSMX_UPSERT_MIN_BLOCKS=6" python -m libsmatrix_b200.build --force > /dev/null 2>&1
run regs40
