#!/bin/bash
# round 2, first call: where does the Zipf/CF build (config 3 shape) spend its time on the round-1 build?
set -x
nproc; lscpu | grep -i -E "numa|model name|socket" ; nvidia-smi topo -m 2>&1 | head -20
python scripts/cf_scale.py 4000000 4000000 > gpurun_out/r2_c3_baseline.txt 2>&1
tail -15 gpurun_out/r2_c3_baseline.txt
