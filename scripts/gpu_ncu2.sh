#!/bin/bash
mkdir -p gpurun_out
ARGS="--steps 3 --warmup 1 --batch 16777216 --total-ops 503316480 --rows 3250000 --gets 33554432 --no-e2e --no-cpu --no-probes"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_upsert' -s 84 -c 6 -o gpurun_out/prof_r1b python bench.py $ARGS > gpurun_out/ncu2.log 2>&1; echo "ncu exit $?"
tail -2 gpurun_out/ncu2.log | cut -c1-300
