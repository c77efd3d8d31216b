#!/bin/bash
mkdir -p gpurun_out
ARGS="--steps 3 --warmup 1 --batch 16777216 --total-ops 503316480 --rows 3250000 --gets 33554432 --no-e2e --no-cpu --no-probes --arena-gib 14"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_upsert|k_get|k_partition_scatter|k_migrate' -s 118 -c 12 -o gpurun_out/r1_prof_quarter python bench.py $ARGS > gpurun_out/r1_ncu_full.log 2>&1; echo "ncu full exit $?"
tail -3 gpurun_out/r1_ncu_full.log | cut -c1-200
ls -la gpurun_out/r1_prof_quarter*
