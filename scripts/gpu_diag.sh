#!/bin/bash
mkdir -p gpurun_out
SMALL="--steps 3 --warmup 1 --batch 16777216 --total-ops 117440512 --rows 1300000 --gets 16777216 --no-e2e --no-cpu --no-probes"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/diag_launches.csv python bench.py $SMALL > gpurun_out/diag1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_upsert' -s 22 -c 2 -o gpurun_out/diag_prof python bench.py $SMALL > gpurun_out/diag2.log 2>&1
ls -la gpurun_out | tail -5
