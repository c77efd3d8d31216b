#!/bin/bash
# round-1 evidence: parity, full bench (both arms), ncu launch list of the bench command, dram traffic of
# the dominant kernel at full scale, and a full-set capture at 1/4 scale (replay of a 60 GB table x40 is impractical)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/r1_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r1_pytest_gpu.log
tail -3 gpurun_out/r1_pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 30 --warmup 5 > gpurun_out/r1_bench_reference.json 2> gpurun_out/r1_bench_reference.err
timeout 900 python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; echo "bench exit $?"
tail -2 gpurun_out/r1_bench.err
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r1_launches.csv \
  python bench.py --no-e2e --no-cpu --no-probes > gpurun_out/r1_ncu_launches.log 2>&1; echo "ncu list exit $?"
timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
  -k regex:'k_upsert|k_get' -s 60 -c 32 --csv --log-file gpurun_out/r1_traffic_full.csv \
  python bench.py --no-e2e --no-cpu --no-probes --gets 134217728 > gpurun_out/r1_ncu_traffic.log 2>&1; echo "ncu traffic exit $?"
ARGS="--steps 3 --warmup 1 --batch 16777216 --total-ops 503316480 --rows 3250000 --gets 33554432 --no-e2e --no-cpu --no-probes --arena-gib 16"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_upsert' -s 64 -c 6 -o gpurun_out/r1_prof_quarter_upsert python bench.py $ARGS > gpurun_out/r1_ncu_full.log 2>&1; echo "ncu full upsert exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_get$|k_getrow_fill|k_rowlen' -c 4 -o gpurun_out/r1_prof_quarter_reads python bench.py $ARGS >> gpurun_out/r1_ncu_full.log 2>&1; echo "ncu full reads exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_partition_scatter|k_migrate' -s 50 -c 6 -o gpurun_out/r1_prof_quarter_aux python bench.py $ARGS >> gpurun_out/r1_ncu_full.log 2>&1; echo "ncu full aux exit $?"
ls -la gpurun_out | grep r1_
