#!/bin/bash
# second GPU session: full-scale bench, reference arm, ncu launch list + full capture of the top kernels
set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?"
tail -3 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
timeout 600 python bench.py --impl reference --steps 30 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
cat gpurun_out/bench_ref.json
# launch list: 1/4 scale, every kernel with its device time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv \
  python bench.py --steps 3 --warmup 1 --batch 16777216 --total-ops 503316480 --rows 3250000 --gets 33554432 --no-e2e --no-cpu --no-probes > gpurun_out/ncu_launches.log 2>&1; echo "ncu list exit $?"
# full capture of the update kernel late in the build and of the get kernel
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_upsert|k_get' -s 66 -c 4 -o gpurun_out/prof_r1 \
  python bench.py --steps 3 --warmup 1 --batch 16777216 --total-ops 503316480 --rows 3250000 --gets 33554432 --no-e2e --no-cpu --no-probes > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out
