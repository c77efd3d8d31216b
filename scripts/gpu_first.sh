#!/bin/bash
# first GPU session: smoke, parity tests, small bench
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --batch 16777216 --total-ops 167772160 --rows 1300000 --gets 33554432 --cpu-sample 4000000 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench exit $?"
tail -3 gpurun_out/bench_small.err; cat gpurun_out/bench_small.json
