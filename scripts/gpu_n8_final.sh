#!/bin/bash
mkdir -p gpurun_out
scripts/gpu_multi_bench_only.sh 8
timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q -k "8 and peer" > gpurun_out/pytest_sharded_8.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_sharded_8.log
tail -3 gpurun_out/pytest_sharded_8.log
