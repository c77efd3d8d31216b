#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/pytest_sharded_2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_sharded_2.log
tail -3 gpurun_out/pytest_sharded_2.log
scripts/gpu_multi_bench_only.sh 2
