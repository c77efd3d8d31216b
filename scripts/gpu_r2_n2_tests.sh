#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_c_example.py -m gpu -x -q -k "2- or multi_gpu" > gpurun_out/r2_pytest_sharded_n2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest_sharded_n2.log
