#!/bin/bash
# 2 GPUs: the C router (IPC between processes) and the torch fallback, parity incl. sharded getrow
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "2-" > gpurun_out/r2_pytest_sharded_n2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest_sharded_n2.log
nvidia-smi topo -m | head -8; nproc; lscpu | grep -i numa
