#!/bin/bash
# usage: gpu_multi.sh N  — NCCL sharded parity test + N-GPU bench
N=$1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/pytest_sharded_$N.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_sharded_$N.log
tail -5 gpurun_out/pytest_sharded_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $?"
tail -5 gpurun_out/bench_n$N.err | cut -c1-300
cat gpurun_out/bench_n$N.json | cut -c1-1500
