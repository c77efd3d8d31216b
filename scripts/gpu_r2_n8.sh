#!/bin/bash
# 8 GPUs: router parity at world 8 (C router), the C multi-GPU example, c2 weak scaling and c5
nvidia-smi topo -m > gpurun_out/r2_topo_8gpu.txt 2>&1; nproc; lscpu | grep -i "numa node" ; free -g | head -2
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "8-c" > gpurun_out/r2_pytest_sharded_n8.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_sharded_n8.log
timeout 600 python -m pytest tests/test_c_example.py -m gpu -x -q > gpurun_out/r2_pytest_cexample_n8.log 2>&1; echo "c example rc=$?"; tail -3 gpurun_out/r2_pytest_cexample_n8.log
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 "${@:2}"; }
run 29811 --steps 20 > gpurun_out/r2_bench_n8_c2.json 2> gpurun_out/r2_bench_n8_c2.err; echo "c2 rc=$?"; tail -3 gpurun_out/r2_bench_n8_c2.err
run 29812 --workload c5 > gpurun_out/r2_bench_n8_c5.json 2> gpurun_out/r2_bench_n8_c5.err; echo "c5 rc=$?"; tail -3 gpurun_out/r2_bench_n8_c5.err
python - <<'PY'
import json
for f in ('r2_bench_n8_c2','r2_bench_n8_c5'):
    try: d=json.load(open('gpurun_out/%s.json'%f))
    except Exception as e: print(f,'unreadable',e); continue
    print(f, {k:d.get(k) for k in ('metric','value','ms_per_step','steps','get_mops','parity','checks','nnz')})
    print(' step_ms',d.get('step_ms')); print(' kern',d.get('step_upsert_kernel_ms')); print(' phases',d.get('host_phase_ms_per_step'))
    r=d['roofline']; print(' roofline',{k:r.get(k) for k in ('achieved','frac','kernel_share_of_step')}, r.get('nvlink'))
    print(' e2e',d.get('e2e')); print(' cpu',d.get('cpu_baseline'))
PY
