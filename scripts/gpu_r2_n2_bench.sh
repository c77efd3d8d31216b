#!/bin/bash
# 2 GPUs: C router performance on c2 (weak scaling, last 20 steps) and c5 (strong scaling)
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:2}"; }
run 29711 --steps 20 > gpurun_out/r2_bench_n2_c2.json 2> gpurun_out/r2_bench_n2_c2.err; echo "c2 rc=$?"; tail -3 gpurun_out/r2_bench_n2_c2.err
run 29712 --workload c5 --steps 20 > gpurun_out/r2_bench_n2_c5.json 2> gpurun_out/r2_bench_n2_c5.err; echo "c5 rc=$?"; tail -3 gpurun_out/r2_bench_n2_c5.err
python - <<'PY'
import json
for f in ('r2_bench_n2_c2','r2_bench_n2_c5'):
    try: d=json.load(open('gpurun_out/%s.json'%f))
    except Exception as e: print(f,'unreadable',e); continue
    print(f, {k:d.get(k) for k in ('metric','value','ms_per_step','get_mops','parity','checks','upsert_rounds','table','nnz')})
    print(' step_ms',d.get('step_ms')); print(' kern',d.get('step_upsert_kernel_ms')); print(' phases',d.get('host_phase_ms_per_step'))
    r=d['roofline']; print(' roofline',{k:r.get(k) for k in ('achieved','frac','kernel_share_of_step')}, r.get('nvlink'))
    print(' e2e',d.get('e2e')); print(' cpu',d.get('cpu_baseline'))
PY
