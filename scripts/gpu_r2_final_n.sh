#!/bin/bash
# usage: gpu_r2_final_n.sh N workload...   — final multi-GPU lines (+ sharded parity tests at that world size)
N=$1; shift
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_c_example.py -m gpu -x -q -k "$N- or multi_gpu" > gpurun_out/r2_pytest_sharded_n$N.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_sharded_n$N.log
port=29900
for w in "$@"; do
  port=$((port+1))
  if [ "$w" = "c2" ]; then extra="--steps 20"; out=r2_bench_n$N; else extra=""; out=r2_bench_${w}_n$N; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --workload $w --warmup 3 $extra > gpurun_out/$out.json 2> gpurun_out/$out.err; echo "$w rc=$?"
  python - $out <<'PY'
import json,sys
try: d=json.load(open('gpurun_out/%s.json'%sys.argv[1]))
except Exception as e: print(sys.argv[1],'unreadable',e); sys.exit(0)
print(sys.argv[1], {k:d.get(k) for k in ('metric','value','ms_per_step','steps','get_mops','rowlen_mops','nnz')})
print(' parity',d['parity']['mismatches'],d['parity']['ranks'],'checks',d['checks'])
print(' step_ms',d.get('step_ms'))
r=d['roofline']; print(' roofline',{k:r.get(k) for k in ('achieved','frac','kernel_share_of_step')}, r.get('nvlink'))
e=d.get('e2e') or {}; print(' e2e',{k:e.get(k) for k in ('value','get_mops','ms_per_step','pcie_frac','h2d_ceiling')}); print(' cpu',(d.get('cpu_baseline') or {}).get('value'))
PY
done
