#!/bin/bash
for t in 1048576 1000000000; do
  SMATRIX_SHARD_TAPER_MIN=$t timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29961 bench.py --gpus 2 --steps 12 --warmup 3 --no-cpu --no-probes --no-parity > gpurun_out/r2_n2_taper_$t.json 2> gpurun_out/r2_n2_taper_$t.err; echo "rc=$?"
  python - $t <<'PY'
import json,sys
d=json.load(open('gpurun_out/r2_n2_taper_%s.json'%sys.argv[1])); e=d['e2e']
print('taper_min',sys.argv[1],'e2e',round(e['value']),'get',round(e['get_mops']),'ms/step',round(e['ms_per_step'],2),e['step_ms'][:8])
PY
done
