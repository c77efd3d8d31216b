#!/bin/bash
# A/B of the L2 fetch granularity (compile-time), then the new bench on c2 / c3 / c4(scaled)
for v in "" _all128 _all64; do
  SMATRIX_B200_LIB=$PWD/libsmatrix_b200/lib/libsmatrix_b200$v.so python bench.py --steps 8 --no-e2e --no-cpu --no-probes --no-parity > gpurun_out/r2_ab_fetch$v.json 2> gpurun_out/r2_ab_fetch$v.err
  python - "$v" <<'PY'
import json,sys
d=json.load(open('gpurun_out/r2_ab_fetch%s.json'%sys.argv[1]))
print('variant',repr(sys.argv[1]),'incr',round(d['value']),'get',round(d['get_mops']),'kern/step',d['step_upsert_kernel_ms'][-3:],'step',d['step_ms'][-3:], 'reads', {k:round(v,1) for k,v in d['reads'].items() if 'mops' in k or 'gpairs' in k})
PY
done
python bench.py > gpurun_out/r2_bench_c2_2.json 2> gpurun_out/r2_bench_c2_2.err; echo "c2 rc=$?"; tail -3 gpurun_out/r2_bench_c2_2.err
python bench.py --workload c3 --phase-series > gpurun_out/r2_bench_c3_2.json 2> gpurun_out/r2_bench_c3_2.err; echo "c3 rc=$?"; tail -20 gpurun_out/r2_bench_c3_2.err
python bench.py --workload c4 --scale 0.1 > gpurun_out/r2_bench_c4_s01.json 2> gpurun_out/r2_bench_c4_s01.err; echo "c4 rc=$?"; tail -5 gpurun_out/r2_bench_c4_s01.err
python - <<'PY'
import json
for f in ('r2_bench_c2_2','r2_bench_c3_2','r2_bench_c4_s01'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
    except Exception as e:
        print(f,'unreadable',e); continue
    print(f, {k:d.get(k) for k in ('metric','value','ms_per_step','get_mops','parity','checks','upsert_rounds','gpu_launches','table','rowlen_mops','build')})
    print(' step_ms',d.get('step_ms')); print(' kern',d.get('step_upsert_kernel_ms')); print(' phases',d.get('host_phase_ms_per_step'))
    r=d['roofline']; print(' roofline',{k:r.get(k) for k in ('achieved','frac','kernel_share_of_step','avg_launch_ms','step_level_gbs')}, r.get('random_sector'))
    print(' e2e',d.get('e2e')); print(' cpu',d.get('cpu_baseline'))
PY
