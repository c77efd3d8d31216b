#!/bin/bash
# single GPU, final build: smoke, parity tests, the bench lines for profiles/
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu_final.log
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_reference_n1.json 2> gpurun_out/r2_bench_reference_n1.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1_steps20.json 2> gpurun_out/r2_bench_n1_steps20.err; echo "c2/20 rc=$?"
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "c2 rc=$?"
python bench.py --workload c3 > gpurun_out/r2_bench_c3_n1.json 2> gpurun_out/r2_bench_c3_n1.err; echo "c3 rc=$?"
python bench.py --workload c4 > gpurun_out/r2_bench_c4_n1.json 2> gpurun_out/r2_bench_c4_n1.err; echo "c4 rc=$?"
python - <<'PY'
import json
for f in ('r2_bench_n1_steps20','r2_bench_n1','r2_bench_c3_n1','r2_bench_c4_n1'):
    try: d=json.load(open('gpurun_out/%s.json'%f))
    except Exception as e: print(f,'unreadable',e); continue
    print(f, {k:d.get(k) for k in ('metric','value','ms_per_step','steps','get_mops','upsert_rounds','gpu_launches','table','rowlen_mops','build')})
    print(' parity',{k:d['parity'][k] for k in ('mismatches','gets','rowlens','getrow_pairs')}, 'checks',d['checks'])
    print(' step_ms',d.get('step_ms')); print(' kern',d.get('step_upsert_kernel_ms')); print(' phases',d.get('host_phase_ms_per_step'))
    r=d['roofline']; print(' roofline',{k:r.get(k) for k in ('achieved','frac','kernel_share_of_step','avg_launch_ms','traffic','step_level_gbs')}, r.get('random_sector'), r.get('get'))
    print(' e2e',d.get('e2e')); print(' cpu',d.get('cpu_baseline')); print(' reads',d.get('reads')); print(' clocks',d.get('clocks'))
print(open('gpurun_out/r2_bench_reference_n1.json').read()[:700])
PY
