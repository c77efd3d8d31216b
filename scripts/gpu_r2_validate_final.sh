#!/bin/bash
# validation of the final round-2 build on one GPU (slice-ordered reads, 256-slice write chunks): every GPU test, smoke(),
# the c2 line as the driver runs it, DRAM traffic of the write and read kernels on the full-scale table, c3, the whole
# build from empty, memcheck.  Most important first: the call's time limit may cut the tail.
t0=$(date +%s)
python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest_gpu.log 2>&1; echo "pytest rc=$? ($(( $(date +%s)-t0 )) s)"; tail -3 gpurun_out/r2h_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2h_smoke.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2h_bench_n1_steps20.json 2> gpurun_out/r2h_bench_n1_steps20.err; echo "c2/20 rc=$? ($(( $(date +%s)-t0 )) s)"
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,launch__grid_size
timeout 150 ncu --metrics $M --clock-control none -k regex:"k_upsert|k_partition_scatter|k_partition_count|k_get|k_gather|k_parts_prefix" --launch-skip 78 --launch-count 80 --csv --log-file gpurun_out/r2h_dram_traffic_c2.csv \
  python bench.py --steps 4 --warmup 1 --no-e2e --no-cpu --no-probes --no-parity > /dev/null 2> gpurun_out/r2h_traffic_c2.err; echo "c2 traffic rc=$? ($(( $(date +%s)-t0 )) s)"
python bench.py --workload c3 > gpurun_out/r2h_bench_c3_n1.json 2> gpurun_out/r2h_bench_c3_n1.err; echo "c3 rc=$? ($(( $(date +%s)-t0 )) s)"
python bench.py --steps 30 --warmup 3 --no-cpu --no-e2e --no-parity --no-probes > gpurun_out/r2h_bench_n1_whole.json 2> gpurun_out/r2h_bench_n1_whole.err; echo "c2/30 rc=$? ($(( $(date +%s)-t0 )) s)"
timeout 120 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitize_small.py > gpurun_out/r2h_sanitize_memcheck.log 2>&1; echo "memcheck rc=$? ($(( $(date +%s)-t0 )) s)"; tail -3 gpurun_out/r2h_sanitize_memcheck.log
python - <<'PY'
import json
for f in ('r2h_bench_n1_steps20','r2h_bench_c3_n1','r2h_bench_n1_whole'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, {k:d.get(k) for k in ('metric','value','ms_per_step','steps','get_mops')}, (d.get('parity') or {}).get('mismatches'), d['checks'])
    print(' step_ms',d.get('step_ms')); print(' phases',d.get('host_phase_ms_per_step'), 'share', d['roofline'].get('kernel_share_of_step'))
    g=d['roofline'].get('get') or {}; print(' get', g.get('achieved'), g.get('sliced_fraction'), g.get('input_order'))
    if d.get('e2e'): print(' e2e',d['e2e']['value'],d['e2e']['get_mops'],d['e2e']['h2d_ceiling']['frac'])
PY
