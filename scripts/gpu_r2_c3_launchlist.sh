#!/bin/bash
# per-launch times + DRAM / L2-atomic traffic of the Zipf/CF build (round-1 build, 256 M ops)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,launch__grid_size,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none -c 200 --csv --log-file gpurun_out/r2_c3_baseline_ncu.csv \
  python scripts/cf_scale.py 4000000 4000000 > gpurun_out/r2_c3_baseline_ncu.txt 2>&1
tail -5 gpurun_out/r2_c3_baseline_ncu.txt
