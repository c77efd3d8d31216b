#!/bin/bash
mkdir -p gpurun_out
./scripts/probe_gran > gpurun_out/gran_plain.txt 2>&1
cat gpurun_out/gran_plain.txt
ncu --metrics dram__sectors_read.sum,dram__sectors_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gran_ncu.csv ./scripts/probe_gran > gpurun_out/gran_ncu.txt 2>&1
echo done
