#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|hazard|Invalid|Uninitialized" gpurun_out/sanitize_$tool.log | head -8
done
