#!/bin/bash
timeout 1500 python scripts/fullscale_parity.py > gpurun_out/r2_fullscale_parity.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2_fullscale_parity.log | cut -c1-900
