#!/bin/bash
set -x
CMD="python scripts/probe_flat.py"
N=1000000 $CMD > gpurun_out/scan_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:flat_scan_kernel -s 3 -c 1 -f -o gpurun_out/prof_scan $CMD > gpurun_out/ncu_scan.log 2>&1
echo rc=$?
cat gpurun_out/scan_plain.log | tail -9
