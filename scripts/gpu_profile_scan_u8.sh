#!/bin/bash
# ncu capture of the single-query streaming scan over u8 rows (1M x 960), after the same command ran clean without ncu
export CASES=1:10 DTYPE=u8
CMD="python scripts/probe_flat.py"
$CMD > gpurun_out/scan_u8_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:flat_scan_kernel -s 3 -c 1 -f -o gpurun_out/prof_scan_u8 $CMD > gpurun_out/ncu_scan_u8.log 2>&1
echo rc=$?
tail -1 gpurun_out/scan_u8_plain.log
