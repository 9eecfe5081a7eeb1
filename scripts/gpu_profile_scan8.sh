#!/bin/bash
# ncu capture of the 8-query variant of the streaming scan (K1), after the same command ran clean without ncu
set -x
export CASES=8:10
CMD="python scripts/probe_flat.py"
$CMD > gpurun_out/scan8_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:flat_scan_kernelILi8 -s 3 -c 1 -f -o gpurun_out/prof_scan8 $CMD > gpurun_out/ncu_scan8.log 2>&1
echo rc=$?
tail -3 gpurun_out/scan8_plain.log
