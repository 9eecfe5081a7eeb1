#!/bin/bash
# ncu captures per /opt/skills/guides/B200_PROFILING.md: launch list of the bench command, then one --set full
# capture of the dominant kernel (the FILTER launches of flat_gemm_kernel). Each only after the same command exited 0
# without ncu.
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --cpu-queries 2 --no-other-configs"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:flat_gemm_kernel -s 2 -c 1 -f -o gpurun_out/r02_prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
tail -2 gpurun_out/plain.log
