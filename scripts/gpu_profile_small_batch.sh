#!/bin/bash
# ncu capture of the single-CTA filter pass of a 128-query Flat batch (HBM-bound on the 2-byte operand rows), after the same
# command ran clean without ncu
export CASES=128:10,32:10,8:10 FLATPATH=0 WARM=2 REPS=5
CMD="python scripts/probe_flat.py"
$CMD > gpurun_out/small_batch_plain.log 2>&1 && \
CASES=128:10 ncu --set full --clock-control none --import-source on -k regex:flat_gemm_kernel -s 4 -c 2 -f -o gpurun_out/r02_prof_small_batch $CMD > gpurun_out/ncu_small_batch.log 2>&1
echo rc=$?
cat gpurun_out/small_batch_plain.log
ncu -i gpurun_out/r02_prof_small_batch.ncu-rep --page raw --csv > gpurun_out/r02_small_batch_raw.csv 2>/dev/null
