#!/bin/bash
set -x
CMD="python bench_aux.py --what pq --nq 64 --cpu-queries 1"
$CMD > gpurun_out/pq_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pq_adc_global_kernel -s 20 -c 1 -f -o gpurun_out/prof_pq $CMD > gpurun_out/ncu_pq.log 2>&1
echo rc=$?
tail -2 gpurun_out/pq_plain.log | cut -c1-200
