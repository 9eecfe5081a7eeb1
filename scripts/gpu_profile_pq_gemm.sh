#!/bin/bash
# ncu capture of the tensor-core ADC filter (second pq_gemm_kernel launch = MODE 1 over all 1M codes)
set -x
CMD="python bench_aux.py --what pq --nq 1000 --cpu-queries 1 --ef 240:240:60"
$CMD > gpurun_out/pqg_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pq_gemm_kernel -s 1 -c 1 -f -o gpurun_out/prof_pq_gemm $CMD > gpurun_out/ncu_pq_gemm.log 2>&1
echo pq_gemm rc=$?
tail -2 gpurun_out/pqg_plain.log
