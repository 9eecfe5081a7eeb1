#!/bin/bash
set -x
CMD1="python bench_aux.py --what ivf --nq 1000 --cpu-queries 1"
$CMD1 > gpurun_out/ivf_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ivf_list_scan_kernel -s 2 -c 1 -f -o gpurun_out/prof_ivf $CMD1 > gpurun_out/ncu_ivf.log 2>&1
echo ivf rc=$?
CMD2="python bench.py --steps 1 --warmup 1 --cpu-queries 2"
$CMD2 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pair_dist_kernel -s 8 -c 1 -f -o gpurun_out/prof_rerank $CMD2 > gpurun_out/ncu_rerank.log 2>&1
echo rerank rc=$?
