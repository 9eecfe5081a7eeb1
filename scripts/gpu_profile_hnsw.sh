#!/bin/bash
# ncu capture of one full-size build batch of hnsw_search_kernel (2048 new nodes, ef_construction 200, 960-d)
set -x
CMD="python bench_aux.py --what hnsw --n 200000 --nq 1000 --hnsw-ef 120:120:40"
$CMD > gpurun_out/hnsw_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 150 -c 1 -f -o gpurun_out/prof_hnsw $CMD > gpurun_out/ncu_hnsw.log 2>&1
echo hnsw rc=$?
tail -2 gpurun_out/hnsw_plain.log | cut -c1-400
