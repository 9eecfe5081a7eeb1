#!/bin/bash
# ncu capture of one hnsw_search_kernel SEARCH launch (1000 queries, ef 120, 1M x 960 graph): the build's launches are
# skipped with cudaProfilerStart (VDB_CUPROF=1) + --profile-from-start off
set -x
CMD="python bench_aux.py --what hnsw --n ${N:-1000000} --nq 1000 --hnsw-ef 120:120:40"
$CMD > gpurun_out/hnsw_search_plain.log 2>&1 && \
VDB_CUPROF=1 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:hnsw_search_kernel -s 1 -c 1 -f \
    -o gpurun_out/r02_prof_hnsw_search $CMD > gpurun_out/ncu_hnsw_search.log 2>&1
echo hnsw rc=$?
tail -2 gpurun_out/hnsw_search_plain.log | cut -c1-400
