#!/bin/bash
for q in 2 4; do
  VDB_ADC_GQ=$q timeout 900 python bench_aux.py --what pq > gpurun_out/pq_gq$q.log 2>&1
  echo "GQ=$q"; grep "PQ search" gpurun_out/pq_gq$q.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['ef'], round(d['qps']), d['recall@10'], d['gpu_vs_oracle_exact_id_rate'])" | head -3
done
