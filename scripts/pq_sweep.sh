#!/bin/bash
for q in 1 2 4; do
  VDB_ADC_NQ=$q timeout 900 python bench_aux.py --what pq > gpurun_out/pq_nq$q.log 2>&1
  echo "NQ=$q"; grep "PQ search" gpurun_out/pq_nq$q.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['ef'], round(d['qps']), d['recall@10'], d['gpu_vs_oracle_exact_id_rate'])"
done
