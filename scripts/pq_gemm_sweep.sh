#!/bin/bash
# CTA-group size x generator groups of pq_gemm_kernel (C4 shape)
for c in 1 2; do for g in 1 2 4; do
  echo "== CTAS=$c GEN=$g"
  VDB_PQ_CTAS=$c VDB_PQ_GEN=$g timeout 300 python scripts/probe_pq.py 2>&1 | grep "nq=1000 ef=240"
done; done
