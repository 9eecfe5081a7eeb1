#!/bin/bash
# per-phase and per-kernel times of the sharded tensor search on one 125k-row shard (what an 8-way split leaves per GPU)
python scripts/probe_shard_phases.py > gpurun_out/shard_phases.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/shard_launches.csv python scripts/probe_shard_phases.py > gpurun_out/shard_ncu.log 2>&1
echo rc=$?
tail -4 gpurun_out/shard_phases.log
