#!/bin/bash
# DRAM traffic of every launch of one bench step (for roofline.traffic) + a default bench run.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
read SKIP CNT < <(python scripts/launch_window.py 2>/dev/null)
B="python bench.py --steps 1 --warmup 3 --e2e-nfe 0 --no-cpu-baseline --profile-ops 0 --no-graph"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -s $SKIP -c $CNT --csv --log-file gpurun_out/traffic.csv $B > gpurun_out/ncu_traffic.log 2>&1
echo "traffic rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $CNT --csv \
    --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launches rc=$?"
( time python bench.py ) > gpurun_out/bench_default.log 2>&1
tail -n 4 gpurun_out/bench_default.log | cut -c1-400
( time python bench.py --impl reference ) > gpurun_out/bench_reference.log 2>&1
tail -n 4 gpurun_out/bench_reference.log | cut -c1-300
