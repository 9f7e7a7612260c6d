#!/bin/bash
# single-op timings + full ncu captures of selected shapes (scripts/one_op.py)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python scripts/one_op.py ${SHAPES:-qkv16 proj16 c32 c32t c16 c8 c8cat g32 g32cat g16} 2>&1 | tail -n 20
for s in ${CAPTURE:-}; do
  ONE_OP_REPS=4 timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv -s 5 -c 1 -f \
      -o gpurun_out/prof_$s python scripts/one_op.py $s > gpurun_out/ncu_$s.log 2>&1
  echo "$s rc=$?"
done
