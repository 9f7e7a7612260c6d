#!/bin/bash
# ncu evidence: per-launch time list of one bench step + full captures of the top kernels.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --e2e-nfe 0 --no-cpu-baseline --profile-ops 0 --no-graph"
# warm-up = 1 + 3*(launches+1) launches; capture the single timed step after it
read SKIP CNT < <(python scripts/launch_window.py 2>/dev/null)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $CNT --csv \
    --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
for k in ${KERNELS-conv_tc_kernel conv_gn_tc attn gn_apply sscs_update}; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s ${KSKIP:-40} -c ${KCNT:-2} \
      -f -o gpurun_out/prof_$k $B > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
ls -la gpurun_out
