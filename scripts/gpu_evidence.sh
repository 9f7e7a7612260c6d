#!/bin/bash
# Round-2 evidence in ONE gpurun call (everything lands in gpurun_out/, summarised into profiles/ by
# scripts/summarize_profiles.py + scripts/make_traffic_json.py):
#   1. per-launch duration + DRAM bytes of exactly one bench step, per precision tier
#   2. `ncu --set full` of one launch per convolution CLASS (selected by shape through
#      scripts/one_op.py, both tiers), of the attention kernel and of the GroupNorm apply
#   3. compute-sanitizer memcheck over the kernel tests
#   4. default bench + reference arm logs
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
STEP="python bench.py --steps 1 --warmup 3 --e2e-nfe 0 --no-cpu-baseline --no-gpu-eager --no-secondary --profile-ops 0 --no-graph --profiler-range"
for tier in ${TIERS:-bf16x3 bf16}; do
  # the window is the timed step itself: bench.py brackets it with cudaProfilerStart/Stop
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      --profile-from-start off --csv --log-file gpurun_out/traffic_$tier.csv $STEP --precision $tier > gpurun_out/ncu_traffic_$tier.log 2>&1
  echo "traffic $tier rc=$?"
done
if [ "${CAPTURES:-1}" = "1" ]; then
  for s in ${SHAPES_X3:-c32t c32cat c16 c8 qkv16}; do
    ONE_OP_X3=1 ONE_OP_REPS=4 timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_tc -s 5 -c 1 -f \
        -o gpurun_out/prof_x3_$s python scripts/one_op.py $s > gpurun_out/ncu_x3_$s.log 2>&1
    echo "x3 $s rc=$?"
  done
  for s in ${SHAPES_X3_GN:-g32sc g16sc g32res}; do      # fused GroupNorm kernel of the split tier (incl. shortcut ring)
    ONE_OP_X3=1 ONE_OP_REPS=4 timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_gn -s 5 -c 1 -f \
        -o gpurun_out/prof_x3_$s python scripts/one_op.py $s > gpurun_out/ncu_x3_$s.log 2>&1
    echo "x3 $s rc=$?"
  done
  for s in ${SHAPES_BF16:-g32cat g32 g16 c16 c8}; do
    ONE_OP_REPS=4 timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv -s 5 -c 1 -f \
        -o gpurun_out/prof_bf16_$s python scripts/one_op.py $s > gpurun_out/ncu_bf16_$s.log 2>&1
    echo "bf16 $s rc=$?"
  done
  # attention + GroupNorm apply + fused update out of a real step of each tier (first matching launch
  # after the warm-up window)
  for tier in ${STEP_TIERS-bf16x3 bf16}; do
    for k in attn_tc gn_apply sscs_update; do
      timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:$k -s ${KSKIP:-3} -c 1 -f \
          -o gpurun_out/prof_${tier}_$k $STEP --precision $tier > gpurun_out/ncu_${tier}_$k.log 2>&1
      echo "$tier $k rc=$?"
    done
  done
fi
if [ "${MEMCHECK:-1}" = "1" ]; then
  timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q --no-header -p no:cacheprovider \
      -m gpu tests/test_gpu_x3.py -k "conv_tc_x3 or conv_gn_fused or attention or memory_bound or split" > gpurun_out/memcheck_x3.log 2>&1
  echo "memcheck x3 rc=$? $(tail -n 1 gpurun_out/memcheck_x3.log)"
  timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q --no-header -p no:cacheprovider \
      -m gpu tests/test_gpu_kernels.py tests/test_gpu_guidance.py -k "conv_tc or conv_gn or attention_tc or groupnorm or fir or axpby or guided_forward" > gpurun_out/memcheck_kernels.log 2>&1
  echo "memcheck kernels rc=$? $(tail -n 1 gpurun_out/memcheck_kernels.log)"
fi
if [ "${BENCH:-1}" = "1" ]; then
  ( time python bench.py ) > gpurun_out/bench_default.log 2>&1
  tail -n 4 gpurun_out/bench_default.log | cut -c1-600
  ( time python bench.py --impl reference ) > gpurun_out/bench_reference.log 2>&1
  tail -n 4 gpurun_out/bench_reference.log | cut -c1-300
fi
ls -la gpurun_out | tail -n 40
