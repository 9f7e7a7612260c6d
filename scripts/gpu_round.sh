#!/bin/bash
# One gpurun call: GPU tests in isolated processes (a trapping kernel poisons its CUDA context),
# then smoke and a short bench.  Everything lands in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { # name, timeout, cmd...
  local name=$1; shift; local to=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1
  echo "rc=$? $(tail -n 1 gpurun_out/$name.log)" | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
PT="python -m pytest -q -rA --no-header -p no:cacheprovider --timeout 600 -m gpu"
run k_other 900 $PT tests/test_gpu_kernels.py -k "not conv_tc"
run k_tc 600 $PT tests/test_gpu_kernels.py -k "conv_tc"
run net_fp32 900 $PT tests/test_gpu_network.py -k "fp32 or contract or shared"
run net_bf16 900 $PT tests/test_gpu_network.py -k "bf16 or checkpoint or in_place"
run sampler 1200 $PT tests/test_gpu_sampler.py
run x3_kernels 900 $PT tests/test_gpu_x3.py -k "not (forward or sampler or trajectory or full_batch)"
run x3_net 1500 $PT tests/test_gpu_x3.py -k "forward or sampler or trajectory or full_batch"
run guidance 900 $PT tests/test_gpu_guidance.py
run smoke 600 python __graft_entry__.py smoke
if [ "${SKIP_BENCH:-0}" != "1" ]; then
  run bench 1500 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 --e2e-nfe ${E2E_NFE:-20} --cpu-steps 1 --cpu-batch 4 ${BENCH_ARGS:-}
fi
grep -h -E "passed|failed|error" gpurun_out/k_*.log gpurun_out/net_*.log gpurun_out/sampler.log gpurun_out/x3_*.log gpurun_out/guidance.log gpurun_out/smoke.log | tail -20
grep -h -E "^FAILED|^ERROR" gpurun_out/*.log | head -20
tail -c 6000 gpurun_out/bench.log
