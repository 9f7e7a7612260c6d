#!/bin/bash
# bf16x3 tier bring-up: kernel tests, network / sampler tests, short bench of the tier
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PT="python -m pytest -q -rA --no-header -p no:cacheprovider --timeout 900 -m gpu"
run() { local name=$1; shift; local to=$1; shift
  echo "=== $name"; timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1
  echo "rc=$? $(tail -n 1 gpurun_out/$name.log)"; }
run x3_kernels 900 $PT tests/test_gpu_x3.py -k "split or conv_tc_x3 or memory_bound or attention"
run x3_net 1500 $PT tests/test_gpu_x3.py -k "forward or sampler or trajectory or full_batch"
grep -h -E "rel-L2|PASSED|FAILED|ERROR|Error|error" gpurun_out/x3_kernels.log gpurun_out/x3_net.log | cut -c1-220 | tail -n 70
if [ "${BENCH:-1}" = "1" ]; then
  timeout 1200 python bench.py --precision bf16x3 --steps 5 --warmup 3 --e2e-nfe 0 --no-cpu-baseline 2>&1 | tail -n 1 | tee gpurun_out/bench_x3.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step', d['ms_per_step'], 'samples/s', d['value'], 'finite', d['finite'])
print('per_kernel_ms', d['per_kernel_ms'])
print('roofline', d['roofline'] and {k: d['roofline'][k] for k in ('achieved','frac','share_of_step')})
for k, v in d.get('conv_classes', {}).items(): print('  ', k, v)
"
fi
