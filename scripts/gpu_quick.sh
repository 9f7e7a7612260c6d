#!/bin/bash
# quick iteration: selected tests + launch list + bench
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PT="python -m pytest -q --no-header -p no:cacheprovider --timeout 600 -m gpu"
if [ -n "${TESTS:-}" ]; then timeout 900 $PT $TESTS 2>&1 | tail -n ${TAILN:-6}; fi
if [ "${LAUNCHES:-1}" = "1" ]; then
  B="python bench.py --steps 1 --warmup 3 --e2e-nfe 0 --no-cpu-baseline --profile-ops 0 --no-graph ${BENCH_ARGS:-}"
  read SKIP CNT < <(python scripts/launch_window.py 2>/dev/null)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $CNT --csv \
      --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launches.log 2>&1
  echo "launch list rc=$?"
fi
if [ "${BENCH:-1}" = "1" ]; then
  timeout 1200 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 --e2e-nfe ${E2E_NFE:-0} --no-cpu-baseline ${BENCH_ARGS:-} 2>&1 | tail -n 1 | tee gpurun_out/bench_quick.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step', d['ms_per_step'], 'samples/s', d['value'], 'finite', d['finite'])
print('per_kernel_ms', d['per_kernel_ms'])
print('roofline', d['roofline'] and {k: d['roofline'][k] for k in ('achieved','frac','share_of_step')})
print('update', d['roofline_update'].get('frac'), d['roofline_update'].get('large'))
print('clocks', d['clocks'])
for k, v in d.get('conv_classes', {}).items(): print('  ', k, v)
"
fi
