#!/bin/bash
# multi-GPU bench lines for BASELINE configs[2] (CIFAR-10, B_total = 2048, strong scaling) and
# configs[3] (CelebA-64, B_total = 512) and configs[4] (WORKLOADS=cfg: classifier-free guidance, B_total = 1024) on N GPUs of one box:  N=<n> bash scripts/gpu_scale.sh
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-2}
COMMON="--steps ${STEPS:-5} --warmup 3 --e2e-nfe ${E2E_NFE:-50} --no-cpu-baseline --no-gpu-eager ${EXTRA:-}"
run() { # tag, args...
  local tag=$1; shift
  if [ "$N" = "1" ]; then
    timeout ${TO:-900} python bench.py --gpus 1 $COMMON "$@" > gpurun_out/scale_${tag}_n$N.log 2>&1
  else
    timeout ${TO:-900} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
        --master-port 29517 bench.py --gpus $N $COMMON "$@" > gpurun_out/scale_${tag}_n$N.log 2>&1
  fi
  echo "$tag N=$N rc=$?"
  tail -n 1 gpurun_out/scale_${tag}_n$N.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    o=d.get('bf16_path') or {}
    print('  ', d['dtype'], 'value', round(d['value'],3), 'e2e', d['e2e'] and round(d['e2e']['value'],3), 'ms/step', round(d['ms_per_step'],2), d['scaling'], d['config']['batch_total'], '| bf16', o.get('value') and round(o['value'],3), o.get('e2e') and round(o['e2e']['value'],3))
except Exception as e:
    print('  (no JSON line)', e)
"
}
for w in ${WORKLOADS:-cifar celeba}; do
  if [ "$w" = "cifar" ]; then run cifar2048 --batch-total 2048; fi
  if [ "$w" = "celeba" ]; then run celeba512 --workload celeba64 --batch-total 512; fi
  if [ "$w" = "cfg" ]; then run cfg1024 --workload cifar10_cfg --batch-total 1024; fi
done
