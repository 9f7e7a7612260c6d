set -u
cd /root/repo 2>/dev/null || cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --e2e-nfe 20 > gpurun_out/bench_2gpu.log 2>&1
echo rc=$?; tail -n 3 gpurun_out/bench_2gpu.log | cut -c1-1500
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --cpu-batch 2 > gpurun_out/bench_2gpu_ref.log 2>&1
echo rc=$?; tail -n 2 gpurun_out/bench_2gpu_ref.log | cut -c1-600
