#!/bin/bash
# final 2-GPU checkpoint: the tests that need two devices (fused NVLink exchange over CUDA IPC, batch sharding from one process), bench line at N = 2
mkdir -p gpurun_out
exec > gpurun_out/final7_2gpu.log 2>&1
nvidia-smi --query-gpu=name --format=csv
echo "== pytest (two-device tests)"; timeout 900 python -m pytest tests -m gpu -x -q -k "two or dist or shard or devices or four_step" 2>&1 | tail -4
echo "== bench 2 GPUs"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "exit $?"; tail -2 gpurun_out/bench_2gpu.err; cut -c1-300 gpurun_out/bench_2gpu.json
