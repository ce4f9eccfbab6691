#!/bin/bash
# N-GPU session: BASELINE configs[3] (RGB+LiDAR only, B=64/GPU) on 4 GPUs, or the default line on all GPUs
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
if [ "$N" -ge 8 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/c17_bench_n8.json 2> gpurun_out/c17_bench_n8.err
  head -c 300 gpurun_out/c17_bench_n8.json; echo; tail -2 gpurun_out/c17_bench_n8.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 4 --steps 20 --warmup 5 --workload rgb_lidar --batch 64 > gpurun_out/c17_bench_rgb_lidar_n4.json 2> gpurun_out/c17_bench_rgb_lidar_n4.err
  head -c 300 gpurun_out/c17_bench_rgb_lidar_n4.json; echo; tail -2 gpurun_out/c17_bench_rgb_lidar_n4.err
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/c17_bench_n4.json 2> gpurun_out/c17_bench_n4.err
  head -c 300 gpurun_out/c17_bench_n4.json; echo; tail -2 gpurun_out/c17_bench_n4.err
fi
