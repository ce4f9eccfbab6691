#!/bin/bash
# BASELINE configs[3] as named: RGB+LiDAR only, B=64 per GPU, 4 x B200
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --workload rgb_lidar --batch 64 --no-extra --steps 10 --warmup 3 > gpurun_out/c57_bench_rgb_lidar_n4.json 2> gpurun_out/c57_bench_rgb_lidar_n4.err
head -c 400 gpurun_out/c57_bench_rgb_lidar_n4.json; echo; tail -2 gpurun_out/c57_bench_rgb_lidar_n4.err
