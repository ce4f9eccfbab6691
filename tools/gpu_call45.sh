#!/bin/bash
# 2-GPU check of HEAD: NCCL data-parallel parity test + the N=2 bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 900 -x -k "two_gpu" > gpurun_out/c45_pytest_two_gpu.log 2>&1
echo "rc=$?" >> gpurun_out/c45_pytest_two_gpu.log
tail -5 gpurun_out/c45_pytest_two_gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-extra > gpurun_out/c45_bench_n2.json 2> gpurun_out/c45_bench_n2.err
head -c 300 gpurun_out/c45_bench_n2.json; echo; tail -3 gpurun_out/c45_bench_n2.err
