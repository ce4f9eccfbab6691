#!/bin/bash
# 2-GPU session: on-hardware data-parallel parity test, N=2 bench line (three-bucket exchange), BEV re-bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 900 -k "two_gpu or bev" > gpurun_out/c8_pytest_dp.log 2>&1
echo "rc=$?" >> gpurun_out/c8_pytest_dp.log
timeout 300 python tools/bev_bench.py > gpurun_out/c8_bev_bench.json 2> gpurun_out/c8_bev_bench.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c8_bench_n2.json 2> gpurun_out/c8_bench_n2.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c8_bench_n1.json 2> gpurun_out/c8_bench_n1.err
tail -5 gpurun_out/c8_pytest_dp.log; head -c 700 gpurun_out/c8_bev_bench.json; echo; head -c 200 gpurun_out/c8_bench_n2.json; echo; tail -2 gpurun_out/c8_bench_n2.err; head -c 200 gpurun_out/c8_bench_n1.json
