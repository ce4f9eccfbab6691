#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -x -k "polyline or segment_max or gat_vectornet" > gpurun_out/c42_pytest_ops.log 2>&1
echo "rc=$?" >> gpurun_out/c42_pytest_ops.log
tail -3 gpurun_out/c42_pytest_ops.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x -k "vectornet or subgraph or train_step" > gpurun_out/c42_pytest_parity.log 2>&1
echo "rc=$?" >> gpurun_out/c42_pytest_parity.log
tail -3 gpurun_out/c42_pytest_parity.log
timeout 600 python bench.py --workload vectornet --batch 128 --no-extra > gpurun_out/c42_bench_vectornet.json 2> gpurun_out/c42_bench_vectornet.err
head -c 250 gpurun_out/c42_bench_vectornet.json; echo; tail -2 gpurun_out/c42_bench_vectornet.err
timeout 600 python bench.py --no-extra --steps 20 --warmup 5 > gpurun_out/c42_bench_tf32.json 2> gpurun_out/c42_bench_tf32.err
head -c 200 gpurun_out/c42_bench_tf32.json; echo
