#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x -k "subgraph or vectornet" > gpurun_out/c52_pytest_parity.log 2>&1
echo "rc=$?" >> gpurun_out/c52_pytest_parity.log
tail -25 gpurun_out/c52_pytest_parity.log
timeout 600 python bench.py --workload vectornet --batch 128 --no-extra > gpurun_out/c52_bench_vectornet.json 2> gpurun_out/c52_bench_vectornet.err
head -c 250 gpurun_out/c52_bench_vectornet.json; echo; tail -2 gpurun_out/c52_bench_vectornet.err
timeout 300 python tools/subgraph_bench.py > gpurun_out/c52_subgraph_bench.json 2> gpurun_out/c52_subgraph_bench.err
tail -30 gpurun_out/c52_subgraph_bench.json; tail -3 gpurun_out/c52_subgraph_bench.err
