#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gpt_bench.py > gpurun_out/c13_gpt_bench.json 2> gpurun_out/c13_gpt_bench.err
timeout 300 python tools/subgraph_bench.py > gpurun_out/c13_subgraph_bench.json 2> gpurun_out/c13_subgraph_bench.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x -k "fused_subgraph or whole_gpt" > gpurun_out/c13_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c13_pytest.log
cat gpurun_out/c13_gpt_bench.json | head -c 6000; tail -3 gpurun_out/c13_gpt_bench.err; cat gpurun_out/c13_subgraph_bench.json; tail -5 gpurun_out/c13_pytest.log
