#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/c3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c3_pytest.log
timeout 300 python tools/bev_bench.py > gpurun_out/c3_bev_bench.json 2> gpurun_out/c3_bev_bench.err
timeout 600 python tools/precision_yardstick.py 16 > gpurun_out/c3_yardstick_b16.json 2> gpurun_out/c3_yardstick_b16.err
timeout 600 python tools/precision_yardstick.py 32 > gpurun_out/c3_yardstick_b32.json 2> gpurun_out/c3_yardstick_b32.err
timeout 600 python tools/precision_yardstick.py 2 > gpurun_out/c3_yardstick_b2.json 2> gpurun_out/c3_yardstick_b2.err
tail -4 gpurun_out/c3_pytest.log; cat gpurun_out/c3_bev_bench.json; tail -2 gpurun_out/c3_yardstick_b16.err
