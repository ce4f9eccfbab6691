#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x -k "packed or trainer or loader_to_engine or inference or eval_forward or cuda_graph" > gpurun_out/c50_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c50_pytest.log
tail -12 gpurun_out/c50_pytest.log
