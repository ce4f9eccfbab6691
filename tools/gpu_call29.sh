#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -x -k "persistent_gemm" > gpurun_out/c29_pytest_ops.log 2>&1
echo "rc=$?" >> gpurun_out/c29_pytest_ops.log
tail -3 gpurun_out/c29_pytest_ops.log
timeout 600 python tools/gemm_persist_bench.py > gpurun_out/c29_gemm_persist.json 2> gpurun_out/c29_gemm_persist.err
head -16 gpurun_out/c29_gemm_persist.json | cut -c1-220; tail -3 gpurun_out/c29_gemm_persist.err
