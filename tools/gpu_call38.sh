#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x -k "phase1 or packed or loader_to_engine or bev" > gpurun_out/c38_pytest_loader.log 2>&1
echo "rc=$?" >> gpurun_out/c38_pytest_loader.log
tail -30 gpurun_out/c38_pytest_loader.log
