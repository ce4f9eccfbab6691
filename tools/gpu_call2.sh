#!/bin/bash
# GPU session 2: bf16 op tests first (fast feedback), then the whole GPU suite, then bf16 bench lines.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "bf16" > gpurun_out/c2_pytest_bf16_ops.log 2>&1
echo "rc=$?" >> gpurun_out/c2_pytest_bf16_ops.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/c2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c2_pytest.log
timeout 600 python bench.py --dtype bf16 --batch 32 --steps 20 --warmup 5 > gpurun_out/c2_bench_bf16_b32.json 2> gpurun_out/c2_bench_bf16_b32.err
timeout 600 python bench.py --dtype tf32 --batch 32 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c2_bench_tf32_b32.json 2> gpurun_out/c2_bench_tf32_b32.err
timeout 600 python bench.py --dtype bf16 --batch 16 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c2_bench_bf16_b16.json 2> gpurun_out/c2_bench_bf16_b16.err
tail -5 gpurun_out/c2_pytest_bf16_ops.log; tail -5 gpurun_out/c2_pytest.log; head -c 300 gpurun_out/c2_bench_bf16_b32.json; tail -3 gpurun_out/c2_bench_bf16_b32.err
