#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -x -k "stem or batchnorm or layernorm or conv_epilogue or column_sums or pool_token or bf16_twins or gat_vectornet" > gpurun_out/c40_pytest_ops.log 2>&1
echo "rc=$?" >> gpurun_out/c40_pytest_ops.log
tail -3 gpurun_out/c40_pytest_ops.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 900 -x > gpurun_out/c40_pytest_parity.log 2>&1
echo "rc=$?" >> gpurun_out/c40_pytest_parity.log
tail -3 gpurun_out/c40_pytest_parity.log
timeout 600 python bench.py --no-extra --steps 20 --warmup 5 > gpurun_out/c40_bench_tf32.json 2> gpurun_out/c40_bench_tf32.err
timeout 600 python bench.py --dtype bf16 --batch 32 --no-extra --steps 20 --warmup 5 > gpurun_out/c40_bench_bf16.json 2> gpurun_out/c40_bench_bf16.err
for f in tf32 bf16; do echo $f; head -c 200 gpurun_out/c40_bench_$f.json; echo; tail -2 gpurun_out/c40_bench_$f.err; done
