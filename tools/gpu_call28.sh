#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -x -k "persistent_gemm" > gpurun_out/c28_pytest_ops.log 2>&1
echo "rc=$?" >> gpurun_out/c28_pytest_ops.log
tail -5 gpurun_out/c28_pytest_ops.log
timeout 600 python tools/gemm_persist_bench.py > gpurun_out/c28_gemm_persist.json 2> gpurun_out/c28_gemm_persist.err
head -20 gpurun_out/c28_gemm_persist.json | cut -c1-220; tail -3 gpurun_out/c28_gemm_persist.err
for f in 0 1; do
  MMFN_GEMM_PERSIST=$f timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c28_bench_persist$f.json 2> gpurun_out/c28_bench_persist$f.err
  python -c "
import json; d=json.load(open('gpurun_out/c28_bench_persist$f.json')); print('persist$f', d['value'], d['ms_per_step'], d['configs2_bf16_b32']['value'], d['configs2_bf16_b32']['ms_per_step'])"
done
