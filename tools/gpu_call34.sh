#!/bin/bash
# direct stem convolution: op tests, microbench, parity suite, A/B bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -x -k "stem" > gpurun_out/c34_pytest_ops.log 2>&1
echo "rc=$?" >> gpurun_out/c34_pytest_ops.log
tail -15 gpurun_out/c34_pytest_ops.log
timeout 300 python tools/stem_bench.py 32 bf16 > gpurun_out/c34_stem_bench_bf16_b32.json 2> gpurun_out/c34_stem_bench.err
timeout 300 python tools/stem_bench.py 16 tf32 > gpurun_out/c34_stem_bench_tf32_b16.json 2>> gpurun_out/c34_stem_bench.err
cat gpurun_out/c34_stem_bench_bf16_b32.json; tail -3 gpurun_out/c34_stem_bench.err
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 900 -x > gpurun_out/c34_pytest_parity.log 2>&1
echo "rc=$?" >> gpurun_out/c34_pytest_parity.log
tail -5 gpurun_out/c34_pytest_parity.log
for f in 0 1; do
  MMFN_DIRECT_STEM=$f timeout 600 python bench.py --no-extra --steps 20 --warmup 5 > gpurun_out/c34_bench_tf32_direct$f.json 2> gpurun_out/c34_bench_tf32_direct$f.err
  MMFN_DIRECT_STEM=$f timeout 600 python bench.py --dtype bf16 --batch 32 --no-extra --steps 20 --warmup 5 > gpurun_out/c34_bench_bf16_direct$f.json 2> gpurun_out/c34_bench_bf16_direct$f.err
done
for f in tf32_direct0 tf32_direct1 bf16_direct0 bf16_direct1; do echo $f; head -c 200 gpurun_out/c34_bench_$f.json; echo; tail -2 gpurun_out/c34_bench_$f.err; done
