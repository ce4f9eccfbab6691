#!/bin/bash
# fused stem tail (BN -> ReLU -> maxpool): op test, parity tests, A/B bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -x -k "stem_tail or batchnorm or conv_epilogue" > gpurun_out/c33_pytest_ops.log 2>&1
echo "rc=$?" >> gpurun_out/c33_pytest_ops.log
tail -5 gpurun_out/c33_pytest_ops.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 900 -x > gpurun_out/c33_pytest_parity.log 2>&1
echo "rc=$?" >> gpurun_out/c33_pytest_parity.log
tail -5 gpurun_out/c33_pytest_parity.log
for f in 0 1; do
  MMFN_FUSE_STEM=$f timeout 600 python bench.py --no-extra --steps 20 --warmup 5 > gpurun_out/c33_bench_tf32_stem$f.json 2> gpurun_out/c33_bench_tf32_stem$f.err
  MMFN_FUSE_STEM=$f timeout 600 python bench.py --dtype bf16 --batch 32 --no-extra --steps 20 --warmup 5 > gpurun_out/c33_bench_bf16_stem$f.json 2> gpurun_out/c33_bench_bf16_stem$f.err
done
for f in tf32_stem0 tf32_stem1 bf16_stem0 bf16_stem1; do echo $f; head -c 200 gpurun_out/c33_bench_$f.json; echo; tail -2 gpurun_out/c33_bench_$f.err; done
