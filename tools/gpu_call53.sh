#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -x -k "attention_fwd_fused or attention_products or block_gradients or dropout_masks" > gpurun_out/c53_pytest_ops.log 2>&1
echo "rc=$?" >> gpurun_out/c53_pytest_ops.log
tail -12 gpurun_out/c53_pytest_ops.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 900 -x -k "train_step or whole_gpt or config2 or trajectory or cuda_graph or transfuser" > gpurun_out/c53_pytest_parity.log 2>&1
echo "rc=$?" >> gpurun_out/c53_pytest_parity.log
tail -6 gpurun_out/c53_pytest_parity.log
for f in 0 1; do
  MMFN_ATTN_FWD_SMALL=$f timeout 600 python bench.py --no-extra --steps 20 --warmup 5 > gpurun_out/c53_bench_tf32_small$f.json 2> gpurun_out/c53_bench_tf32_small$f.err
  echo small$f; head -c 200 gpurun_out/c53_bench_tf32_small$f.json | cut -c40-200; echo; tail -1 gpurun_out/c53_bench_tf32_small$f.err
done
