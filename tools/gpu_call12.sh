#!/bin/bash
# fused VectorNet sub-graph + whole-GPT forward kernels: parity tests, microbench, step ablation, A/B bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -x -k "fused_subgraph or config5" > gpurun_out/c12_pytest_vn.log 2>&1
echo "rc=$?" >> gpurun_out/c12_pytest_vn.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 900 -k "whole_gpt" > gpurun_out/c12_pytest_gpt.log 2>&1
echo "rc=$?" >> gpurun_out/c12_pytest_gpt.log
timeout 300 python tools/subgraph_bench.py > gpurun_out/c12_subgraph_bench.json 2> gpurun_out/c12_subgraph_bench.err
for f in 0 1; do
  MMFN_FUSE_GPT=$f timeout 600 python bench.py --steps 20 --warmup 5 --no-extra > gpurun_out/c12_bench_tf32_gpt$f.json 2> gpurun_out/c12_bench_tf32_gpt$f.err
  MMFN_FUSE_GPT=$f timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --dtype bf16 --batch 32 > gpurun_out/c12_bench_bf16_gpt$f.json 2> gpurun_out/c12_bench_bf16_gpt$f.err
done
tail -15 gpurun_out/c12_pytest_vn.log; tail -30 gpurun_out/c12_pytest_gpt.log; cat gpurun_out/c12_subgraph_bench.json; tail -3 gpurun_out/c12_subgraph_bench.err
for f in gpurun_out/c12_bench_*.json; do echo $f; head -c 260 $f; echo; done
tail -3 gpurun_out/c12_bench_tf32_gpt1.err
