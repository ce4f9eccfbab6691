#!/bin/bash
mkdir -p gpurun_out
for r in 1024 512; do
  MMFN_BN_SMALL_ROWS=$r timeout 600 python bench.py --no-extra --steps 20 --warmup 5 > gpurun_out/c46_bench_tf32_bn$r.json 2> gpurun_out/c46_bench_tf32_bn$r.err
  MMFN_BN_SMALL_ROWS=$r timeout 600 python bench.py --dtype bf16 --batch 32 --no-extra --steps 20 --warmup 5 > gpurun_out/c46_bench_bf16_bn$r.json 2> gpurun_out/c46_bench_bf16_bn$r.err
  for f in tf32 bf16; do echo $f $r; head -c 190 gpurun_out/c46_bench_${f}_bn$r.json | cut -c50-190; echo; tail -1 gpurun_out/c46_bench_${f}_bn$r.err; done
done
