#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/c6_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c6_pytest.log
timeout 300 python tools/attn_bench.py --trace > gpurun_out/c6_attn_trace.json 2> gpurun_out/c6_attn_trace.err
timeout 300 python tools/attn_bench.py > gpurun_out/c6_attn_bench.json 2> gpurun_out/c6_attn_bench.err
timeout 300 python tools/bev_bench.py > gpurun_out/c6_bev_bench.json 2> gpurun_out/c6_bev_bench.err
for cfg in "1 1" "0 1" "1 0"; do set -- $cfg
  MMFN_FUSE_BN=$1 MMFN_BF16_ATTN=$2 timeout 600 python bench.py --dtype bf16 --batch 32 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c6_bench_bf16_b32_bn$1_attn$2.json 2> gpurun_out/c6_bench_bf16_b32_bn$1_attn$2.err
done
for bn in 1 0; do
  MMFN_FUSE_BN=$bn timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c6_bench_tf32_b16_bn$bn.json 2> gpurun_out/c6_bench_tf32_b16_bn$bn.err
done
tail -4 gpurun_out/c6_pytest.log; cat gpurun_out/c6_attn_trace.json; echo; head -c 600 gpurun_out/c6_bev_bench.json; echo
for f in gpurun_out/c6_bench_*.json; do echo $f; head -c 140 $f; echo; done
