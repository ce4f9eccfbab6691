#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "attention_fwd_bf16" > gpurun_out/c4_pytest_attn.log 2>&1
echo "rc=$?" >> gpurun_out/c4_pytest_attn.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/c4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c4_pytest.log
timeout 300 python tools/attn_bench.py > gpurun_out/c4_attn_bench.json 2> gpurun_out/c4_attn_bench.err
timeout 300 python tools/bev_bench.py > gpurun_out/c4_bev_bench.json 2> gpurun_out/c4_bev_bench.err
timeout 600 python bench.py --dtype bf16 --batch 32 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c4_bench_bf16_b32.json 2> gpurun_out/c4_bench_bf16_b32.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -c 9 -o gpurun_out/c4_attn_ncu python tools/attn_bench.py --one > gpurun_out/c4_ncu.log 2>&1
tail -4 gpurun_out/c4_pytest_attn.log; tail -4 gpurun_out/c4_pytest.log; cat gpurun_out/c4_attn_bench.json | head -c 3000; echo; cat gpurun_out/c4_bev_bench.json | head -c 600; echo; head -c 200 gpurun_out/c4_bench_bf16_b32.json
