#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "attention_fwd_bf16 or conv_epilogue or dropout or bev" > gpurun_out/c5_pytest_new.log 2>&1
echo "rc=$?" >> gpurun_out/c5_pytest_new.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/c5_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c5_pytest.log
timeout 300 python tools/attn_bench.py --trace > gpurun_out/c5_attn_trace.json 2> gpurun_out/c5_attn_trace.err
timeout 300 python tools/attn_bench.py > gpurun_out/c5_attn_bench.json 2> gpurun_out/c5_attn_bench.err
timeout 300 python tools/bev_bench.py > gpurun_out/c5_bev_bench.json 2> gpurun_out/c5_bev_bench.err
timeout 600 python bench.py --dtype bf16 --batch 32 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c5_bench_bf16_b32.json 2> gpurun_out/c5_bench_bf16_b32.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c5_bench_tf32_b16.json 2> gpurun_out/c5_bench_tf32_b16.err
timeout 600 python tools/ablate.py 32 bf16 > gpurun_out/c5_ablate_bf16_b32.log 2>&1; cp gpurun_out/ablate.json gpurun_out/c5_ablate_bf16_b32.json
tail -3 gpurun_out/c5_pytest_new.log; tail -4 gpurun_out/c5_pytest.log; cat gpurun_out/c5_attn_trace.json; echo; head -c 1500 gpurun_out/c5_attn_bench.json; echo; head -c 700 gpurun_out/c5_bev_bench.json; echo; head -c 150 gpurun_out/c5_bench_bf16_b32.json; echo; head -c 150 gpurun_out/c5_bench_tf32_b16.json
