#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -x -k "stem" > gpurun_out/c35_pytest_ops.log 2>&1
echo "rc=$?" >> gpurun_out/c35_pytest_ops.log
tail -3 gpurun_out/c35_pytest_ops.log
timeout 300 python tools/stem_bench.py 32 bf16 > gpurun_out/c35_stem_bench_bf16_b32.json 2> gpurun_out/c35_stem_bench.err
grep -E "direct|fused" gpurun_out/c35_stem_bench_bf16_b32.json; tail -3 gpurun_out/c35_stem_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"stem_|bn_relu_maxpool" -s 8 -c 8 -o gpurun_out/c35_stem -f python tools/ncu_stem.py 32 > gpurun_out/c35_ncu.log 2>&1
tail -3 gpurun_out/c35_ncu.log
ncu -i gpurun_out/c35_stem.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_full_summary.py > gpurun_out/c35_ncu_stem_kernels.json
ncu -i gpurun_out/c35_stem.ncu-rep --page details --csv 2>/dev/null > gpurun_out/c35_ncu_stem_details.csv
ls -la gpurun_out/c35_stem.ncu-rep
