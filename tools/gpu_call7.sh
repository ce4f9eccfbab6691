#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "bev" > gpurun_out/c7_pytest_bev.log 2>&1
echo "rc=$?" >> gpurun_out/c7_pytest_bev.log
timeout 300 python tools/bev_bench.py > gpurun_out/c7_bev_bench.json 2> gpurun_out/c7_bev_bench.err
timeout 900 ncu --set full --clock-control none -k regex:"tc_kernel|conv3x3|attn_fwd|bn_|bev_|adamw|im2col|softmax" -c 150 -o gpurun_out/c7_full python tools/ncu_targets.py 1 bf16 > gpurun_out/c7_ncu_full.log 2>&1
ncu -i gpurun_out/c7_full.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_full_summary.py > gpurun_out/c7_ncu_full_kernels.json
rm -f gpurun_out/c7_full.ncu-rep
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c7_launches_tf32_b16.csv python tools/graph_step_launches.py 16 tf32 > gpurun_out/c7_l1.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c7_launches_bf16_b32.csv python tools/graph_step_launches.py 32 bf16 > gpurun_out/c7_l2.log 2>&1
python tools/ncu_summary.py gpurun_out/c7_launches_tf32_b16.csv 45 > gpurun_out/c7_launches_tf32_b16_summary.txt 2>&1
python tools/ncu_summary.py gpurun_out/c7_launches_bf16_b32.csv 45 > gpurun_out/c7_launches_bf16_b32_summary.txt 2>&1
tail -3 gpurun_out/c7_pytest_bev.log; head -c 500 gpurun_out/c7_bev_bench.json; echo; head -12 gpurun_out/c7_launches_bf16_b32_summary.txt; tail -3 gpurun_out/c7_ncu_full.log
