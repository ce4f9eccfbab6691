#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c26_launches_vectornet.csv python bench.py --workload vectornet --batch 128 --steps 2 --warmup 1 --no-graph > gpurun_out/c26_vn.log 2>&1
python tools/ncu_summary.py gpurun_out/c26_launches_vectornet.csv 40 > gpurun_out/c26_launches_vectornet_summary.txt 2>&1
head -45 gpurun_out/c26_launches_vectornet_summary.txt
