#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --no-extra --steps 20 --warmup 5 > gpurun_out/c44_bench_tf32.json 2> gpurun_out/c44_bench_tf32.err
timeout 600 python bench.py --dtype bf16 --batch 32 --no-extra --steps 20 --warmup 5 > gpurun_out/c44_bench_bf16.json 2> gpurun_out/c44_bench_bf16.err
for f in tf32 bf16; do echo $f; head -c 200 gpurun_out/c44_bench_$f.json; echo; tail -2 gpurun_out/c44_bench_$f.err; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c44_launches_vectornet.csv python bench.py --workload vectornet --batch 128 --steps 2 --warmup 1 --no-graph > gpurun_out/c44_vn.log 2>&1
python tools/ncu_summary.py gpurun_out/c44_launches_vectornet.csv 40 > gpurun_out/c44_launches_vectornet_summary.txt 2>&1
head -32 gpurun_out/c44_launches_vectornet_summary.txt
