#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/c9_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c9_pytest.log
timeout 300 python tools/bev_bench.py > gpurun_out/c9_bev_bench_onevisit.json 2> gpurun_out/c9_bev_bench.err
MMFN_BEV_STRIPS=1 timeout 300 python tools/bev_bench.py > gpurun_out/c9_bev_bench_strip.json 2>> gpurun_out/c9_bev_bench.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/c9_bench_n1.json 2> gpurun_out/c9_bench_n1.err
tail -4 gpurun_out/c9_pytest.log; head -c 600 gpurun_out/c9_bev_bench_onevisit.json; echo; head -c 600 gpurun_out/c9_bev_bench_strip.json; echo; head -c 200 gpurun_out/c9_bench_n1.json; echo; python -c "
import json; d=json.loads(open('gpurun_out/c9_bench_n1.json').read().strip().splitlines()[-1]); print(d.get('configs2_bf16_b32'))"
