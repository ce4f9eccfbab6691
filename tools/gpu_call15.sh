#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 900 -x -k "whole_gpt" > gpurun_out/c15_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c15_pytest.log
timeout 600 python tools/gpt_bench.py > gpurun_out/c15_gpt_bench.json 2> gpurun_out/c15_gpt_bench.err
tail -5 gpurun_out/c15_pytest.log
python - <<'PY'
import json
rows=json.loads(open('gpurun_out/c15_gpt_bench.json').read().strip().splitlines()[-1])
for r in rows:
    print(r['prec'], r['B'], r['C'], 'fwd', round(r['fwd_us_fused0']), '->', round(r['fwd_us_fused1']), 'fwdbwd', round(r['fwdbwd_us_fused0']), '->', round(r['fwdbwd_us_fused1']), 'block_us', r['block_us'], r['phase_ns'])
PY
tail -3 gpurun_out/c15_gpt_bench.err
