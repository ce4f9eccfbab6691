#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "bev or loader" > gpurun_out/c11_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c11_pytest.log
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 300 -k "stem or epilogue" >> gpurun_out/c11_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c11_pytest.log
for v in 0 1; do MMFN_BEV_DEDUPE=$v timeout 200 python tools/bev_bench.py > gpurun_out/c11_bev_dedupe$v.json 2>> gpurun_out/c11_bev.err; done
python - <<'PY'
import json
for v in (0,1):
    d=json.loads(open(f'gpurun_out/c11_bev_dedupe{v}.json').read().strip().splitlines()[-1])
    print('dedupe',v,[(r['frames'], round(r['warm_us'],1), round(r['warm_frac'],3), round(r['cold_us'],1)) for r in d])
PY
tail -6 gpurun_out/c11_pytest.log
