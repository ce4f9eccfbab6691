#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2 3 4 7; do
  MMFN_BEV_DEBUG=$v timeout 200 python tools/bev_bench.py > gpurun_out/c10_bev_dbg$v.json 2>> gpurun_out/c10_bev.err
done
python - <<'PY'
import json
for v in (0,1,2,3,4,7):
    try:
        d=json.loads(open(f'gpurun_out/c10_bev_dbg{v}.json').read().strip().splitlines()[-1])
        print('dbg',v,[(r['frames'], round(r['warm_us'],1)) for r in d])
    except Exception as e: print(v,'ERR',e)
PY
