#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:"subgraph_|ln_bwd_|segmax_|wgrad_n64|colsum_vec" -s 7 -c 7 -o gpurun_out/c51_vn -f python tools/ncu_vectornet.py > gpurun_out/c51_ncu.log 2>&1
tail -3 gpurun_out/c51_ncu.log
ncu -i gpurun_out/c51_vn.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_full_summary.py > gpurun_out/c51_ncu_vectornet_kernels.json
rm -f gpurun_out/c51_vn.ncu-rep
python - <<'PY'
import json
for r in json.load(open('gpurun_out/c51_ncu_vectornet_kernels.json')):
    print(r['kernel'][:50], r.get('grid'), r.get('duration_ns'), r.get('traffic_bytes'), r.get('dram_pct_of_peak'), r.get('tensor_pipe_active_pct'), r.get('registers'))
PY
