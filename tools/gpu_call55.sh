#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "trajectory_tracks" > gpurun_out/c55_traj_$i.log 2>&1
  tail -1 gpurun_out/c55_traj_$i.log
  python - <<'PY'
import json
d = json.load(open('gpurun_out/trajectory_tf32.json'))
print("max rel dev %.4f  last-two %.4f  first %.2e" % (max(d['rel_dev']), abs(sum(d['loss_gpu'][-2:]) - sum(d['loss_oracle'][-2:])) / sum(d['loss_oracle'][-2:]), d['rel_dev'][0]))
PY
done
