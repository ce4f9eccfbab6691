#!/bin/bash
# last verification of HEAD: smoke + full GPU suite
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c56_smoke.log 2>&1
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/c56_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c56_pytest.log
tail -2 gpurun_out/c56_smoke.log; tail -4 gpurun_out/c56_pytest.log
python - <<'PY'
import json
d = json.load(open('gpurun_out/trajectory_tf32.json'))
print("trajectory: max rel dev %.4f  last-two %.4f" % (max(d['rel_dev']), abs(sum(d['loss_gpu'][-2:]) - sum(d['loss_oracle'][-2:])) / sum(d['loss_oracle'][-2:])))
PY
