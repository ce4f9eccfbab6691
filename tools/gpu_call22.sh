#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 600 -x -k "attention_bwd_small" > gpurun_out/c22_pytest_ops.log 2>&1
echo "rc=$?" >> gpurun_out/c22_pytest_ops.log
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 --deselect tests/test_gpu_parity.py::test_two_gpu_data_parallel_step_matches_hand_summed_gradients > gpurun_out/c22_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c22_pytest.log
for f in 0 1; do
  MMFN_FUSE_ATTN_BWD_SMALL=$f timeout 600 python bench.py --steps 20 --warmup 5 --no-extra > gpurun_out/c22_bench_tf32_ab$f.json 2> gpurun_out/c22_bench_tf32_ab$f.err
done
timeout 600 python tools/gpt_bench.py > gpurun_out/c22_gpt_bench.json 2> gpurun_out/c22_gpt_bench.err
tail -4 gpurun_out/c22_pytest_ops.log; tail -6 gpurun_out/c22_pytest.log
for f in gpurun_out/c22_bench_tf32_ab*.json; do echo $f; head -c 230 $f; echo; done
python - <<'PY'
import json
rows=json.loads(open('gpurun_out/c22_gpt_bench.json').read().strip().splitlines()[-1])
for r in rows:
    print(r['prec'], r['B'], r['C'], 'fwd', round(r['fwd_us_fused0']), '->', round(r['fwd_us_fused1']), 'fwdbwd', round(r['fwdbwd_us_fused0']), '->', round(r['fwdbwd_us_fused1']), r['fwdbwd_launches_fused0'], r['fwdbwd_launches_fused1'])
PY
