#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 600 -x -k "attention_bwd_small" > gpurun_out/c19_pytest_ops.log 2>&1
echo "rc=$?" >> gpurun_out/c19_pytest_ops.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 900 -x -k "whole_gpt or bf16" > gpurun_out/c19_pytest_parity.log 2>&1
echo "rc=$?" >> gpurun_out/c19_pytest_parity.log
timeout 600 python tools/gpt_bench.py > gpurun_out/c19_gpt_bench.json 2> gpurun_out/c19_gpt_bench.err
for f in 0 1; do
  MMFN_FUSE_ATTN_BWD_SMALL=$f timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --dtype bf16 --batch 32 > gpurun_out/c19_bench_bf16_ab$f.json 2> gpurun_out/c19_bench_bf16_ab$f.err
done
for cfg in "bf16 0 32" "tf32 0 16" "bf16 1 32"; do
  set -- $cfg
  timeout 600 ncu --set full --clock-control none -k regex:"gpt_small_fwd" -s 1 -c 1 -o gpurun_out/c19_gpt_$1_$2 -f python tools/gpt_one.py $1 $2 $3 > gpurun_out/c19_ncu_gpt.log 2>&1
  ncu -i gpurun_out/c19_gpt_$1_$2.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_full_summary.py > gpurun_out/c19_ncu_gpt_$1_site$2.json
  rm -f gpurun_out/c19_gpt_$1_$2.ncu-rep
done
tail -4 gpurun_out/c19_pytest_ops.log; tail -6 gpurun_out/c19_pytest_parity.log
python - <<'PY'
import json
rows=json.loads(open('gpurun_out/c19_gpt_bench.json').read().strip().splitlines()[-1])
for r in rows:
    print(r['prec'], r['B'], r['C'], 'fwd', round(r['fwd_us_fused0']), '->', round(r['fwd_us_fused1']), 'fwdbwd', round(r['fwdbwd_us_fused0']), '->', round(r['fwdbwd_us_fused1']), r['fwdbwd_launches_fused0'], r['fwdbwd_launches_fused1'])
for f in ('c19_ncu_gpt_bf16_site0','c19_ncu_gpt_tf32_site0','c19_ncu_gpt_bf16_site1'):
    try:
        for r in json.load(open(f'gpurun_out/{f}.json')): print(f, r['kernel'][:60], r.get('grid'), r.get('duration_ns'), r.get('traffic_bytes'), r.get('tensor_pipe_active_pct'), r.get('registers'))
    except Exception as e: print(f, 'ERR', e)
PY
for f in gpurun_out/c19_bench_bf16_ab*.json; do echo $f; head -c 230 $f; echo; done
