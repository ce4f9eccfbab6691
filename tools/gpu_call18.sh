#!/bin/bash
# final profiling session: ncu --set full of the headline kernels (incl. the whole-GPT and sub-graph kernels), launch lists of
# the final graph step in both configurations, smoke(), final bench lines
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c18_smoke.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"tc_kernel|conv3x3|attn_fwd|bn_|bev_|adamw|im2col|softmax" -c 160 -o gpurun_out/c18_full -f python tools/ncu_targets.py 1 bf16 > gpurun_out/c18_ncu_full.log 2>&1
ncu -i gpurun_out/c18_full.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_full_summary.py > gpurun_out/c18_ncu_full_kernels.json
rm -f gpurun_out/c18_full.ncu-rep
for cfg in "bf16 0 32" "tf32 0 16" "bf16 1 32"; do
  set -- $cfg
  timeout 600 ncu --set full --clock-control none -k regex:"gpt_small" -s 3 -c 3 -o gpurun_out/c18_gpt_$1_$2 -f python tools/gpt_one.py $1 $2 $3 > gpurun_out/c18_ncu_gpt.log 2>&1
  ncu -i gpurun_out/c18_gpt_$1_$2.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_full_summary.py > gpurun_out/c18_ncu_gpt_$1_site$2.json
  rm -f gpurun_out/c18_gpt_$1_$2.ncu-rep
done
timeout 600 ncu --set full --clock-control none -k regex:"subgraph_fused" -s 6 -c 3 -o gpurun_out/c18_sub -f python tools/subgraph_bench.py > gpurun_out/c18_ncu_sub.log 2>&1
ncu -i gpurun_out/c18_sub.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_full_summary.py > gpurun_out/c18_ncu_subgraph.json
rm -f gpurun_out/c18_sub.ncu-rep
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c18_launches_tf32_b16.csv python tools/graph_step_launches.py 16 tf32 > gpurun_out/c18_l1.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c18_launches_bf16_b32.csv python tools/graph_step_launches.py 32 bf16 > gpurun_out/c18_l2.log 2>&1
python tools/ncu_summary.py gpurun_out/c18_launches_tf32_b16.csv 45 > gpurun_out/c18_launches_tf32_b16_summary.txt 2>&1
python tools/ncu_summary.py gpurun_out/c18_launches_bf16_b32.csv 45 > gpurun_out/c18_launches_bf16_b32_summary.txt 2>&1
timeout 900 python bench.py > gpurun_out/c18_bench_default.json 2> gpurun_out/c18_bench_default.err
timeout 600 python bench.py --workload vectornet --batch 128 > gpurun_out/c18_bench_vectornet.json 2> gpurun_out/c18_bench_vectornet.err
timeout 600 python bench.py --workload rgb_lidar --batch 64 --no-extra > gpurun_out/c18_bench_rgb_lidar.json 2> gpurun_out/c18_bench_rgb_lidar.err
cat gpurun_out/c18_smoke.log | tail -2; head -8 gpurun_out/c18_launches_tf32_b16_summary.txt; head -c 250 gpurun_out/c18_bench_default.json; echo; head -c 250 gpurun_out/c18_bench_vectornet.json; echo; head -c 250 gpurun_out/c18_bench_rgb_lidar.json; echo
python - <<'PY'
import json
for f in ('c18_ncu_gpt_bf16_site0','c18_ncu_gpt_tf32_site0','c18_ncu_gpt_bf16_site1','c18_ncu_subgraph'):
    try:
        for r in json.load(open(f'gpurun_out/{f}.json')): print(f, r['kernel'][:60], r.get('grid'), r.get('duration_ns'), r.get('traffic_bytes'), r.get('tensor_pipe_active_pct'))
    except Exception as e: print(f, 'ERR', e)
PY
