#!/bin/bash
# final verification of HEAD: full GPU suite, smoke, ncu --set full table (incl. the persistent GEMM), launch lists, bench lines
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c49_smoke.log 2>&1
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 --deselect tests/test_gpu_parity.py::test_two_gpu_data_parallel_step_matches_hand_summed_gradients > gpurun_out/c49_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c49_pytest.log
timeout 900 ncu --set full --clock-control none -k regex:"tc_kernel|tc_gemm_persist|conv3x3|attn_fwd|attn_bwd_small|bn_|bev_|adamw|im2col|softmax|stem_|ln_bwd_param" -c 200 -o gpurun_out/c49_full -f python tools/ncu_targets.py 1 bf16 > gpurun_out/c49_ncu_full.log 2>&1
ncu -i gpurun_out/c49_full.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_full_summary.py > gpurun_out/c49_ncu_full_kernels.json
rm -f gpurun_out/c49_full.ncu-rep
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c49_launches_tf32_b16.csv python tools/graph_step_launches.py 16 tf32 > gpurun_out/c49_l1.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c49_launches_bf16_b32.csv python tools/graph_step_launches.py 32 bf16 > gpurun_out/c49_l2.log 2>&1
python tools/ncu_summary.py gpurun_out/c49_launches_tf32_b16.csv 45 > gpurun_out/c49_launches_tf32_b16_summary.txt 2>&1
python tools/ncu_summary.py gpurun_out/c49_launches_bf16_b32.csv 45 > gpurun_out/c49_launches_bf16_b32_summary.txt 2>&1
timeout 900 python bench.py > gpurun_out/c49_bench_default.json 2> gpurun_out/c49_bench_default.err
timeout 600 python bench.py --dtype bf16 --batch 32 --no-extra --steps 20 --warmup 5 > gpurun_out/c49_bench_bf16.json 2> gpurun_out/c49_bench_bf16.err
timeout 600 python bench.py --workload rgb_lidar --batch 64 --no-extra > gpurun_out/c49_bench_rgb_lidar.json 2> gpurun_out/c49_bench_rgb_lidar.err
tail -2 gpurun_out/c49_smoke.log; tail -4 gpurun_out/c49_pytest.log; head -10 gpurun_out/c49_launches_bf16_b32_summary.txt
for f in default bf16 rgb_lidar; do head -c 250 gpurun_out/c49_bench_$f.json; echo; done
python - <<'PY'
import json
for r in json.load(open('gpurun_out/c49_ncu_full_kernels.json')):
    if 'persist' in r['kernel'] or 'attn_bwd_small' in r['kernel']: print(r['kernel'][:90], r.get('grid'), r.get('duration_ns'), r.get('traffic_bytes'), r.get('tensor_pipe_active_pct'))
PY
timeout 600 python bench.py --workload vectornet --batch 128 --no-extra > gpurun_out/c49_bench_vectornet.json 2> gpurun_out/c49_bench_vectornet.err
head -c 250 gpurun_out/c49_bench_vectornet.json; echo
