#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 600 -x -k "attention_bwd_small" > gpurun_out/c20_pytest_ops.log 2>&1
echo "rc=$?" >> gpurun_out/c20_pytest_ops.log
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -x --deselect tests/test_gpu_parity.py::test_two_gpu_data_parallel_step_matches_hand_summed_gradients > gpurun_out/c20_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c20_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --dtype bf16 --batch 32 > gpurun_out/c20_bench_bf16.json 2> gpurun_out/c20_bench_bf16.err
timeout 600 python tools/ablate.py 32 bf16 > gpurun_out/c20_ablate_bf16.json 2> gpurun_out/c20_ablate.err
tail -4 gpurun_out/c20_pytest_ops.log; tail -6 gpurun_out/c20_pytest.log; head -c 230 gpurun_out/c20_bench_bf16.json; echo; tail -c 1500 gpurun_out/c20_ablate_bf16.json; tail -3 gpurun_out/c20_ablate.err
