#!/bin/bash
# full GPU suite + A/B bench lines for the whole-GPT kernels
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 -x --deselect tests/test_gpu_parity.py::test_two_gpu_data_parallel_step_matches_hand_summed_gradients > gpurun_out/c16_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c16_pytest.log
for f in 0 1; do
  MMFN_FUSE_GPT=$f timeout 600 python bench.py --steps 20 --warmup 5 --no-extra > gpurun_out/c16_bench_tf32_gpt$f.json 2> gpurun_out/c16_bench_tf32_gpt$f.err
  MMFN_FUSE_GPT=$f timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --dtype bf16 --batch 32 > gpurun_out/c16_bench_bf16_gpt$f.json 2> gpurun_out/c16_bench_bf16_gpt$f.err
done
tail -8 gpurun_out/c16_pytest.log
for f in gpurun_out/c16_bench_*.json; do echo $f; head -c 230 $f; echo; done
tail -3 gpurun_out/c16_bench_tf32_gpt1.err
