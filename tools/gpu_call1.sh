#!/bin/bash
# GPU session 1 (round 2): probes, the whole GPU test-suite with the new parity tests, TF32 bench line, the
# torch-eager competitor, configs[3]/[4] lines.  Everything lands under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt 2>&1
timeout 60 ./tools/exp/bf16_desc_probe > gpurun_out/c1_bf16_desc_probe.txt 2>&1
timeout 60 ./tools/exp/bf16_gemm_probe > gpurun_out/c1_bf16_gemm_probe.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/c1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c1_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c1_bench_tf32_b16.json 2> gpurun_out/c1_bench_tf32_b16.err
for cfg in "tf32 16" "bf16 32" "tf32 32" "bf16 16"; do set -- $cfg
  timeout 400 python bench.py --impl torch-eager --dtype $1 --batch $2 --steps 10 --warmup 3 > gpurun_out/c1_eager_$1_b$2.json 2> gpurun_out/c1_eager_$1_b$2.err
done
timeout 400 python bench.py --workload rgb_lidar --steps 10 --warmup 3 > gpurun_out/c1_bench_rgb_lidar_b64.json 2> gpurun_out/c1_bench_rgb_lidar_b64.err
timeout 400 python bench.py --workload vectornet --steps 20 --warmup 3 > gpurun_out/c1_bench_vectornet_b128.json 2> gpurun_out/c1_bench_vectornet_b128.err
timeout 300 python bench.py --impl torch-eager --workload rgb_lidar --batch 64 --steps 10 > gpurun_out/c1_eager_rgb_lidar_b64.json 2> gpurun_out/c1_eager_rgb_lidar_b64.err
tail -3 gpurun_out/c1_pytest.log; head -c 600 gpurun_out/c1_bench_tf32_b16.json; echo; cat gpurun_out/c1_bf16_desc_probe.txt
