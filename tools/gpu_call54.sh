#!/bin/bash
# last verification of HEAD (no ncu): smoke, full GPU suite, bench lines
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c54_smoke.log 2>&1
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/c54_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/c54_pytest.log
timeout 900 python bench.py > gpurun_out/c54_bench_default.json 2> gpurun_out/c54_bench_default.err
timeout 600 python bench.py --dtype bf16 --batch 32 --no-extra --steps 20 --warmup 5 > gpurun_out/c54_bench_bf16.json 2> gpurun_out/c54_bench_bf16.err
timeout 600 python bench.py --workload vectornet --batch 128 --no-extra > gpurun_out/c54_bench_vectornet.json 2> gpurun_out/c54_bench_vectornet.err
tail -2 gpurun_out/c54_smoke.log; tail -4 gpurun_out/c54_pytest.log
for f in default bf16 vectornet; do head -c 250 gpurun_out/c54_bench_$f.json; echo; tail -1 gpurun_out/c54_bench_$f.err; done
