#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 900 -x > gpurun_out/c48_pytest_parity.log 2>&1
echo "rc=$?" >> gpurun_out/c48_pytest_parity.log
tail -4 gpurun_out/c48_pytest_parity.log
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --no-extra --steps 20 --warmup 5 > gpurun_out/c48_tf32_$name.json 2> gpurun_out/c48_tf32_$name.err
  env "$@" timeout 600 python bench.py --dtype bf16 --batch 32 --no-extra --steps 20 --warmup 5 > gpurun_out/c48_bf16_$name.json 2> gpurun_out/c48_bf16_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
out = []
for f in ("tf32", "bf16"):
    try:
        d = json.loads(open(f"gpurun_out/c48_{f}_{n}.json").read().strip().splitlines()[-1])
        out.append(f"{f} {d['value']:.1f} ({d['ms_per_step']:.3f} ms)")
    except Exception as e:
        out.append(f"{f} FAILED {e}")
print(n, " | ".join(out))
PY
}
run default MMFN_DUMMY=1
run inner0 MMFN_BF16_ONLY_INNER=0
