#!/bin/bash
# tuning sweeps on the step (one knob at a time against the default)
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --no-extra --steps 20 --warmup 5 > gpurun_out/c47_tf32_$name.json 2> gpurun_out/c47_tf32_$name.err
  env "$@" timeout 600 python bench.py --dtype bf16 --batch 32 --no-extra --steps 20 --warmup 5 > gpurun_out/c47_bf16_$name.json 2> gpurun_out/c47_bf16_$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
out = []
for f in ("tf32", "bf16"):
    try:
        d = json.loads(open(f"gpurun_out/c47_{f}_{n}.json").read().strip().splitlines()[-1])
        out.append(f"{f} {d['value']:.1f} ({d['ms_per_step']:.3f} ms)")
    except Exception as e:
        out.append(f"{f} FAILED {e}")
print(n, " | ".join(out))
PY
}
run default MMFN_DUMMY=1
run bnfuse256 MMFN_FUSE_BN_MAX_CTAS=256
run bnfuse2048 MMFN_FUSE_BN_MAX_CTAS=2048
run gpt2 MMFN_FUSE_GPT=2
run leaf2 MMFN_AUX_LEAF=2
run leaf8 MMFN_AUX_LEAF=8
run persist222 MMFN_GEMM_PERSIST_MIN=222
run persist148 MMFN_GEMM_PERSIST_MIN=148
