#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gpt_small_fwd -s 2 -c 1 -o gpurun_out/c16_gpt_bf16_c64 -f python tools/gpt_one.py bf16 0 32 > gpurun_out/c16_ncu.log 2>&1
tail -3 gpurun_out/c16_ncu.log
