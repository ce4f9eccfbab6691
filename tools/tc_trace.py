#!/usr/bin/env python3
"""GPU developer aid: phase timeline (ns, %globaltimer) of CTA 0 of one tensor-core GEMM launch."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmfn_b200 import ops
from mmfn_b200._lib import lib

dev = "cuda"
buf = torch.zeros(16, dtype=torch.int64, device=dev)
names = ["entry", "setup done", "first TMA landed", "MMAs issued", "acc visible", "stores issued", "tmem freed", "chunk0 in regs", "chunk0 retiled", "chunk0 stored"]
for (M, N, K) in [(3072, 128, 128), (3072, 64, 64), (4096, 2048, 512), (128, 128, 32)]:
    A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev); C = torch.empty(M, N, device=dev)
    for it in range(3):
        ops.gemm(A, B, C)
    torch.cuda.synchronize()
    lib().tc_set_trace(buf.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.gemm(A, B, C); e1.record()
    torch.cuda.synchronize()
    lib().tc_set_trace(0)
    t = buf.cpu().tolist()
    print(f"M{M} N{N} K{K}: event {e0.elapsed_time(e1)*1e3:.1f} us; " + ", ".join(f"{n}=+{t[i]-t[0]}" for i, n in enumerate(names)))
    # back-to-back throughput
    e0.record()
    for it in range(50):
        ops.gemm(A, B, C)
    e1.record(); torch.cuda.synchronize()
    print(f"   50 launches: {e0.elapsed_time(e1)*1e3/50:.1f} us each")

# bf16 configuration: transformer-4 MLP GEMMs at B=32 (the shapes the bench names) -- bf16 operands, bf16 result + ReLU
ops.set_precision("bf16")
bf = torch.bfloat16
for (M, N, K, act) in [(8192, 2048, 512, 1), (8192, 512, 2048, 0), (8192, 1536, 512, 0), (6144, 1024, 256, 1)]:
    A = torch.randn(M, K, device=dev).to(bf); B = (torch.randn(N, K, device=dev) * 0.05).to(bf)
    C = torch.empty(M, N, device=dev, dtype=bf); bias = torch.zeros(N, device=dev)
    for it in range(3):
        ops.gemm(A, B, C, bias=bias, act=act)
    torch.cuda.synchronize()
    lib().tc_set_trace(buf.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.gemm(A, B, C, bias=bias, act=act); e1.record()
    torch.cuda.synchronize()
    lib().tc_set_trace(0)
    t = buf.cpu().tolist()
    print(f"bf16 M{M} N{N} K{K}: event {e0.elapsed_time(e1)*1e3:.1f} us; " + ", ".join(f"{n}=+{t[i]-t[0]}" for i, n in enumerate(names)))
    e0.record()
    for it in range(50):
        ops.gemm(A, B, C, bias=bias, act=act)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 50
    print(f"   50 launches: {us:.1f} us each = {2.0 * M * N * K / us / 1e6:.0f} TFLOP/s")
ops.set_precision("tf32")
