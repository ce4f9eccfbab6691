#!/usr/bin/env python3
"""Stem microbench (1 GPU): 7x7/2 convolution forward + weight gradient, im2col + GEMM vs the direct kernel, and the
BatchNorm -> ReLU -> MaxPool tail, unfused vs fused.  CUDA-event time of 20 back-to-back calls (working set > L2)."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmfn_b200 import ops

dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
bf16 = (sys.argv[2] == "bf16") if len(sys.argv) > 2 else True
ops.BF16 = bf16


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


out = {"batch": B, "bf16": bf16, "unit": "us"}
for C in (3, 2):
    x = torch.randn(B, 256, 256, C, device=dev)
    w = torch.randn(64, 7, 7, C, device=dev) * 0.1
    dw = torch.zeros_like(w)
    r = {}
    z, col, w_pad = ops.conv2d_fwd_im2col(x, w, 2, 3, None, bf16=bf16)
    r["fwd_im2col_gemm"] = timeit(lambda: ops.conv2d_fwd_im2col(x, w, 2, 3, w_pad, bf16=bf16))
    r["fwd_direct"] = timeit(lambda: ops.conv2d_stem7_fwd(x, w))
    dz32 = torch.randn(B, 128, 128, 64, device=dev)
    dz = dz32.to(torch.bfloat16) if bf16 else dz32
    r["wgrad_im2col_gemm"] = timeit(lambda: ops.conv2d_wgrad_im2col_(dz, col, dw))
    r["wgrad_direct"] = timeit(lambda: ops.conv2d_stem7_wgrad_(dz, x, dw))
    r["algorithmic_MB_fwd"] = (x.numel() * 4 + z.numel() * 4) / 1e6
    r["col_MB"] = col.numel() * col.element_size() / 1e6
    del col
    g, b = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev)
    rm, rv = torch.zeros(64, device=dev), torch.ones(64, device=dev)
    dg, db = torch.zeros(64, device=dev), torch.zeros(64, device=dev)
    dout = torch.randn(B, 64, 64, 64, device=dev)

    def unfused_fwd():
        y, mean, rstd = ops.bn_train_fwd(z, g, b, rm, rv, relu=True)
        return y, mean, rstd, ops.maxpool_fwd(y, want16=bf16)
    y, mean, rstd, (o, idx) = unfused_fwd()
    r["tail_fwd_unfused"] = timeit(unfused_fwd)
    r["tail_fwd_fused"] = timeit(lambda: ops.stem_bn_relu_maxpool_fwd(z, g, b, rm, rv, want16=bf16))
    _, idx2, mean2, rstd2 = ops.stem_bn_relu_maxpool_fwd(z, g, b, rm, rv, want16=bf16)
    r["tail_bwd_unfused"] = timeit(lambda: ops.bn_train_bwd(ops.maxpool_bwd(dout, idx, y.shape), z, y, mean, rstd, g, dg, db, out_bf16=bf16))
    r["tail_bwd_fused"] = timeit(lambda: ops.stem_bn_relu_maxpool_bwd(dout, idx2, z, mean2, rstd2, g, dg, db, out_bf16=bf16))
    out[f"C{C}"] = {k: round(v, 1) for k, v in r.items()}
print(json.dumps(out, indent=1))
