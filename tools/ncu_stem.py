#!/usr/bin/env python3
"""ncu target: one launch each of the stem kernels at the configs[2] shapes (B=32, bf16 configuration)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmfn_b200 import ops

dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ops.BF16 = True
for C in (3, 2):
    x = torch.randn(B, 256, 256, C, device=dev)
    w = torch.randn(64, 7, 7, C, device=dev) * 0.1
    dw = torch.zeros_like(w)
    g, b = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev)
    rm, rv = torch.zeros(64, device=dev), torch.ones(64, device=dev)
    dg, db = torch.zeros(64, device=dev), torch.zeros(64, device=dev)
    dout = torch.randn(B, 64, 64, 64, device=dev)
    for _ in range(2):
        z = ops.conv2d_stem7_fwd(x, w)
        out, idx, mean, rstd = ops.stem_bn_relu_maxpool_fwd(z, g, b, rm, rv, want16=True)
        dz = ops.stem_bn_relu_maxpool_bwd(dout, idx, z, mean, rstd, g, dg, db, out_bf16=True)
        ops.conv2d_stem7_wgrad_(dz, x, dw)
    torch.cuda.synchronize()
