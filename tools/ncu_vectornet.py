#!/usr/bin/env python3
"""ncu target: the VectorNet kernels at the BASELINE configs[4] shapes (128 samples x 256 polylines x 19 vectors = 622 592 rows)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmfn_b200 import ops

dev = "cuda"
B, L, P = 128, 256, 20
G, V = B * L, P - 1
torch.manual_seed(0)
lane = torch.randn(B, L, P, 5, device=dev)
layers = []
for k in (7, 128, 128):
    layers.append((torch.randn(64, k, device=dev) * 0.2, torch.zeros(64, device=dev), torch.ones(64, device=dev), torch.zeros(64, device=dev)))
for _ in range(2):
    o = ops.subgraph_fused_fwd(lane, layers)                      # tensor-core sub-graph forward
    dy = torch.randn(G * V, 64, device=dev)
    dg, db = torch.zeros(64, device=dev), torch.zeros(64, device=dev)
    x, mean, rstd = o["y"][1], o["mean"][1], o["rstd"][1]
    ops.layernorm_bwd(dy, x, layers[1][2], layers[1][3], mean, rstd, dg, db, act=1, parts=2)      # parameter gradients (C = 64)
    dx = ops.layernorm_bwd(dy, x, layers[1][2], layers[1][3], mean, rstd, None, None, act=1, parts=1)   # dx, half-warp per row
    d2 = torch.randn(G * V, 128, device=dev)
    ops.subgraph_pool_bwd(d2, o["arg"][1], G, V)
    ops.segmax_bwd(torch.randn(G, 128, device=dev), o["argf"], G, V)
    dw = torch.zeros(64, 7, device=dev)
    ops.wgrad_n64_k7_(dy, o["vec"], dw)
    bias = torch.zeros(64, device=dev)
    ops.colsum_(dy, bias)
torch.cuda.synchronize()
