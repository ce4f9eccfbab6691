#!/usr/bin/env python3
"""GPU developer aid: in-graph cost of ONE launch of each hot-path kernel at its production shape.

Each op is captured N times back to back (a dependent chain on one stream) into a CUDA graph; the replay
time / N is what the kernel costs on the step's critical path (launch ramp + CTA life + memory flush), which
is what matters at B=16 where most kernels live for only a few microseconds.  ncu's per-launch durations
are cold-cache and carry ~3 us of replay overhead, so they cannot rank kernels this small."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmfn_b200 import ops                                 # noqa: E402

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
only = sys.argv[2] if len(sys.argv) > 2 else ""
N = 40
results = {}


def chain(name, fn, flops=0.0, nbytes=0.0):
    if only and only not in name:
        return
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(N):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / reps / N
    extra = ""
    if flops:
        extra += f"  {flops / us / 1e6:7.1f} TFLOP/s"
    if nbytes:
        extra += f"  {nbytes / us / 1e3:7.1f} GB/s"
    results[name] = round(us, 2)
    print(f"{name:58s} {us:7.2f} us{extra}", flush=True)


def R(*shape):
    return torch.randn(*shape, device=dev)


def gpt_ops(T, C, nh=4):
    M = B * T
    hs = C // nh
    tag = f"T{T} C{C}"
    x, g, b = R(M, C), torch.ones(C, device=dev), torch.zeros(C, device=dev)
    chain(f"{tag} ln_fwd", lambda: ops.layernorm_fwd(x, g, b), nbytes=8.0 * M * C)
    _, mean, rstd = ops.layernorm_fwd(x, g, b)
    dy, dg, db = R(M, C), torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    chain(f"{tag} ln_bwd dx(+drop copy)", lambda: ops.layernorm_bwd(dy, x, g, b, mean, rstd, None, None, dres=x, parts=1, drop=(0.1, 3)), nbytes=20.0 * M * C)
    chain(f"{tag} ln_bwd params", lambda: ops.layernorm_bwd(dy, x, g, b, mean, rstd, dg, db, parts=2), nbytes=8.0 * M * C)
    wqkv, bqkv, qkv = R(3 * C, C) * 0.02, torch.zeros(3 * C, device=dev), torch.empty(M, 3 * C, device=dev)
    chain(f"{tag} gemm qkv fwd", lambda: ops.gemm(x, wqkv, qkv, bias=bqkv), flops=2.0 * M * 3 * C * C)
    if ops.attention_fwd_ok(T, C, nh):
        chain(f"{tag} attention_fwd (drop .1)", lambda: ops.attention_fwd(qkv, B, T, C, nh, 0.1, 5), flops=4.0 * B * nh * T * T * hs)
    wp, y = R(C, C) * 0.02, torch.empty(M, C, device=dev)
    chain(f"{tag} gemm proj fwd (+res+drop)", lambda: ops.gemm(x, wp, y, bias=b, res=x, drop_p=0.1, seed=9), flops=2.0 * M * C * C)
    w1, b1, a = R(4 * C, C) * 0.02, torch.zeros(4 * C, device=dev), torch.empty(M, 4 * C, device=dev)
    chain(f"{tag} gemm fc1 fwd (relu)", lambda: ops.gemm(x, w1, a, bias=b1, act=1), flops=8.0 * M * C * C)
    w2 = R(C, 4 * C) * 0.02
    chain(f"{tag} gemm fc2 fwd (+res+drop)", lambda: ops.gemm(a, w2, y, bias=b, res=x, drop_p=0.1, seed=9), flops=8.0 * M * C * C)
    # backward
    da = torch.empty(M, 4 * C, device=dev)
    chain(f"{tag} gemm fc2 dgrad (mask)", lambda: ops.gemm(dy, w2.t(), da, mask=a), flops=8.0 * M * C * C)
    chain(f"{tag} gemm fc1 dgrad", lambda: ops.gemm(da, w1.t(), y), flops=8.0 * M * C * C)
    dw1 = torch.zeros(4 * C, C, device=dev)
    chain(f"{tag} gemm fc1 wgrad", lambda: ops.gemm(da.t(), x.t(), dw1, accum=1), flops=8.0 * M * C * C)
    chain(f"{tag} colsum 4C", lambda: ops.colsum_(da, b1), nbytes=4.0 * M * 4 * C)
    dqkv = torch.empty_like(qkv)
    k, q, v = (qkv[:, i * C:(i + 1) * C].view(B, T, nh, hs).permute(0, 2, 1, 3) for i in range(3))
    dk, dq, dv = (dqkv[:, i * C:(i + 1) * C].view(B, T, nh, hs).permute(0, 2, 1, 3) for i in range(3))
    dyh = y.view(B, T, nh, hs).permute(0, 2, 1, 3)
    dP = torch.empty(B, nh, T, T, device=dev)
    P = torch.softmax(R(B, nh, T, T), -1)
    chain(f"{tag} attn bwd dP = dy v^T", lambda: ops.gemm(dyh, v, dP), flops=2.0 * B * nh * T * T * hs)
    chain(f"{tag} attn bwd softmax_bwd", lambda: ops.softmax_bwd(P, dP, 1.0 / math.sqrt(hs), 0.1, 5), nbytes=12.0 * B * nh * T * T)
    chain(f"{tag} attn bwd dq = dS k", lambda: ops.gemm(dP, k.transpose(-1, -2), dq), flops=2.0 * B * nh * T * T * hs)
    chain(f"{tag} attn bwd dv = Pd^T dy", lambda: ops.gemm(P.transpose(-1, -2), dyh.transpose(-1, -2), dv), flops=2.0 * B * nh * T * T * hs)
    chain(f"{tag} gemm qkv dgrad", lambda: ops.gemm(dqkv, wqkv.t(), y), flops=6.0 * M * C * C)


def conv_ops(H, C, stride=1):
    tag = f"{H}x{H} C{C}"
    x, w = R(B, H, H, C), R(C, 3, 3, C) * 0.05
    fl = 2.0 * B * H * H * C * C * 9
    chain(f"{tag} conv3x3 fwd", lambda: ops.conv2d_fwd(x, w, 1, 1), flops=fl)
    dw = torch.zeros_like(w)
    chain(f"{tag} conv3x3 wgrad", lambda: ops.conv2d_wgrad_(x, x, dw, 1, 1), flops=fl)
    g, bt = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    chain(f"{tag} bn_train_fwd (+res+relu)", lambda: ops.bn_train_fwd(x, g, bt, rm, rv, res=x, relu=True), nbytes=16.0 * x.numel())
    y, mean, rstd = ops.bn_train_fwd(x, g, bt, rm, rv, relu=True)
    dg, dbt = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    chain(f"{tag} bn_train_bwd (+dres)", lambda: ops.bn_train_bwd(x, x, y, mean, rstd, g, dg, dbt, True), nbytes=24.0 * x.numel())
    chain(f"{tag} filter flip", lambda: ops.filter_crsk(w, flip=True), nbytes=8.0 * w.numel())


for (T, C) in [(192, 64), (192, 128), (192, 256), (256, 512)]:
    gpt_ops(T, C)
for (H, C) in [(64, 64), (32, 128), (16, 256), (8, 512)]:
    conv_ops(H, C)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(results, open("gpurun_out/chain_bench.json", "w"), indent=1)
