#!/usr/bin/env python3
"""GPU: the persistent GEMM kernel (one CTA per SM, TMA ring across tiles, double-buffered TMEM accumulator, two epilogue
groups) against the one-tile-per-CTA kernel on the large GEMMs of transformers 3 / 4 -- in-graph time per launch (chain of
20 dependent launches, 10 replays), both operand precisions."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmfn_b200 import ops  # noqa: E402
from mmfn_b200._lib import lib  # noqa: E402

dev = "cuda"


def chain_us(fn, n=20, reps=10):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(); fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / reps / n


out = []
for prec, B in (("bf16", 32), ("tf32", 16)):
    ops.set_precision(prec)
    dt = torch.bfloat16 if prec == "bf16" else torch.float32
    for (T, C) in ((256, 512), (192, 256)):
        M = B * T
        shapes = [("qkv", M, 3 * C, C, {}), ("fc1 relu", M, 4 * C, C, dict(act=1)), ("fc2 res drop", M, C, 4 * C, dict(res=True, drop_p=0.1, seed=3)),
                  ("fc2 dgrad mask", M, 4 * C, C, dict(mask=True, tr=True))]
        for name, m, n, k, opt in shapes:
            A = torch.randn(m, k, device=dev).to(dt)
            W = (torch.randn(k, n, device=dev) * 0.05).to(dt).t() if opt.get("tr") else (torch.randn(n, k, device=dev) * 0.05).to(dt)
            Cout = torch.empty(m, n, device=dev, dtype=dt if prec == "bf16" and not opt.get("res") else torch.float32)
            kw = dict(bias=torch.zeros(n, device=dev))
            if opt.get("act"):
                kw["act"] = 1
            if opt.get("res"):
                kw.update(res=torch.randn(m, n, device=dev), drop_p=opt["drop_p"], seed=opt["seed"])
            if opt.get("mask"):
                kw = dict(mask=torch.randn(m, n, device=dev).to(dt))
            r = dict(prec=prec, shape=f"{name} {m}x{n}x{k}", tiles=((m + 127) // 128) * ((n + 127) // 128))
            for persist in (0, 1):
                lib().set_gemm_persist(persist)
                us = chain_us(lambda: ops.gemm(A, W, Cout, **kw))
                r[f"us_persist{persist}"] = round(us, 2)
                r[f"tflops_persist{persist}"] = round(2.0 * m * n * k / us / 1e6, 1)
            out.append(r)
            print(json.dumps(r), flush=True)
lib().set_gemm_persist(1)
# per-tile timeline of CTA 0 (bf16 fc1 of transformer 4): ns relative to the first stamp
ops.set_precision("bf16")
m, n, k = 8192, 2048, 512
A = torch.randn(m, k, device=dev).to(torch.bfloat16); W = (torch.randn(n, k, device=dev) * 0.05).to(torch.bfloat16)
Cout = torch.empty(m, n, device=dev, dtype=torch.bfloat16); bias = torch.zeros(n, device=dev)
for _ in range(3):
    ops.gemm(A, W, Cout, bias=bias, act=1)
tr = torch.zeros(16 * 8, dtype=torch.int64, device=dev)
lib().tc_set_persist_trace(tr.data_ptr())
ops.gemm(A, W, Cout, bias=bias, act=1)
torch.cuda.synchronize()
lib().tc_set_persist_trace(0)
t = tr.view(16, 8).cpu()
t0 = int(t[0, 0])
names = ["prod first", "prod last", "mma has acc", "mma first operands", "mma committed", "epi sees acc", "epi stores issued"]
for it in range(8):
    print(f"tile {it}: " + ", ".join(f"{nm}={int(t[it, i]) - t0}" for i, nm in enumerate(names)), flush=True)
ops.set_precision("tf32")
print(json.dumps(out))
