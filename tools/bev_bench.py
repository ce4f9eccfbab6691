#!/usr/bin/env python3
"""GPU: BEV scatter microbench (north_star: achieved HBM GB/s for the scatter).  Algorithmic bytes per frame =
32768 x 16 B (XYZI rows) + 2 x 256 x 256 x 4 B (fp32 grid) = 1 048 576 B.  CUDA-graph chains, L2 flushed
variant included (a 512 MB memset between replays) since 16 frames (16.8 MB) sit in the 126 MB L2 otherwise."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmfn_b200 import ops, synthetic  # noqa: E402

PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0


def chain_us(fn, chain=20, replays=10):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(chain):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / (replays * chain)


def cold_us(fn, reps=10):
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(reps):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(1e3 * e0.elapsed_time(e1))
    return float(np.median(ts))


out = []
for frames in (16, 32, 64, 128):
    pts = torch.from_numpy(np.stack([synthetic.synth_points(9000 + i) for i in range(frames)])).cuda()
    # MMFN_BEV_STRIPS=1: the shared-memory strip kernel with its automatic strip count instead of the one-visit kernel
    strips = (4 if frames >= 64 else 8 if frames >= 24 else 16) if os.environ.get("MMFN_BEV_STRIPS") else -1
    fn = lambda: ops.bev_scatter(pts, strips)
    warm, cold = chain_us(fn), cold_us(fn)
    byts = frames * 1048576.0
    out.append(dict(frames=frames, algorithmic_bytes=byts, warm_us=warm, warm_gbs=byts / warm / 1e3, warm_frac=byts / warm / 1e3 / PEAK,
                    cold_us=cold, cold_gbs=byts / cold / 1e3, cold_frac=byts / cold / 1e3 / PEAK, peak_gbs=PEAK))
print(json.dumps(out))
