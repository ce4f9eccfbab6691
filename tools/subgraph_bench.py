#!/usr/bin/env python3
"""GPU: the polyline Subgraph forward (reference model_rad.py:260-283) -- one-launch kernel (csrc/vectornet.cu) against the
per-layer kernels it replaces (lane_to_vector + 3 x [GEMM, LayerNorm, max-pool/concat] + segment max = 11 launches).
Times both with CUDA events (median of 20 after 5 warm-ups, L2 flushed between runs) at the training shape
(B=16 x 128 lanes x 10 nodes) and at BASELINE configs[4] (B=128 x 256 polylines x 20 nodes); reports the HBM traffic each
formulation needs (tensors written for the backward + inputs) and the fp32 FMA rate of the fused kernel."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmfn_b200 import ops, synthetic  # noqa: E402
from mmfn_b200.config import GlobalConfig  # noqa: E402
from mmfn_b200.model_rad import MMFN  # noqa: E402


def timed(fn, flush, n=20, warm=5):
    ts = []
    for i in range(warm + n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    dev = torch.device("cuda:0")
    ops.set_precision("tf32")
    model = MMFN(GlobalConfig(), dev)
    vn = model.net.vectornet
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = []
    for B, L, P in ((16, 128, 10), (32, 128, 10), (128, 256, 20)):
        lane = synthetic.synth_batch(B, first_index=300, n_lanes=L, n_nodes=P, n_points=0)["lane"].to(dev)
        G, V = B * L, P - 1
        layers = [(lin.w, lin.b, ln.g, ln.b) for lin, ln in vn.sub]

        def unfused():
            x = ops.lane_to_vector(lane)
            for lin, ln in vn.sub:
                x, _ = ops.subgraph_pool_fwd(ln.fwd(lin.fwd(x)), G, V)
            return ops.segmax_fwd(x, G, V)

        def fused():
            return ops.subgraph_fused_fwd(lane, layers)
        rows = G * V
        macs = rows * (7 * 64 + 2 * 64 * 64) + G * 2 * 64 * 64
        written = rows * 4 * (7 + 3 * 64 + 6 + 2 * 128) + G * 4 * (3 * 64 + 2 * 128)
        r = dict(B=B, L=L, P=P, polylines=G, us_unfused=timed(unfused, flush), us_fused=timed(fused, flush),
                 fused_bytes_mb=(written + lane.numel() * 4) / 1e6)
        r["fused_gbps"] = r["fused_bytes_mb"] / r["us_fused"] * 1e3
        r["fused_fp32_tflops"] = 2 * macs / r["us_fused"] * 1e-6
        r["speedup"] = r["us_unfused"] / r["us_fused"]
        out.append(r)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
