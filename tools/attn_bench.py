#!/usr/bin/env python3
"""GPU: fused attention forward kernels per transformer scale (north_star: tensor-pipe utilisation of the fusion
attention).  Times mmfn_attention_fwd_tf32 (fp32 qkv, TF32 MMAs, P / Pd stored as fp32) and mmfn_attention_fwd_bf16
(bf16 qkv, K / V resident, single-exp softmax; with P / Pd stored as bf16 for the unfused backward, and in the
stats-only mode) as CUDA-graph chains.  Algorithmic FLOPs = 4 T^2 C per (sample, layer); the roofline
denominator is the measured sustained bf16 matmul peak (MEASURED_PEAKS.json) for the bf16 kernel and a live cuBLAS
TF32 measurement for the TF32 kernel.   --one: a single launch of the bf16 kernel at (T=256, C=512, B=32) for ncu."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmfn_b200 import ops  # noqa: E402
from bench import _graph_chain_us, measure_matmul_peak, measured_peaks  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    ops.set_precision("bf16")
    if "--one" in sys.argv:
        B, T, C, nh = 32, 256, 512, 4
        qkv = torch.randn(B * T, 3 * C, device=dev).to(torch.bfloat16)
        for _ in range(3):
            ops.attention_fwd_bf16(qkv, B, T, C, nh, 0.1, 7, save_probs=False)
            ops.attention_fwd_bf16(qkv, B, T, C, nh, 0.1, 7, save_probs=True)
            ops.attention_fwd(qkv.float(), B, T, C, nh, 0.1, 7)
        torch.cuda.synchronize()
        return
    if "--trace" in sys.argv:
        # %globaltimer stamps of CTA 0 (ns): 0 entry, 1 setup done, 2 Q/K landed, 3 S visible to the softmax, 4 row statistics
        # exchanged, 5 all P tiles written, 6 first PV MMA issued, 7 all MMAs issued, 8 O visible, 9 y stored
        from mmfn_b200._lib import lib
        res = {}
        for B, T, C in ((16, 256, 512), (32, 256, 512), (32, 192, 256), (32, 192, 64)):
            qkv = torch.randn(B * T, 3 * C, device=dev).to(torch.bfloat16)
            buf = torch.zeros(16, dtype=torch.int64, device=dev)
            for _ in range(3):
                ops.attention_fwd_bf16(qkv, B, T, C, 4, 0.1, 7, save_probs=False)
            torch.cuda.synchronize()
            lib().tc_set_trace(buf.data_ptr())
            ops.attention_fwd_bf16(qkv, B, T, C, 4, 0.1, 7, save_probs=False)
            torch.cuda.synchronize()
            lib().tc_set_trace(0)
            t = buf.cpu().tolist()
            res[f"B{B}_T{T}_C{C}"] = [x - t[0] for x in t[:10]]
        print(json.dumps(res))
        return
    peaks = measured_peaks()
    tf32_peak = measure_matmul_peak(dev, True)[1]
    out = {"bf16_peak_tflops": peaks["tf_sust"], "tf32_peak_tflops": tf32_peak, "rows": []}
    nh = 4
    for B in (16, 32):
        for T, C in ((192, 64), (192, 128), (192, 256), (256, 512)):
            qkv32 = torch.randn(B * T, 3 * C, device=dev)
            qkv16 = qkv32.to(torch.bfloat16)
            fl = 4.0 * B * T * T * C
            row = dict(B=B, T=T, C=C, hs=C // nh, gflop=fl / 1e9)
            for name, fn, peak in (
                    ("tf32_store_p_drop", lambda: ops.attention_fwd(qkv32, B, T, C, nh, 0.1, 7), tf32_peak),
                    ("bf16_store_p_drop", lambda: ops.attention_fwd_bf16(qkv16, B, T, C, nh, 0.1, 7), peaks["tf_sust"]),
                    ("bf16_store_p", lambda: ops.attention_fwd_bf16(qkv16, B, T, C, nh, 0.0, 7), peaks["tf_sust"]),
                    ("bf16_stats_only", lambda: ops.attention_fwd_bf16(qkv16, B, T, C, nh, 0.1, 7, save_probs=False), peaks["tf_sust"])):
                us = _graph_chain_us(fn, chain=20)
                row[name] = dict(us=round(us, 2), tflops=round(fl / us / 1e6, 1), frac=round(fl / us / 1e6 / peak, 3))
            out["rows"].append(row)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
