#!/usr/bin/env python3
"""GPU: the two narrow fusion transformers (n_embd 64 / 128, 8 blocks, T = 192) -- whole-GPT kernels (csrc/gpt_small.cu)
against the per-op chain they replace.  CUDA-event timing of gpt.fwd and gpt.fwd + gpt.bwd, each captured in a CUDA graph
(as in the training step) so that launch overhead is what the step sees; median of 20 replays after 5 warm-ups."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmfn_b200 import ops  # noqa: E402
from mmfn_b200._lib import lib  # noqa: E402
from mmfn_b200.config import GlobalConfig  # noqa: E402
from mmfn_b200.model_rad import MMFN, _Aux  # noqa: E402


def timed_graph(body, n=20, warm=5):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        body(); body()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    l0 = lib().launches
    g = torch.cuda.CUDAGraph()
    cap = torch.cuda.Stream()
    cap.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cap):
        g.capture_begin()
        body()
        g.capture_end()
    torch.cuda.current_stream().wait_stream(cap)
    torch.cuda.synchronize()
    launches = lib().launches - l0
    ts = []
    for i in range(warm + n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], launches


def main():
    dev = torch.device("cuda:0")
    out = []
    for prec in ("tf32", "bf16"):
        ops.set_precision(prec)
        model = MMFN(GlobalConfig(), dev)
        if prec == "bf16":
            model.store.sync_shadow()
        for B in (16, 32):
            for site in (0, 1):
                gpt = model.net.gpts[site]
                C, T = gpt.C, gpt.T
                feats = [torch.randn(B, 16, 16, C, device=dev) for _ in range(3)]
                vel = torch.randn(B, 1, device=dev)
                dtok = torch.randn(B, T, C, device=dev)
                dfeats = [torch.zeros_like(f) for f in feats]
                r = dict(prec=prec, B=B, C=C, T=T)
                for fused in (0, 1):
                    ops.FUSE_GPT = 2 * fused

                    def fwd():
                        gpt.fwd(feats, vel, 7, True)

                    def fwd_bwd():
                        gpt.fwd(feats, vel, 7, True)
                        gpt.bwd(dtok, dfeats)
                        _Aux.join_all()
                    r[f"fwd_us_fused{fused}"], r[f"fwd_launches_fused{fused}"] = timed_graph(fwd)
                    r[f"fwdbwd_us_fused{fused}"], r[f"fwdbwd_launches_fused{fused}"] = timed_graph(fwd_bwd)
                # phase timeline of the first CTA (one eager launch): ns per phase, averaged over blocks 1..7
                ops.FUSE_GPT = 2
                tr = torch.zeros(10 * 8, dtype=torch.int64, device=dev)
                lib().gpt_small_trace(tr.data_ptr())
                gpt.fwd(feats, vel, 7, True)
                torch.cuda.synchronize()
                lib().gpt_small_trace(0)
                t = tr.view(8, 10)[:, :9].cpu()
                names = ["ln1", "qkv", "sync1", "attention", "sync2", "proj", "ln2", "mlp"]
                r["phase_ns"] = {n: round(float((t[1:, i + 1] - t[1:, i]).float().mean()), 0) for i, n in enumerate(names)}
                r["block_us"] = round(float((t[1:, 8] - t[1:, 0]).float().mean()) / 1e3, 2)
                out.append(r)
                print(json.dumps(r), flush=True)
        del model
    ops.FUSE_GPT = 1
    ops.set_precision("tf32")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
