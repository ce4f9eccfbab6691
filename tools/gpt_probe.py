#!/usr/bin/env python3
"""GPU developer aid: run one fusion transformer (fwd+bwd) eagerly so `ncu --cache-control none` can list
per-kernel durations with a warm L2.  usage: gpt_probe.py <gpt index 0..3> [B]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmfn_b200.config import GlobalConfig
from mmfn_b200.model_rad import MMFN, _Aux

i = int(sys.argv[1]); B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda:0")
model = MMFN(GlobalConfig(), dev)
net = model.net
_Aux.enabled = False
WID = (64, 128, 256, 512); HW = (64, 32, 16, 8)
nmod = 3 if i < 3 else 4
feats = [torch.randn(B, HW[i], HW[i], WID[i], device=dev) for _ in range(nmod)]
dfe = [torch.zeros_like(f) for f in feats]
dtok = torch.randn(B, nmod * 64, WID[i], device=dev) * 1e-3
vel = torch.rand(B, device=dev)
for it in range(3):
    torch.cuda.synchronize()
    if it == 2:
        torch.cuda.profiler.start()
    net.gpts[i].fwd(feats, vel, 7, True)
    net.gpts[i].bwd(dtok, dfe)
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
