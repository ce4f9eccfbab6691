#!/usr/bin/env python3
"""GPU: launch the whole-GPT forward kernel a few times (for ncu): python tools/gpt_one.py [bf16|tf32] [site] [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmfn_b200 import ops  # noqa: E402
from mmfn_b200.config import GlobalConfig  # noqa: E402
from mmfn_b200.model_rad import MMFN  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
site = int(sys.argv[2]) if len(sys.argv) > 2 else 1
B = int(sys.argv[3]) if len(sys.argv) > 3 else 32
dev = torch.device("cuda:0")
ops.set_precision(prec)
ops.FUSE_GPT = 2
model = MMFN(GlobalConfig(), dev)
if prec == "bf16":
    model.store.sync_shadow()
gpt = model.net.gpts[site]
feats = [torch.randn(B, 16, 16, gpt.C, device=dev) for _ in range(3)]
vel = torch.randn(B, 1, device=dev)
for _ in range(3):
    gpt.fwd(feats, vel, 7, True)
torch.cuda.synchronize()
print("done")
