#!/usr/bin/env python3
"""GPU: error of the CUDA path against the CPU oracle, exact-fp32 kernels vs the TF32 tensor-core
kernels (waypoint L1, loss, worst per-tensor gradient error).  Feeds the tolerances in tests/."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmfn_b200 import ops, synthetic  # noqa: E402
from mmfn_b200.config import GlobalConfig  # noqa: E402
from mmfn_b200.engine import TrainEngine  # noqa: E402
from mmfn_b200.model_rad import MMFN  # noqa: E402
from oracle import bev_oracle, mmfn_oracle  # noqa: E402


def main(B=2):
    dev = torch.device("cuda:0")
    cfg = GlobalConfig(embd_pdrop=0.0, attn_pdrop=0.0, resid_pdrop=0.0)
    b = synthetic.synth_batch(B)
    db = {k: v.to(dev) for k, v in b.items()}
    lidar_ref = torch.from_numpy(np.stack([bev_oracle.lidar_to_histogram_features(p[:, :3].numpy()) for p in b["points"]]))
    inputs = (b["rgb_u8"].float(), lidar_ref, b["lane"], b["lane_num"], b["radar"], b["radar_adj"],
              b["target_point"], b["velocity"])
    out = {}
    ograds = None
    for mode in (False, True):
        ops.TF32 = mode
        model = MMFN(cfg, dev)
        sd = synthetic.fill_golden_weights(model.state_dict(), 42)
        model.load_state_dict(sd)
        eng = TrainEngine(model)
        loss = eng.forward_backward(db)
        torch.cuda.synchronize()
        if ograds is None:
            oloss, opred, ograds = mmfn_oracle.train_step({k: v.clone() for k, v in sd.items()}, cfg,
                                                          dict(inputs=inputs, gt_waypoints=b["gt_waypoints"]))
        errs, coss = [], []
        dot = n1 = n2 = 0.0
        for k, g in ograds.items():
            if g is None:
                continue
            got = model.store.torch_view(k, grad=True).cpu()
            errs.append(((got - g).norm().item() / max(g.norm().item(), 1e-6), k))
            d, a, b_ = (got.double() * g.double()).sum().item(), got.double().norm().item(), g.double().norm().item()
            dot, n1, n2 = dot + d, n1 + a * a, n2 + b_ * b_
            if b_ > 1e-5:
                coss.append((d / max(a * b_, 1e-30), k))
        errs.sort(reverse=True)
        coss.sort()
        out["tf32" if mode else "fp32"] = {
            "global_grad_cosine": dot / (n1 ** 0.5 * n2 ** 0.5), "worst_tensor_cosines": coss[:5],
            "waypoint_L1": (eng.last_pred.cpu() - opred).abs().mean().item(),
            "loss_abs_err": abs(loss.item() - oloss.item()),
            "grad_rel_err_worst": errs[:5], "grad_rel_err_median": errs[len(errs) // 2][0],
        }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 2)
