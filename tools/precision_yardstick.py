#!/usr/bin/env python3
"""GPU: how far do reduced-precision paths sit from the fp32 result -- OURS and the REFERENCE'S OWN?

Same golden weights, same seeded batch, dropout 0, train-mode BatchNorm.  Truth = the CPU fp32 oracle (pinned to the
reference).  Columns:
  ref_fp32_gpu   unmodified reference nn.Module (baseline/_ref) on the GPU, TF32 disabled        -> fp32 re-ordering noise
  ref_tf32       the same with torch.backends.{cuda.matmul,cudnn}.allow_tf32 = True              -> cuBLAS / cuDNN TF32
  ref_bf16       the same under torch.autocast(bfloat16)                                         -> what BASELINE configs[2] means
  ours_fp32 / ours_tf32 / ours_bf16   this repo's exact-fp32 SIMT, TF32 tcgen05 and bf16 tcgen05 paths
For each: waypoint L1 / max, |loss error|, whole-model gradient cosine, median / p90 per-tensor relative gradient error
(attn.key.bias tensors excluded: their true gradient is exactly zero -- a constant added to every key shifts all
scores of a softmax row equally -- so both sides hold pure rounding noise).
Writes one JSON object; feeds the tolerances asserted in tests/ and the parity table in DESIGN.md.
"""
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mmfn_b200 import ops, synthetic  # noqa: E402
from mmfn_b200.config import GlobalConfig  # noqa: E402
from mmfn_b200.engine import TrainEngine  # noqa: E402
from mmfn_b200.model_rad import MMFN  # noqa: E402
from oracle import bev_oracle, mmfn_oracle  # noqa: E402


def stats(grads, ograds, pred, opred, loss, oloss):
    dot = n1 = n2 = 0.0
    rel = []
    for k, g in ograds.items():
        if g is None or k.endswith("attn.key.bias"):
            continue
        got = grads[k].detach().double().cpu()
        gd = g.double()
        dot += (got * gd).sum().item(); n1 += got.pow(2).sum().item(); n2 += gd.pow(2).sum().item()
        rel.append(((got - gd).norm() / gd.norm().clamp_min(1e-12)).item())
    rel.sort()
    d = (pred.detach().float().cpu() - opred).abs()
    return dict(waypoint_l1=d.mean().item(), waypoint_max=d.max().item(), loss_abs_err=abs(float(loss) - float(oloss)),
                grad_cosine=dot / (n1 ** 0.5 * n2 ** 0.5), grad_rel_median=rel[len(rel) // 2], grad_rel_p90=rel[int(0.9 * len(rel))],
                grad_rel_max=rel[-1])


def reference_columns(B, b, lidar_ref, sd, dev, ograds, opred, oloss):
    ref = os.path.join(ROOT, "baseline", "_ref", "team_code")
    if not os.path.isdir(ref):
        return {"unavailable": "baseline/_ref missing (tools/install_ref.py)"}
    sys.path.insert(0, ref)
    sys.modules.setdefault("torch._six", types.SimpleNamespace(string_classes=(str, bytes)))
    import torchvision
    _o = torchvision.models.resnet34
    torchvision.models.resnet34 = lambda pretrained=False, **k: _o(weights=None, **k)
    from mmfn_utils.models import model_rad
    from mmfn_utils.datasets.config import GlobalConfig as RefConfig
    out = {}
    d = dict(fronts=b["rgb_u8"].float(), lidars=lidar_ref, lane=b["lane"], lane_num=b["lane_num"].float(), radar=b["radar"],
             radar_adj=b["radar_adj"], tp=b["target_point"], vel=b["velocity"], gt=b["gt_waypoints"])
    d = {k: v.to(dev) for k, v in d.items()}
    for name, tf32, autocast in (("ref_fp32_gpu", False, False), ("ref_tf32", True, False), ("ref_bf16", True, True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        net = model_rad.MMFN(RefConfig(embd_pdrop=0.0, attn_pdrop=0.0, resid_pdrop=0.0), dev).to(dev)
        net.load_state_dict(sd)
        net.train()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            vm = [[d["lane"]], [d["lane_num"]], d["lane"].shape[1]]
            pred = net([d["fronts"]], [d["lidars"]], None, vm, [d["radar"]], [d["radar_adj"]], d["tp"], d["vel"])
            loss = F.l1_loss(pred.float(), d["gt"], reduction="none").mean()
        loss.backward()
        grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in net.named_parameters()}
        out[name] = stats(grads, ograds, pred, opred, loss.item(), oloss)
        del net
    return out


def main(B):
    dev = torch.device("cuda:0")
    cfg = GlobalConfig(embd_pdrop=0.0, attn_pdrop=0.0, resid_pdrop=0.0)
    b = synthetic.synth_batch(B)
    db = {k: v.to(dev) for k, v in b.items()}
    lidar_ref = torch.from_numpy(np.stack([bev_oracle.lidar_to_histogram_features(p[:, :3].numpy()) for p in b["points"]]))
    inputs = (b["rgb_u8"].float(), lidar_ref, b["lane"], b["lane_num"], b["radar"], b["radar_adj"], b["target_point"], b["velocity"])
    out = {"B": B}
    sd = ograds = None
    for mode in ("fp32", "tf32", "bf16"):
        ops.set_precision(mode)
        model = MMFN(cfg, dev)
        if sd is None:
            sd = synthetic.fill_golden_weights(model.state_dict(), 42)
            oloss, opred, ograds = mmfn_oracle.train_step({k: v.clone() for k, v in sd.items()}, cfg,
                                                          dict(inputs=inputs, gt_waypoints=b["gt_waypoints"]))
            oloss = oloss.item()
        model.load_state_dict(sd)
        eng = TrainEngine(model)
        loss = eng.forward_backward(db).item()
        torch.cuda.synchronize()
        grads = {k: model.store.torch_view(k, grad=True) for k, g in ograds.items() if g is not None}
        out["ours_" + mode] = stats(grads, ograds, eng.last_pred, opred, loss, oloss)
        del eng, model
    ops.set_precision("tf32")
    out.update(reference_columns(B, b, lidar_ref, sd, dev, ograds, opred, oloss))
    print(json.dumps(out))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 16)
