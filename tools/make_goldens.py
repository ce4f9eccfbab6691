#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Runs only in the build container (the reference is not on the GPU box).  Two shims, both
outside the reference tree: torchvision resnet34(pretrained=True) ->
random init (no network), and a stub for the removed torch._six module.
Outputs (small):
  bev_golden.npz     : packed BEV histograms of the reference lidar_to_histogram_features
  mmfn_golden_b2.npz : pred_wp / loss / per-parameter gradient norms+probes of model_rad.MMFN
                       (dropout 0, train-mode BN) on synth_batch(2) with fill_golden_weights(42),
                       plus BN running stats after the step and intermediate activation probes
  state_dict_keys.json: reference state_dict keys, shapes, dtypes
"""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/team_code")
sys.modules["torch._six"] = types.SimpleNamespace(string_classes=(str, bytes))
import torchvision  # noqa: E402

_orig34 = torchvision.models.resnet34
torchvision.models.resnet34 = lambda pretrained=False, **k: _orig34(weights=None, **k)

from mmfn_utils.models import model_rad  # noqa: E402
from mmfn_utils.datasets.config import GlobalConfig  # noqa: E402
from mmfn_utils.datasets.dataloader import lidar_to_histogram_features  # noqa: E402
from mmfn_b200 import synthetic  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)


def probe(t, n=16):
    """A few deterministic elements + norm of a tensor: small but position-sensitive."""
    f = t.detach().reshape(-1).double()
    idx = torch.linspace(0, f.numel() - 1, n).long()
    return np.concatenate([[f.norm().item(), f.sum().item()], f[idx].numpy()])


def bev():
    out = {}
    for seed, n in [(1234, 32768), (1235, 32768), (77, 1000), (78, 1), (79, 0)]:
        pts = synthetic.synth_points(seed, n)
        feat = lidar_to_histogram_features(pts[:, :3])
        out[f"s{seed}_n{n}"] = np.round(feat * 5).astype(np.uint8)
    np.savez_compressed(os.path.join(GOLD, "bev_golden.npz"), **out)


def model(B=2):
    torch.manual_seed(0)
    torch.set_num_threads(8)
    cfg = GlobalConfig(embd_pdrop=0.0, attn_pdrop=0.0, resid_pdrop=0.0)
    net = model_rad.MMFN(cfg, "cpu")
    sd = net.state_dict()
    json.dump({k: [list(v.shape), str(v.dtype)] for k, v in sd.items()},
              open(os.path.join(GOLD, "state_dict_keys.json"), "w"), indent=0)
    net.load_state_dict(synthetic.fill_golden_weights(sd, 42))
    net.train()
    b = synthetic.synth_batch(B)
    lidar = torch.from_numpy(np.stack([lidar_to_histogram_features(p[:, :3].numpy()) for p in b["points"]]))
    fronts = [b["rgb_u8"].float()]
    vectormaps = [[b["lane"]], [b["lane_num"].float()], b["lane"].shape[1]]
    acts = {}
    enc = net.encoder
    hooks = [
        enc.image_encoder.features.layer1.register_forward_hook(lambda m, i, o: acts.__setitem__("img_l1", o)),
        enc.lidar_encoder._model.layer1.register_forward_hook(lambda m, i, o: acts.__setitem__("lid_l1", o)),
        enc.vectornet_encoder.register_forward_hook(lambda m, i, o: acts.__setitem__("map_gen", o)),
        enc.radar_encoder.register_forward_hook(lambda m, i, o: acts.__setitem__("rad", o)),
        enc.register_forward_hook(lambda m, i, o: acts.__setitem__("fused", o)),
    ]
    pred = net(fronts, [lidar], None, vectormaps, [b["radar"]], [b["radar_adj"]], b["target_point"], b["velocity"])
    loss = torch.nn.functional.l1_loss(pred, b["gt_waypoints"], reduction="none").mean()
    loss.backward()
    for h in hooks:
        h.remove()
    out = {"pred_wp": pred.detach().numpy(), "loss": np.float64(loss.item())}
    for k, v in acts.items():
        out["act/" + k] = probe(v)
    unused = []
    for k, p in net.named_parameters():
        if p.grad is None:
            unused.append(k)
        else:
            out["grad/" + k] = probe(p.grad, 6)
    for k, v in net.state_dict().items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            out["buf/" + k] = probe(v, 4)
    out["unused"] = np.array(unused)
    # one AdamW step with torch.optim defaults as phase2_train_net.py:256
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4)
    opt.step()
    for k, p in net.named_parameters():
        if k.endswith("conv1.weight") or "blocks.0.attn.key" in k or k.startswith("decoder") or "generator.3.bias" in k:
            out["adam/" + k] = probe(p.data, 6)
    np.savez_compressed(os.path.join(GOLD, f"mmfn_golden_b{B}.npz"), **out)
    print("loss", loss.item(), "unused params", len(unused))


def describe(x):
    """Type/shape/dtype tree of a collated batch (lists, dicts, tensors, scalars)."""
    if isinstance(x, torch.Tensor):
        return ["tensor", list(x.shape), str(x.dtype)]
    if isinstance(x, dict):
        return {k: describe(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [describe(v) for v in x]
    return [type(x).__name__, x if isinstance(x, (int, float, bool, str)) else None]


def collate():
    """Reference collate_single_cpu + PRE_Data adjacency on 3 synthetic samples with 70/128/93 lanes."""
    from mmfn_utils.datasets.data_utils import collate_single_cpu
    samples = [synthetic.synth_sample(i, lidar_to_histogram_features) for i in range(3)]
    for smp in samples:                                     # PRE_Data.__getitem__, dataloader.py:379-384
        rows = [smp["radar"][0][:, 1] - smp["radar"][0][i, 1] for i in range(81)]
        smp["radar_adj"] = np.array(rows)
    out = collate_single_cpu(samples)
    json.dump(describe(out), open(os.path.join(GOLD, "collate_structure.json"), "w"), indent=0)
    np.savez_compressed(os.path.join(GOLD, "collate_golden.npz"),
                        lanes=out["vectormaps"][0][0].numpy(), lane_nums=out["vectormaps"][0][1].numpy(),
                        lmax=np.int64(out["vectormaps"][0][2]), radar_adj=out["radar_adj"].numpy(),
                        radar=out["radar"][0].numpy(), velocity=out["velocity"].numpy(),
                        target_point=torch.stack(out["target_point"], 1).numpy(),
                        waypoints=torch.stack([torch.stack(w, 1) for w in out["waypoints"]], 1).numpy(),
                        fronts_sum=np.int64(out["fronts"][0].long().sum()), lidars_sum=np.float64(out["lidars"][0].double().sum()))


def variants(B=2):
    """model_vec.MMFN (no radar) and model_img.MMFN (rasterised map image; the default train.yaml entry point)."""
    from mmfn_utils.models import model_img, model_vec
    b = synthetic.synth_batch(B)
    lidar = torch.from_numpy(np.stack([lidar_to_histogram_features(p[:, :3].numpy()) for p in b["points"]]))
    vectormaps = [[b["lane"]], [b["lane_num"].float()], b["lane"].shape[1]]
    maps = [synthetic.synth_map_images(B).float()]
    for name, mod in (("vec", model_vec), ("img", model_img)):
        torch.manual_seed(0)
        cfg = GlobalConfig(embd_pdrop=0.0, attn_pdrop=0.0, resid_pdrop=0.0)
        net = mod.MMFN(cfg, "cpu")
        sd = net.state_dict()
        json.dump({k: [list(v.shape), str(v.dtype)] for k, v in sd.items()},
                  open(os.path.join(GOLD, f"{name}_state_dict_keys.json"), "w"), indent=0)
        net.load_state_dict(synthetic.fill_golden_weights(sd, 42))
        net.train()
        pred = net([b["rgb_u8"].float()], [lidar], maps, vectormaps, [b["radar"]], [b["radar_adj"]], b["target_point"], b["velocity"])
        loss = torch.nn.functional.l1_loss(pred, b["gt_waypoints"], reduction="none").mean()
        loss.backward()
        out = {"pred_wp": pred.detach().numpy(), "loss": np.float64(loss.item())}
        unused = []
        for k, p in net.named_parameters():
            if p.grad is None:
                unused.append(k)
            else:
                out["grad/" + k] = probe(p.grad, 6)
        out["unused"] = np.array(unused)
        np.savez_compressed(os.path.join(GOLD, f"{name}_golden_b{B}.npz"), **out)
        print(name, "loss", loss.item(), "unused params", len(unused))


def transfuser(B=2):
    """benchmarks/transfuser/model.py:TransFuser (RGB + LiDAR only, BASELINE configs[3]) on synth_batch(2)."""
    from benchmarks.transfuser import model as tf_model
    from benchmarks.transfuser.config import GlobalConfig as TFConfig
    torch.manual_seed(0)
    torch.set_num_threads(8)
    cfg = TFConfig()
    cfg.embd_pdrop = cfg.attn_pdrop = cfg.resid_pdrop = 0.0
    net = tf_model.TransFuser(cfg, "cpu")
    sd = net.state_dict()
    json.dump({k: [list(v.shape), str(v.dtype)] for k, v in sd.items()},
              open(os.path.join(GOLD, "transfuser_state_dict_keys.json"), "w"), indent=0)
    net.load_state_dict(synthetic.fill_golden_weights(sd, 42))
    net.train()
    b = synthetic.synth_batch(B)
    lidar = torch.from_numpy(np.stack([lidar_to_histogram_features(p[:, :3].numpy()) for p in b["points"]]))
    pred = net([b["rgb_u8"].float()], [lidar], b["target_point"], b["velocity"])
    loss = torch.nn.functional.l1_loss(pred, b["gt_waypoints"], reduction="none").mean()
    loss.backward()
    out = {"pred_wp": pred.detach().numpy(), "loss": np.float64(loss.item())}
    for k, p in net.named_parameters():
        assert p.grad is not None, k
        out["grad/" + k] = probe(p.grad, 6)
    for k, v in net.state_dict().items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            out["buf/" + k] = probe(v, 4)
    np.savez_compressed(os.path.join(GOLD, f"transfuser_golden_b{B}.npz"), **out)
    print("transfuser loss", loss.item())


def eval_mode(B=1):
    """Inference path of the reference (what the e2e agents run, e2e_agent/mmfn_radar.py:296-306): eval() mode --
    BatchNorm running statistics, dropout off -- with the reference DEFAULT dropout config."""
    torch.manual_seed(0)
    cfg = GlobalConfig()
    net = model_rad.MMFN(cfg, "cpu")
    net.load_state_dict(synthetic.fill_golden_weights(net.state_dict(), 42))
    net.eval()
    b = synthetic.synth_batch(B)
    lidar = torch.from_numpy(np.stack([lidar_to_histogram_features(p[:, :3].numpy()) for p in b["points"]]))
    vectormaps = [[b["lane"]], [b["lane_num"].float()], b["lane"].shape[1]]
    with torch.no_grad():
        pred = net([b["rgb_u8"].float()], [lidar], None, vectormaps, [b["radar"]], [b["radar_adj"]], b["target_point"], b["velocity"])
    np.savez_compressed(os.path.join(GOLD, f"mmfn_eval_golden_b{B}.npz"), pred_wp=pred.numpy())
    print("eval pred", pred.abs().mean().item())


def control():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from control_fixture import control_inputs
    # MMFN.control_pid (model_rad.py:697-739) driven over a sequence: the two PID controllers carry state.
    net = model_rad.MMFN(GlobalConfig(), "cpu")
    rows = []
    for wp, v in control_inputs():
        steer, throttle, brake, meta = net.control_pid(torch.from_numpy(wp.copy()), torch.from_numpy(v.copy()))
        rows.append([float(steer), float(throttle), float(brake), meta["desired_speed"], meta["angle"], meta["delta"],
                     meta["aim"][0], meta["aim"][1], meta["speed"]])
    np.savez_compressed(os.path.join(GOLD, "control_pid_golden.npz"), rows=np.asarray(rows, dtype=np.float64))


if __name__ == "__main__":
    if "--control-only" in sys.argv:
        control()
        sys.exit(0)
    if "--eval-only" in sys.argv:
        eval_mode(1)
        sys.exit(0)
    if "--transfuser-only" in sys.argv:
        transfuser(2)
        sys.exit(0)
    if "--variants-only" in sys.argv:
        variants(2)
        sys.exit(0)
    bev()
    collate()
    model(2)
    transfuser(2)
    variants(2)
    control()
    eval_mode(1)
