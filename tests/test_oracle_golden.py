"""The oracle (oracle/) pinned against vectors produced by the UNMODIFIED reference
(tools/make_goldens.py, run in the build container).  CPU only."""
import json
import os

import numpy as np
import torch

from mmfn_b200 import synthetic
from mmfn_b200.config import GlobalConfig
from oracle import bev_oracle, mmfn_oracle


def probe(t, n=16):
    f = t.detach().reshape(-1).double()
    idx = torch.linspace(0, f.numel() - 1, n).long()
    return np.concatenate([[f.norm().item(), f.sum().item()], f[idx].numpy()])


def test_bev_oracle_matches_reference_goldens(golden_dir):
    gold = np.load(os.path.join(golden_dir, "bev_golden.npz"))
    assert len(gold.files) == 5
    for key in gold.files:
        seed, n = int(key.split("_")[0][1:]), int(key.split("_n")[1])
        pts = synthetic.synth_points(seed, n)
        feat = bev_oracle.lidar_to_histogram_features(pts[:, :3])
        assert feat.dtype == np.float32 and feat.shape == (2, 256, 256)
        ref = (gold[key].astype(np.float64) / 5).astype(np.float32)
        assert np.array_equal(feat, ref), key


def test_bev_oracle_edge_semantics():
    pts = np.array([[16.0, 8.0, -2.0], [-16.0, -24.0, -2.0000002], [16.000002, 0, 0], [0, 8.000001, 0],
                    [np.nan, 0, 0], [0, 0, np.nan], [-16.0, 8.0, -1.9999999]], dtype=np.float32)
    c = bev_oracle.bev_counts(pts)
    assert c[0, 255, 255] == 1 and c[0, 0, 0] == 1 and c[1, 0, 255] == 1 and c.sum() == 3
    many = np.tile(np.array([[1.0, 1.0, 0.0]], dtype=np.float32), (9, 1))
    f = bev_oracle.lidar_to_histogram_features(many)
    assert f[1, 136, 200] == 1.0 and f.sum() == 1.0


def _oracle_step(B=2):
    torch.manual_seed(0)
    cfg = GlobalConfig(embd_pdrop=0.0, attn_pdrop=0.0, resid_pdrop=0.0)
    keys = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "state_dict_keys.json")))
    shapes = {k: torch.empty(v[0], dtype=getattr(torch, v[1].split(".")[1])) for k, v in keys.items()}
    sd = synthetic.fill_golden_weights(shapes, 42)
    b = synthetic.synth_batch(B)
    lidar = torch.from_numpy(np.stack([bev_oracle.lidar_to_histogram_features(p[:, :3].numpy()) for p in b["points"]]))
    inputs = (b["rgb_u8"].float(), lidar, b["lane"], b["lane_num"], b["radar"], b["radar_adj"],
              b["target_point"], b["velocity"])
    return sd, cfg, dict(inputs=inputs, gt_waypoints=b["gt_waypoints"])


def test_model_oracle_matches_reference_goldens(golden_dir):
    gold = np.load(os.path.join(golden_dir, "mmfn_golden_b2.npz"))
    sd, cfg, batch = _oracle_step(2)
    opt = {"t": 0, "m": {}, "v": {}}
    loss, pred, grads = mmfn_oracle.train_step(sd, cfg, batch, opt_state=opt)
    # fp32 CPU vs fp32 CPU, same op order up to a few reassociations
    assert abs(loss.item() - float(gold["loss"])) < 1e-5
    assert np.abs(pred.numpy() - gold["pred_wp"]).max() < 2e-5
    unused = set(gold["unused"].tolist())
    assert {k for k, g in grads.items() if g is None} == unused and len(unused) == 21
    worst = 0.0
    for k, g in grads.items():
        if g is None:
            continue
        ref = gold["grad/" + k]
        got = probe(g, 6)
        # relative to the gradient's norm; entry 1 (the plain sum) is dropped: for weights that
        # feed a BatchNorm it is a pure-cancellation quantity.  key.bias gradients are exactly 0
        # in theory (softmax shift invariance), hence the absolute floor.
        denom = max(abs(ref[0]), 1e-6)
        worst = max(worst, np.abs(np.delete(got - ref, 1)).max() / denom)
    assert worst < 5e-3, worst   # fp32 re-association noise through ~100 train-mode BN layers at B=2
    for k in gold.files:
        if k.startswith("buf/"):
            assert np.allclose(probe(sd[k[4:]], 4), gold[k], rtol=1e-4, atol=1e-5), k
        if k.startswith("adam/"):
            # first AdamW step moves every weight by ~lr*sign(g): elements whose gradient is ~0 may
            # flip sign, so the plain sum (entry 1) is excluded and the norm gets a loose bound
            got, ref = probe(sd[k[5:]], 6), gold[k]
            assert np.allclose(got[2:], ref[2:], rtol=1e-5, atol=1e-6) and abs(got[0] - ref[0]) < 1e-4 * ref[0], k


def test_transfuser_oracle_matches_reference_goldens(golden_dir):
    """RGB+LiDAR-only variant (BASELINE configs[3]): oracle/transfuser_oracle.py vs the unmodified
    benchmarks/transfuser/model.py:TransFuser (goldens by tools/make_goldens.py)."""
    from mmfn_b200.params import param_spec
    from oracle import transfuser_oracle
    gold = np.load(os.path.join(golden_dir, "transfuser_golden_b2.npz"))
    keys = json.load(open(os.path.join(golden_dir, "transfuser_state_dict_keys.json")))
    cfg = GlobalConfig(embd_pdrop=0.0, attn_pdrop=0.0, resid_pdrop=0.0)
    # the native parameter inventory reproduces the reference state_dict (order, shapes, dtypes)
    spec = param_spec(cfg, "transfuser")
    assert [k for k, _, _ in spec] == list(keys.keys())
    for k, shape, kind in spec:
        assert list(shape) == keys[k][0], k
    shapes = {k: torch.empty(v[0], dtype=getattr(torch, v[1].split(".")[1])) for k, v in keys.items()}
    sd = synthetic.fill_golden_weights(shapes, 42)
    b = synthetic.synth_batch(2)
    lidar = torch.from_numpy(np.stack([bev_oracle.lidar_to_histogram_features(p[:, :3].numpy()) for p in b["points"]]))
    batch = dict(inputs=(b["rgb_u8"].float(), lidar, b["target_point"], b["velocity"]), gt_waypoints=b["gt_waypoints"])
    loss, pred, grads = transfuser_oracle.train_step(sd, cfg, batch)
    assert abs(loss.item() - float(gold["loss"])) < 1e-5
    assert np.abs(pred.numpy() - gold["pred_wp"]).max() < 2e-5
    worst = 0.0
    for k, g in grads.items():
        assert g is not None, k
        ref, got = gold["grad/" + k], probe(g, 6)
        worst = max(worst, np.abs(np.delete(got - ref, 1)).max() / max(abs(ref[0]), 1e-6))
    assert worst < 5e-3, worst
    for k in gold.files:
        if k.startswith("buf/"):
            assert np.allclose(probe(sd[k[4:]], 4), gold[k], rtol=1e-4, atol=1e-5), k


def _variant_batch(variant, B=2):
    b = synthetic.synth_batch(B)
    lidar = torch.from_numpy(np.stack([bev_oracle.lidar_to_histogram_features(p[:, :3].numpy()) for p in b["points"]]))
    lane = synthetic.synth_map_images(B).float() if variant == "img" else b["lane"]
    inputs = (b["rgb_u8"].float(), lidar, lane, b["lane_num"], b["radar"], b["radar_adj"], b["target_point"], b["velocity"])
    return dict(inputs=inputs, gt_waypoints=b["gt_waypoints"])


def test_vec_and_img_variant_oracles_match_reference_goldens(golden_dir):
    """model_vec.MMFN (no radar) and model_img.MMFN (rasterised map image, default train.yaml entry point)."""
    import functools
    from mmfn_b200.params import param_spec
    cfg = GlobalConfig(embd_pdrop=0.0, attn_pdrop=0.0, resid_pdrop=0.0)
    for variant, n_unused in (("vec", 21), ("img", 0)):
        gold = np.load(os.path.join(golden_dir, f"{variant}_golden_b2.npz"))
        keys = json.load(open(os.path.join(golden_dir, f"{variant}_state_dict_keys.json")))
        spec = param_spec(cfg, variant)
        assert [k for k, _, _ in spec] == list(keys.keys()), variant
        assert all(list(shape) == keys[k][0] for k, shape, _ in spec), variant
        shapes = {k: torch.empty(v[0], dtype=getattr(torch, v[1].split(".")[1])) for k, v in keys.items()}
        sd = synthetic.fill_golden_weights(shapes, 42)
        fwd = functools.partial(mmfn_oracle.forward, variant=variant)
        loss, pred, grads = mmfn_oracle.train_step(sd, cfg, _variant_batch(variant), forward_fn=fwd)
        assert abs(loss.item() - float(gold["loss"])) < 1e-5, variant
        assert np.abs(pred.numpy() - gold["pred_wp"]).max() < 2e-5, variant
        unused = set(gold["unused"].tolist())
        assert {k for k, g in grads.items() if g is None} == unused and len(unused) == n_unused, variant
        worst = 0.0
        for k, g in grads.items():
            if g is None:
                continue
            ref, got = gold["grad/" + k], probe(g, 6)
            worst = max(worst, np.abs(np.delete(got - ref, 1)).max() / max(abs(ref[0]), 1e-6))
        assert worst < 5e-3, (variant, worst)


def test_model_oracle_inference_path_matches_reference_golden(golden_dir):
    """eval() mode (BatchNorm running statistics, dropout inactive whatever the config says): the path the e2e agents
    run (e2e_agent/mmfn_radar.py:296-306).  The golden running statistics are deliberately far from the batch
    statistics, so the outputs are O(1e3): compared relative to their magnitude."""
    gold = np.load(os.path.join(golden_dir, "mmfn_eval_golden_b1.npz"))["pred_wp"]
    sd, _, batch = _oracle_step(1)
    cfg = GlobalConfig()                                     # reference default: dropout 0.1 -- must be a no-op in eval
    with torch.no_grad():
        pred = mmfn_oracle.forward(sd, cfg, *batch["inputs"], train=False)
    assert pred.shape == gold.shape == (1, 4, 2)
    rel = np.abs(pred.numpy() - gold).mean() / np.abs(gold).mean()
    assert rel < 1e-5, rel
