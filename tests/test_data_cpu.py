"""CPU: the data layer against goldens produced by the reference's own
collate_single_cpu / PRE_Data adjacency code (tools/make_goldens.py:collate)."""
import json
import os
import pickle

import numpy as np
import torch

from mmfn_b200 import data, synthetic
from mmfn_b200.config import GlobalConfig
from oracle import bev_oracle


def describe(x):
    if isinstance(x, torch.Tensor):
        return ["tensor", list(x.shape), str(x.dtype)]
    if isinstance(x, dict):
        return {k: describe(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [describe(v) for v in x]
    return [type(x).__name__, x if isinstance(x, (int, float, bool, str)) else None]


def _samples(tmp_path=None):
    smp = [synthetic.synth_sample(i, bev_oracle.lidar_to_histogram_features) for i in range(3)]
    if tmp_path is None:
        for s in smp:
            s["radar_adj"] = synthetic.radar_adjacency(s["radar"][0])
        return smp
    for i, s in enumerate(smp):                              # through the pickle reader, like phase 2
        with open(os.path.join(tmp_path, f"{i}.pkl"), "wb") as fd:
            pickle.dump(s, fd)
    ds = data.PRE_Data(str(tmp_path), GlobalConfig(), "train")
    assert len(ds) == 3
    order = np.argsort([int(os.path.basename(p).split(".")[0]) for p in ds.preload_dict])
    return [ds[int(j)] for j in order]


def test_collate_matches_reference_structure_and_values(golden_dir, tmp_path):
    gold = np.load(os.path.join(golden_dir, "collate_golden.npz"))
    structure = json.load(open(os.path.join(golden_dir, "collate_structure.json")))
    out = data.collate_single_cpu(_samples(tmp_path))
    assert json.loads(json.dumps(describe(out))) == structure          # same tree, shapes and dtypes
    lanes, lane_nums, lmax = out["vectormaps"][0]
    assert np.array_equal(lanes.numpy(), gold["lanes"]) and np.array_equal(lane_nums.numpy(), gold["lane_nums"])
    assert lmax == int(gold["lmax"]) == int(lane_nums.max())
    assert np.array_equal(out["radar_adj"].numpy(), gold["radar_adj"])          # PRE_Data adjacency, bit-exact
    assert np.array_equal(out["radar"][0].numpy(), gold["radar"])
    assert np.array_equal(out["velocity"].numpy(), gold["velocity"])
    assert np.array_equal(torch.stack(out["target_point"], 1).numpy(), gold["target_point"])
    assert np.array_equal(torch.stack([torch.stack(w, 1) for w in out["waypoints"]], 1).numpy(), gold["waypoints"])
    assert int(out["fronts"][0].long().sum()) == int(gold["fronts_sum"])
    assert float(out["lidars"][0].double().sum()) == float(gold["lidars_sum"])
    # padded lanes are zero beyond each sample's lane count (pad_sequence semantics)
    for b, n in enumerate(lane_nums.tolist()):
        assert lanes[b, n:].abs().sum() == 0


def test_to_engine_batch_reproduces_the_train_loop_tensors():
    out = data.collate_single_cpu(_samples())
    eb = data.to_engine_batch(out, seq_len=1, pad_lanes_to=128)
    B = 3
    assert eb["rgb_u8"].shape == (B, 3, 256, 256) and eb["rgb_u8"].dtype == torch.uint8
    assert eb["lidar"].shape == (B, 2, 256, 256) and eb["lane"].shape == (B, 128, 10, 5)
    assert eb["lane_num"].dtype == torch.int32 and eb["gt_waypoints"].shape == (B, 4, 2)
    ref = synthetic.synth_batch(B)                                              # the same frames, direct path
    assert torch.equal(eb["rgb_u8"], ref["rgb_u8"]) and torch.equal(eb["lane_num"], ref["lane_num"])
    assert torch.equal(eb["lane"], ref["lane"]) and torch.allclose(eb["radar_adj"], ref["radar_adj"])
    assert torch.allclose(eb["gt_waypoints"], ref["gt_waypoints"]) and torch.allclose(eb["target_point"], ref["target_point"])
    assert torch.allclose(eb["velocity"], ref["velocity"])
    try:
        data.to_engine_batch(out, pad_lanes_to=64)
        assert False
    except ValueError:
        pass


def test_collate_error_behaviour():
    import pytest
    with pytest.raises(RuntimeError, match="equal size"):
        data.collate_single_cpu([{"a": [1, 2]}, {"a": [1]}])
    with pytest.raises(TypeError):
        data.collate_single_cpu([object(), object()])
    assert data.collate_single_cpu([1.5, 2.5]).dtype == torch.float64
    assert data.collate_single_cpu(["a", "b"]) == ["a", "b"]
