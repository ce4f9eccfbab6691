"""CPU: mmfn_b200/preprocess.py (frame -> training sample) against goldens produced by the
UNMODIFIED reference dataset class CARLA_Data on the same seeded frames (tools/make_preprocess_goldens.py)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import preprocess_fixture as fx                      # noqa: E402
from mmfn_b200 import preprocess as pp               # noqa: E402
from oracle import bev_oracle                        # noqa: E402

PRED_LEN = 4


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "preprocess_golden.npz"))


def test_frame_to_sample_matches_reference_dataset(gold):
    assert int(gold["n_samples"]) == fx.N_FRAMES - PRED_LEN - 2
    for i in range(int(gold["n_samples"])):
        frames = [fx.raw_frame(i + 1 + k) for k in range(1 + PRED_LEN)]      # current frame + pred_len future frames
        cur = frames[0]
        m = cur["meas"]
        s = pp.frame_to_sample(cur["rgb"], cur["points"], cur["lanes"], cur["radar"], cur["map"],
                               [f["meas"]["x"] for f in frames], [f["meas"]["y"] for f in frames],
                               [f["meas"]["theta"] for f in frames], m["x_command"], m["y_command"],
                               bev_fn=bev_oracle.lidar_to_histogram_features,
                               steer=m["steer"], throttle=m["throttle"], brake=m["brake"], command=m["command"],
                               velocity=m["speed"])
        front = s["fronts"][0].numpy()
        assert front.shape == (3, 256, 256) and front.dtype == np.uint8
        assert int(front.astype(np.int64).sum()) == int(gold[f"s{i}_front_sum"])
        assert np.array_equal(front[:, ::37, ::41], gold[f"s{i}_front_probe"])
        assert np.array_equal(s["maps"][0].numpy()[:, ::37, ::41], gold[f"s{i}_map_probe"])
        assert np.array_equal(s["vectormaps"][0].numpy(), gold[f"s{i}_vectormap"])
        assert np.array_equal(s["radar"][0], gold[f"s{i}_radar"])             # same ops, same order: bit-exact
        assert np.allclose(np.asarray(s["waypoints"]), gold[f"s{i}_waypoints"], rtol=0, atol=1e-9)
        assert np.allclose(np.asarray(s["target_point"]), gold[f"s{i}_target_point"], rtol=0, atol=1e-9)
        assert np.array_equal(np.round(s["lidars"][0] * 5).astype(np.uint8), gold[f"s{i}_lidar_x5"])
        assert [s["steer"], s["throttle"], float(s["brake"]), s["command"], s["velocity"]] == list(gold[f"s{i}_scalars"])
        assert s["waypoints"][0] == (0.0, 0.0) or np.allclose(s["waypoints"][0], 0.0, atol=1e-9)   # ego origin


def test_transform_points_2d_matches_reference_function(gold):
    pts, poses, ref = gold["tf_points"], gold["tf_poses"], gold["tf_out"]
    for p, r in zip(poses, ref):
        out = pp.transform_points_2d(pts, *p)
        assert out.dtype == np.float64
        assert np.allclose(out, r, rtol=0, atol=1e-9)                         # coordinates up to ~1e3: ~1e-13 relative
        assert np.array_equal(out[:, 2], pts[:, 2])                           # z is carried over untouched
    # a frame expressed in itself is the identity
    same = pp.transform_points_2d(pts, 0.3, 5.0, -7.0, 0.3, 5.0, -7.0)
    assert np.allclose(same, pts, rtol=0, atol=1e-12)


def test_radar_to_size_matches_reference_function(gold):
    for name in ("few", "exact", "many"):
        out = pp.radar_to_size(gold[f"radar_in_{name}"], (81, 5))
        assert out.shape == (81, 5)
        assert np.array_equal(out, gold[f"radar_out_{name}"])
    assert np.array_equal(pp.radar_to_size(np.zeros((0, 5)), (81, 5)), np.zeros((81, 5)))


def test_recorded_route_to_pickles_to_engine_batch(gold, tmp_path):
    """raw recording on disk -> preprocess -> <i>.pkl -> PRE_Data -> collate -> engine batch: the whole loader side."""
    import json
    import torch
    from PIL import Image
    from mmfn_b200 import data as mdata
    from mmfn_b200.config import GlobalConfig
    route = tmp_path / "town" / "route_00"
    for sub in ("rgb_front", "maps", "vectormap", "lidar", "radar", "measurements"):
        (route / sub).mkdir(parents=True)
    for f in range(1, fx.N_FRAMES + 1):
        fr = fx.raw_frame(f)
        name = str(f).zfill(4)
        Image.fromarray(fr["rgb"]).save(route / "rgb_front" / (name + ".png"))
        Image.fromarray(fr["map"]).save(route / "maps" / (name + ".png"))
        np.save(route / "vectormap" / (name + ".npy"), fr["lanes"])
        np.save(route / "lidar" / (name + ".npy"), fr["points"])
        np.save(route / "radar" / (name + ".npy"), fr["radar"])
        json.dump(fr["meas"], open(route / "measurements" / (name + ".json"), "w"))
    frames = pp.route_sequences(str(route))
    assert frames == [1, 2]
    out = tmp_path / "pro_train"
    n = pp.write_pickles((pp.load_route_sample(str(route), f, bev_oracle.lidar_to_histogram_features) for f in frames), str(out))
    assert n == 2 and sorted(p.name for p in out.iterdir()) == ["0.pkl", "1.pkl"]
    ds = mdata.PRE_Data(str(out), GlobalConfig())
    assert len(ds) == 2
    samples = sorted((ds[i] for i in range(2)), key=lambda s: s["steer"])
    for i, s in enumerate(samples):
        assert np.array_equal(np.round(np.asarray(s["lidars"][0]) * 5).astype(np.uint8), gold[f"s{i}_lidar_x5"])
        assert np.array_equal(s["radar"][0], gold[f"s{i}_radar"])
        assert s["radar_adj"].shape == (81, 81)
        assert np.allclose(np.asarray(s["waypoints"]), gold[f"s{i}_waypoints"], atol=1e-9)
    batch = mdata.collate_single_cpu(samples)
    lanes, lane_num, lmax = batch["vectormaps"][0]
    assert lanes.shape[0] == 2 and lanes.shape[1] == int(lane_num.max()) == lmax
    eb = mdata.to_engine_batch(batch)
    assert eb["rgb_u8"].shape == (2, 3, 256, 256) and eb["rgb_u8"].dtype == torch.uint8
    assert eb["gt_waypoints"].shape == (2, 4, 2) and eb["radar"].shape == (2, 81, 5)
