#!/usr/bin/env python3
"""Golden vectors for mmfn_b200/preprocess.py from the UNMODIFIED reference dataset class.

Builds a tiny on-disk route in the layout CARLA_Data expects (rgb_front/ maps/ vectormap/ lidar/ radar/ measurements/,
dataloader.py:69-118) from seeded synthetic frames (tests/preprocess_fixture.py), runs the real
CARLA_Data.__getitem__ (dataloader.py:183-268) on it and stores what it returns in tests/golden/preprocess_golden.npz.
Runs only in the build container (the reference is not on the GPU box)."""
import json
import os
import sys
import tempfile
import types

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference/team_code")
sys.modules["torch._six"] = types.SimpleNamespace(string_classes=(str, bytes))

from mmfn_utils.datasets.config import GlobalConfig  # noqa: E402
from mmfn_utils.datasets.dataloader import CARLA_Data, radar_to_size, transform_2d_points  # noqa: E402
import preprocess_fixture as fx  # noqa: E402

out = {}
with tempfile.TemporaryDirectory() as d:
    route = os.path.join(d, "town", "route_00")
    for sub in ("rgb_front", "maps", "vectormap", "lidar", "radar", "measurements"):
        os.makedirs(os.path.join(route, sub))
    for f in range(1, fx.N_FRAMES + 1):
        fr = fx.raw_frame(f)
        name = str(f).zfill(4)
        Image.fromarray(fr["rgb"]).save(os.path.join(route, "rgb_front", name + ".png"))
        Image.fromarray(fr["map"]).save(os.path.join(route, "maps", name + ".png"))
        np.save(os.path.join(route, "vectormap", name + ".npy"), fr["lanes"])
        np.save(os.path.join(route, "lidar", name + ".npy"), fr["points"])
        np.save(os.path.join(route, "radar", name + ".npy"), fr["radar"])
        json.dump(fr["meas"], open(os.path.join(route, "measurements", name + ".json"), "w"))
    cfg = GlobalConfig()
    ds = CARLA_Data([os.path.join(d, "town")], cfg)
    out["n_samples"] = np.array(len(ds))
    for i in range(len(ds)):
        s = ds[i]
        out[f"s{i}_front_sum"] = np.array(s["fronts"][0].numpy().astype(np.int64).sum())
        out[f"s{i}_front_probe"] = s["fronts"][0].numpy()[:, ::37, ::41].copy()
        out[f"s{i}_map_probe"] = s["maps"][0].numpy()[:, ::37, ::41].copy()
        out[f"s{i}_lidar_x5"] = np.round(np.asarray(s["lidars"][0]) * 5).astype(np.uint8)
        out[f"s{i}_vectormap"] = s["vectormaps"][0].numpy()
        out[f"s{i}_radar"] = np.asarray(s["radar"][0])
        out[f"s{i}_waypoints"] = np.asarray(s["waypoints"], dtype=np.float64)
        out[f"s{i}_target_point"] = np.asarray(s["target_point"], dtype=np.float64)
        out[f"s{i}_scalars"] = np.array([s["steer"], s["throttle"], float(s["brake"]), s["command"], s["velocity"]], dtype=np.float64)

# stand-alone functions on harder inputs
rng = np.random.default_rng(5)
pts = rng.normal(0, 30, size=(257, 3))
poses = rng.normal(0, 200, size=(6, 6))
poses[:, 0] = rng.uniform(-4, 4, 6); poses[:, 3] = rng.uniform(-4, 4, 6)
out["tf_points"], out["tf_poses"] = pts, poses
out["tf_out"] = np.stack([transform_2d_points(pts, *p) for p in poses])
for name, n in (("few", 17), ("exact", 81), ("many", 140)):
    r = fx.raw_radar(900 + n, n)
    out[f"radar_in_{name}"] = r
    with np.errstate(divide="ignore", invalid="ignore"):
        out[f"radar_out_{name}"] = np.asarray(radar_to_size(r, (81, 5)))
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "preprocess_golden.npz"), **out)
print("wrote preprocess_golden.npz:", {k: getattr(v, "shape", None) for k, v in list(out.items())[:12]})
