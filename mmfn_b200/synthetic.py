"""Seeded synthetic frames and weights.  There is no dataset or
checkpoint in the sandbox, so benchmarks, tests and goldens all draw from here; everything is
generated on the CPU with numpy / torch CPU generators so it is identical on every machine.

One frame = 400x300 RGB + 32 768-point LiDAR sweep + lane set + 81 radar points, i.e. what
CARLA_Data.__getitem__ (team_code/mmfn_utils/datasets/dataloader.py:183-268) reads from disk.
"""
import zlib

import numpy as np
import torch

N_POINTS = 32768


def synth_points(seed, n=N_POINTS):
    """(n, 4) float32 XYZI sweep: ~36 % of points fall outside the BEV grid, 1 % sit on bin edges or
    their float neighbours, plus NaN/inf/denormal rows and one saturated pillar."""
    rng = np.random.default_rng(seed)
    p = np.empty((n, 4), dtype=np.float32)
    if n == 0:
        return p
    p[:, 0] = rng.uniform(-20, 20, n)
    p[:, 1] = rng.uniform(-28, 12, n)
    p[:, 2] = rng.uniform(-4, 2, n)
    p[:, 3] = rng.uniform(0, 1, n)
    k = max(1, n // 100)
    sel = rng.choice(n, size=k, replace=False)
    snapped = (np.round(p[sel, :2] * 8) / 8).astype(np.float32)
    mode = rng.integers(0, 3, size=(k, 2))
    snapped = np.where(mode == 1, np.nextafter(snapped, np.float32(np.inf)), snapped)
    snapped = np.where(mode == 2, np.nextafter(snapped, np.float32(-np.inf)), snapped)
    p[sel, :2] = snapped.astype(np.float32)
    specials = np.array([[16.0, 8.0, -2.0, 0], [-16.0, -24.0, -2.0, 0], [16.0, -24.0, 0.0, 0],
                         [np.nan, 0, 0, 0], [0, np.nan, 0, 0], [0, 0, np.nan, 0],
                         [np.inf, 0, 0, 0], [0, -np.inf, 0, 0], [1e-40, -1e-40, -2.0000002, 0],
                         [15.999999, 7.9999995, 5.0, 0]], dtype=np.float32)
    m = min(len(specials), n)
    p[:m] = specials[:m]
    if n >= 2048:
        p[100:600, :3] = np.array([3.3, -3.3, 0.5], dtype=np.float32)
    return p


def synth_frame(index, n_lanes=128, n_nodes=10, n_points=N_POINTS):
    """Raw sensor frame for sample `index` (seed = 1234 + index)."""
    rng = np.random.default_rng(1234 + index)
    rgb = rng.integers(0, 256, size=(300, 400, 3), dtype=np.uint8)
    points = synth_points(1234 + index, n_points)
    lane = np.zeros((n_lanes, n_nodes, 5), dtype=np.float32)
    lane[:, :, 0:2] = rng.normal(0, 15, size=(n_lanes, n_nodes, 2))
    lane[:, :, 2:5] = rng.integers(0, 3, size=(n_lanes, n_nodes, 3))
    lane_num = int(rng.integers(n_lanes // 2, n_lanes + 1))
    lane[lane_num:] = 0
    radar = np.zeros((81, 5), dtype=np.float32)
    m = int(rng.integers(20, 82))
    radar[:m, 0] = rng.uniform(1, 80, m)
    radar[:m, 1] = rng.uniform(-0.3, 0.3, m)
    radar[:m, 2] = rng.uniform(-0.1, 0.1, m)
    radar[:m, 3] = rng.normal(0, 5, m)
    radar[:m, 4] = rng.integers(0, 2, m)
    velocity = np.float32(rng.uniform(0, 10))
    target_point = (rng.normal(0, 10, 2) * 2).astype(np.float32)
    gt = np.cumsum(rng.normal(0, 1, size=(4, 2)), axis=0).astype(np.float32)
    return dict(rgb=rgb, points=points, lane=lane, lane_num=lane_num, radar=radar,
                velocity=velocity, target_point=target_point, gt_waypoints=gt)


def synth_sample(index, bev_fn, n_lanes=128, n_nodes=10, n_points=N_POINTS):
    """One sample in the dict layout CARLA_Data.__getitem__ / the phase-1 pickles use
    (dataloader.py:183-268): per-timestep lists, tuples for waypoints and target point, python floats.
    `bev_fn(points_xyz) -> (2,256,256)` supplies the LiDAR histogram the reference stores in the pickle."""
    f = synth_frame(index, n_lanes, n_nodes, n_points)
    wps = [(0.0, 0.0)] + [(float(x), float(y)) for x, y in f["gt_waypoints"].astype(np.float64)]
    return {
        "fronts": [torch.from_numpy(center_crop_chw(f["rgb"]))],
        "lidars": [np.asarray(bev_fn(f["points"][:, :3]), dtype=np.float32)],
        "vectormaps": [torch.from_numpy(f["lane"][: f["lane_num"]].astype(np.float64))],
        "radar": [f["radar"].astype(np.float64)],
        "maps": [torch.from_numpy(center_crop_chw(f["rgb"][::-1].copy()))],
        "waypoints": wps,
        "target_point": (float(f["target_point"][0]), float(f["target_point"][1])),
        "steer": 0.1 * index, "throttle": 0.5, "brake": False, "command": 4,
        "velocity": float(f["velocity"]),
    }


def center_crop_chw(rgb, crop=256):
    """scale_and_crop_image with scale=1 (dataloader.py:296-308): HWC uint8 -> CHW uint8 crop."""
    h, w = rgb.shape[:2]
    sx, sy = h // 2 - crop // 2, w // 2 - crop // 2
    return np.ascontiguousarray(np.transpose(rgb[sx:sx + crop, sy:sy + crop], (2, 0, 1)))


def radar_adjacency(radar):
    """PRE_Data.__getitem__ (dataloader.py:379-384): adj[i, j] = radar[j, 1] - radar[i, 1]."""
    az = radar[..., 1]
    return az[..., None, :] - az[..., :, None]


def synth_batch(batch_size, first_index=0, n_lanes=128, n_nodes=10, n_points=N_POINTS):
    """Host-side batch in the layout Engine.train builds (phase2_train_net.py:66-103), except that
    LiDAR stays a raw point cloud: the BEV histogram is built on the GPU (north_star)."""
    fr = [synth_frame(first_index + i, n_lanes, n_nodes, n_points) for i in range(batch_size)]
    return dict(
        rgb_u8=torch.from_numpy(np.stack([center_crop_chw(f["rgb"]) for f in fr])),        # (B,3,256,256) u8
        points=torch.from_numpy(np.stack([f["points"] for f in fr])),                       # (B,N,4) f32
        lane=torch.from_numpy(np.stack([f["lane"] for f in fr])),                           # (B,L,P,5) f32
        lane_num=torch.tensor([f["lane_num"] for f in fr], dtype=torch.int32),              # (B,)
        radar=torch.from_numpy(np.stack([f["radar"] for f in fr])),                         # (B,81,5)
        radar_adj=torch.from_numpy(np.stack([radar_adjacency(f["radar"]) for f in fr])),    # (B,81,81)
        velocity=torch.from_numpy(np.stack([f["velocity"] for f in fr])),                   # (B,)
        target_point=torch.from_numpy(np.stack([f["target_point"] for f in fr])),           # (B,2)
        gt_waypoints=torch.from_numpy(np.stack([f["gt_waypoints"] for f in fr])),           # (B,4,2)
    )


def synth_map_images(batch_size, first_index=0):
    """(B,3,256,256) uint8 rasterised-map crops for the model_img variant (dataloader.py:214-216 stores the
    `maps` image next to `fronts`): a second seeded image per frame."""
    out = []
    for i in range(batch_size):
        rng = np.random.default_rng(777000 + first_index + i)
        out.append(center_crop_chw(rng.integers(0, 256, size=(300, 400, 3), dtype=np.uint8)))
    return torch.from_numpy(np.stack(out))


def fill_golden_weights(state_dict, seed=42):
    """Deterministic, key-addressed weights (independent of construction order) so the reference
    module, the oracle and the CUDA module can be loaded with bit-identical parameters."""
    out = {}
    for k in sorted(state_dict.keys()):
        v = state_dict[k]
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(k.encode())) % (2 ** 31))
        shape = tuple(v.shape)
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros(shape, dtype=torch.int64)
            continue
        r = torch.randn(shape, generator=g, dtype=torch.float32)
        leaf = k.split(".")[-1]
        if k.endswith("running_mean"):
            t = 0.1 * r
        elif k.endswith("running_var"):
            t = 1.0 + 0.2 * torch.rand(shape, generator=g) - 0.1
        elif len(shape) == 4:                                   # conv filters, kaiming fan_out
            t = r * (2.0 / (shape[0] * shape[2] * shape[3])) ** 0.5
        elif leaf == "pos_emb":
            t = 0.02 * r
        elif leaf in ("W", "a"):                                # GAT, xavier_normal(gain=1.414)
            t = r * 1.414 * (2.0 / (shape[0] + shape[1])) ** 0.5
        elif len(shape) == 2:
            if ".transformer" in k:
                t = 0.02 * r
            else:
                t = (torch.rand(shape, generator=g) * 2 - 1) / shape[1] ** 0.5
        elif len(shape) == 1:
            is_norm_w = leaf == "weight" and (".bn" in k or "downsample.1" in k or ".ln" in k or k.endswith("ln_f.weight")
                                              or _is_layernorm_weight(k))
            t = 1.0 + 0.1 * r if is_norm_w else 0.02 * r
        else:
            t = 0.02 * r
        out[k] = t.contiguous()
    return out


def _is_layernorm_weight(k):
    # LayerNorms inside nn.Sequential containers of the VectorNet encoder (index 1 of each MLP)
    return any(s in k for s in ("mlp.1.weight", "pos_emb.1.weight", "agent_fusion.1.weight", "generator.1.weight"))
