"""Host-side frame preprocessing of the training data path: what
CARLA_Data.__getitem__ (team_code/mmfn_utils/datasets/dataloader.py:183-268) does to one raw frame before it is
pickled by phase1_preprocess_data.py and later read back by PRE_Data.

  transform_points_2d     dataloader.py:311-334 -- ego-frame change of an (N, 3) point set, float64
  radar_to_size           dataloader.py:336-346 -- pad / prune the radar returns to (81, 5)
  crop_image_chw          dataloader.py:296-308 -- centre crop to channels-first (scale = 1 path)
  local_waypoints         dataloader.py:241-248 -- future ego positions in the current ego frame
  local_command_point     dataloader.py:250-261 -- navigation target in the current ego frame
  frame_to_sample         dataloader.py:201-268 -- the sample dict; the LiDAR histogram comes from a caller-supplied
                          builder: `mmfn_b200.ops.bev_scatter` on the GPU (bit-exact with the reference function for
                          identical points) or the CPU oracle in tests

  ego_pose_params / lidar_frames_to_bev_gpu / write_packed -- the same frame preprocessing with the per-point work on the
                          GPU (csrc/loader.cu: float64 transform; csrc/bev.cu: histogram) and ONE packed shard file
                          (data.PackedShard reads it back) instead of one pickle per frame

The rigid transform is evaluated in closed form (one rotation by r1 - r2 and one translation) instead of the
reference's two 3x3 matrix products and a matrix inverse: mathematically identical, numerically equal to ~1e-13
relative (tests/test_preprocess_cpu.py pins it against outputs of the reference function).  No torch autograd,
no GPU work here: it is loader-side code.
"""
import numpy as np
import torch


def transform_points_2d(xyz, r1, t1_x, t1_y, r2, t2_x, t2_y):
    """Points given in frame 1 (pose: heading r1, offset (t1_x, t1_y)) expressed in frame 2; z is carried over.

    The reference maps p -> M1 p to the world frame, with M1 = [[c1, s1, t1_x], [-s1, c1, t1_y], [0, 0, 1]], and then
    applies inverse(M2).  With R(a) = [[cos a, sin a], [-sin a, cos a]] this is
        out_xy = R(r2)^T (R(r1) p_xy + t1 - t2).
    """
    xyz = np.asarray(xyz, dtype=np.float64)
    c1, s1 = np.cos(r1), np.sin(r1)
    c2, s2 = np.cos(r2), np.sin(r2)
    wx = c1 * xyz[:, 0] + s1 * xyz[:, 1] + (t1_x - t2_x)
    wy = -s1 * xyz[:, 0] + c1 * xyz[:, 1] + (t1_y - t2_y)
    out = np.empty_like(xyz)
    out[:, 0] = c2 * wx - s2 * wy
    out[:, 1] = s2 * wx + c2 * wy
    out[:, 2] = xyz[:, 2]
    return out


def radar_to_size(data, target_size=(81, 5)):
    """Fewer returns than rows: zero padding at the end.  More: drop the returns with the LARGEST |depth / velocity|
    (time to collision; NaN / inf from zero velocity sort last under argsort of the negated value, exactly as in the
    reference) until `target_size[0]` remain, keeping the original order of the survivors."""
    data = np.asarray(data)
    rows = target_size[0]
    if data.shape[0] >= rows:
        n = data.shape[0] - rows
        with np.errstate(divide="ignore", invalid="ignore"):
            drop = (-np.abs(data[:, 0] / data[:, 3])).argsort()[:n]
        return np.delete(data, drop, 0)
    out = np.zeros(target_size)
    out[: data.shape[0], :] = data
    return out


def crop_image_chw(image_hwc, crop=256):
    """Centre crop of an (H, W, 3) uint8 frame, channels first (scale_and_crop_image with scale = 1)."""
    image_hwc = np.asarray(image_hwc)
    h, w = image_hwc.shape[:2]
    y0, x0 = h // 2 - crop // 2, w // 2 - crop // 2
    return np.ascontiguousarray(np.transpose(image_hwc[y0:y0 + crop, x0:x0 + crop], (2, 0, 1)))


def local_waypoints(xs, ys, thetas, ego_index):
    """Ego positions of the frames in `xs, ys, thetas` (world) expressed in the frame of `ego_index`: the reference
    transforms the ORIGIN of every frame i into the ego frame (dataloader.py:241-247)."""
    ex, ey, et = xs[ego_index], ys[ego_index], thetas[ego_index]
    wps = []
    for x, y, th in zip(xs, ys, thetas):
        p = transform_points_2d(np.zeros((1, 3)), np.pi / 2 - th, -x, -y, np.pi / 2 - et, -ex, -ey)
        wps.append((p[0, 0], p[0, 1]))
    return wps


def local_command_point(x_command, y_command, ego_x, ego_y, ego_theta):
    """Navigation target in the ego frame: R(pi/2 + theta)^T (target - ego) (dataloader.py:250-261)."""
    a = np.pi / 2 + ego_theta
    c, s = np.cos(a), np.sin(a)
    dx, dy = x_command - ego_x, y_command - ego_y
    return (c * dx + s * dy, -s * dx + c * dy)


def frame_to_sample(rgb_hwc, points_xyzi, lanes, radar_raw, map_hwc, xs, ys, thetas, x_command, y_command,
                    bev_fn, seq_len=1, steer=0.0, throttle=0.0, brake=False, command=4, velocity=0.0, crop=256):
    """One training sample in the layout the phase-1 pickles hold (seq_len = 1, the MMFN configuration).
    xs / ys / thetas: ego poses of the current frame followed by the `pred_len` future frames.
    bev_fn(points (N, 3) float32) -> (2, 256, 256) float32 histogram."""
    assert seq_len == 1, "MMFN is trained with seq_len = 1 (config.py:6)"
    thetas = [0.0 if np.isnan(t) else float(t) for t in thetas]              # dataloader.py:222-224
    i = seq_len - 1
    pts = np.array(np.asarray(points_xyzi)[..., :3], dtype=np.float64)
    pts[:, 1] *= -1                                                          # dataloader.py:232
    pts = transform_points_2d(pts, np.pi / 2 - thetas[i], -xs[i], -ys[i], np.pi / 2 - thetas[i], -xs[i], -ys[i])
    lidar = np.asarray(bev_fn(pts.astype(np.float32)), dtype=np.float32)
    return {
        "fronts": [torch.from_numpy(crop_image_chw(rgb_hwc, crop))],
        "lidars": [lidar],
        "vectormaps": [torch.from_numpy(np.asarray(lanes))],
        "radar": [radar_to_size(radar_raw, (81, 5))],
        "maps": [torch.from_numpy(np.ascontiguousarray(np.transpose(np.asarray(map_hwc), (2, 0, 1))))],
        "waypoints": local_waypoints(xs, ys, thetas, i),
        "target_point": local_command_point(x_command, y_command, xs[i], ys[i], thetas[i]),
        "steer": steer, "throttle": throttle, "brake": brake, "command": command, "velocity": velocity,
    }


# --------------------------------------------------------------------------- recorded route -> pickles (phase 1)
def route_sequences(route_dir, seq_len=1, pred_len=4):
    """Frame numbers of the usable samples of one recorded route, as CARLA_Data.__init__ enumerates them
    (dataloader.py:69-80): the first frame and the last pred_len + 1 frames are not used as `current` frames."""
    import os
    n = len(os.listdir(os.path.join(route_dir, "rgb_front")))
    num_seq = (n - pred_len - 2) // seq_len
    return [seq * seq_len + 1 for seq in range(num_seq)]


def load_route_sample(route_dir, frame, bev_fn, pred_len=4, crop=256):
    """Read one recorded frame (+ the poses of its pred_len successors) from the CARLA recording layout
    rgb_front/ maps/ vectormap/ lidar/ radar/ measurements/ (dataloader.py:83-118) and preprocess it."""
    import json
    import os
    from PIL import Image

    def name(f, ext):
        return f"{str(f).zfill(4)}.{ext}"

    meas = []
    for f in range(frame, frame + 1 + pred_len):
        with open(os.path.join(route_dir, "measurements", name(f, "json"))) as fd:
            meas.append(json.load(fd))
    cur = meas[0]
    vm = os.path.join(route_dir, "vectormap", name(frame, "npy"))
    if not os.path.exists(vm):
        raise FileNotFoundError(f"{vm}: the reference falls back to a neighbouring sample's lanes here "
                                "(dataloader.py:206-214); this loader reports the missing file instead")
    return frame_to_sample(
        np.asarray(Image.open(os.path.join(route_dir, "rgb_front", name(frame, "png")))),
        np.load(os.path.join(route_dir, "lidar", name(frame, "npy"))),
        np.load(vm),
        np.load(os.path.join(route_dir, "radar", name(frame, "npy"))),
        np.asarray(Image.open(os.path.join(route_dir, "maps", name(frame, "png")))),
        [m["x"] for m in meas], [m["y"] for m in meas], [m["theta"] for m in meas],
        cur["x_command"], cur["y_command"], bev_fn,
        steer=cur["steer"], throttle=cur["throttle"], brake=cur["brake"], command=cur["command"],
        velocity=cur["speed"], crop=crop)


def write_pickles(samples, out_dir):
    """phase1_preprocess_data.py:41-43: one `<index>.pkl` per sample, the files PRE_Data (data.py) reads back."""
    import os
    import pickle
    os.makedirs(out_dir, exist_ok=True)
    n = 0
    for n, s in enumerate(samples, start=1):
        with open(os.path.join(out_dir, f"{n - 1}.pkl"), "wb") as fd:
            pickle.dump(s, fd)
    return n


# --------------------------------------------------------------------------- GPU phase 1 + packed shard
def ego_pose_params(r1, t1_x, t1_y, r2, t2_x, t2_y):
    """The six float64 numbers mmfn_lidar_ego_transform_f64 takes per frame, evaluated with the numpy calls of
    transform_points_2d so that the device result is bit-identical to it."""
    return np.array([np.cos(r1), np.sin(r1), np.cos(r2), np.sin(r2), t1_x - t2_x, t1_y - t2_y], dtype=np.float64)


def lidar_frames_to_bev_gpu(sweeps, xs, ys, thetas, device="cuda"):
    """Raw sweeps (list of (N_i, >= 3) float32 arrays) + ego poses -> uint8 histogram counts (F, 2, 256, 256) on the
    device: y flip and ego transform in float64 (dataloader.py:229-239), histogram (:271-293), all frames in three
    launches.  Ragged sweeps are padded with points far outside the grid."""
    from . import ops
    F_ = len(sweeps)
    n = max(int(s.shape[0]) for s in sweeps)
    n = max(4, (n + 3) // 4 * 4)
    pts = np.full((F_, n, 3), 1.0e6, dtype=np.float32)             # padding: lands outside every bin
    pose = np.empty((F_, 6), dtype=np.float64)
    for i, s in enumerate(sweeps):
        pts[i, : s.shape[0]] = np.asarray(s)[:, :3]
        th = 0.0 if np.isnan(thetas[i]) else float(thetas[i])
        pose[i] = ego_pose_params(np.pi / 2 - th, -xs[i], -ys[i], np.pi / 2 - th, -xs[i], -ys[i])
    ego = ops.lidar_ego_transform(torch.from_numpy(pts).to(device), torch.from_numpy(pose).to(device))
    return ops.bev_pack_u8(ops.bev_scatter(ego))


PACK_MAGIC = b"MMFNPK01"
PACK_ALIGN = 4096


def samples_to_arrays(samples, lidar_counts=None):
    """Samples in the phase-1 pickle layout (frame_to_sample / synthetic.synth_sample) -> the arrays of a packed shard.
    lidar_counts (F, 2, 256, 256) uint8: histogram counts from lidar_frames_to_bev_gpu; when None they are recovered from
    the float32 histograms the samples hold (value * 5 is an integer 0..5)."""
    F_ = len(samples)
    lanes = [np.asarray(s["vectormaps"][0]) for s in samples]
    lmax = max(l.shape[0] for l in lanes)
    lane = np.zeros((F_, lmax) + lanes[0].shape[1:], dtype=np.float32)
    for i, l in enumerate(lanes):
        lane[i, : l.shape[0]] = l                                   # float64 -> float32: Engine.train's cast, done once
    radar64 = np.stack([np.asarray(s["radar"][0], dtype=np.float64) for s in samples])
    if lidar_counts is None:
        lidar_counts = np.stack([np.rint(np.asarray(s["lidars"][0], dtype=np.float32) * 5.0).astype(np.uint8) for s in samples])
    arrays = {
        "fronts": np.stack([np.asarray(s["fronts"][0], dtype=np.uint8) for s in samples]),
        "maps": np.stack([np.asarray(s["maps"][0], dtype=np.uint8) for s in samples]),
        "lidar_u8": np.ascontiguousarray(np.asarray(lidar_counts, dtype=np.uint8)),
        "lane": lane,
        "lane_num": np.array([l.shape[0] for l in lanes], dtype=np.int32),
        "radar": radar64.astype(np.float32),
        "radar_az64": np.ascontiguousarray(radar64[:, :, 1]),
        "waypoints": np.array([s["waypoints"] for s in samples], dtype=np.float64),
        "target_point": np.array([s["target_point"] for s in samples], dtype=np.float64),
        "velocity": np.array([s["velocity"] for s in samples], dtype=np.float64),
        "steer": np.array([s["steer"] for s in samples], dtype=np.float64),
        "throttle": np.array([s["throttle"] for s in samples], dtype=np.float64),
        "brake": np.array([bool(s["brake"]) for s in samples], dtype=np.uint8),
        "command": np.array([s["command"] for s in samples], dtype=np.int64),
    }
    return arrays


def write_packed(samples, path, lidar_counts=None):
    """ONE shard file instead of len(samples) pickles: magic, a JSON table {name: dtype, shape, offset}, then the raw
    arrays at 4096-byte boundaries (memory-mappable, no unpickling on the training side).  Returns the sample count."""
    import json
    arrays = samples_to_arrays(samples, lidar_counts)
    table, off = {}, 0
    for name, a in arrays.items():
        table[name] = {"dtype": a.dtype.str, "shape": list(a.shape), "offset": off}
        off += (a.nbytes + PACK_ALIGN - 1) // PACK_ALIGN * PACK_ALIGN
    head = json.dumps({"n": len(samples), "arrays": table}).encode()
    data0 = (len(PACK_MAGIC) + 8 + len(head) + PACK_ALIGN - 1) // PACK_ALIGN * PACK_ALIGN
    with open(path, "wb") as fd:
        fd.write(PACK_MAGIC)
        fd.write(np.uint64(len(head)).tobytes())
        fd.write(head)
        for name, a in arrays.items():
            fd.seek(data0 + table[name]["offset"])
            fd.write(np.ascontiguousarray(a).tobytes())
        fd.truncate(data0 + off)
    return len(samples)
