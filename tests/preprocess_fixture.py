"""Seeded raw frames for the preprocessing tests and their golden generator (tools/make_preprocess_goldens.py):
what one CARLA recording step leaves on disk (dataloader.py:69-118)."""
import numpy as np

N_FRAMES = 8          # -> (8 - pred_len - 2) // seq_len = 2 samples


def raw_radar(seed, n):
    rng = np.random.default_rng(seed)
    r = np.zeros((n, 5), dtype=np.float64)
    r[:, 0] = rng.uniform(1, 80, n)
    r[:, 1] = rng.uniform(-0.3, 0.3, n)
    r[:, 2] = rng.uniform(-0.1, 0.1, n)
    r[:, 3] = rng.normal(0, 5, n)
    r[:, 4] = rng.integers(0, 2, n)
    if n > 4:
        r[3, 3] = 0.0          # zero velocity: infinite time-to-collision
    return r


def raw_frame(f):
    rng = np.random.default_rng(4000 + f)
    n_pts = 4096
    pts = np.empty((n_pts, 4), dtype=np.float32)
    pts[:, 0] = rng.uniform(-20, 20, n_pts)
    pts[:, 1] = rng.uniform(-12, 28, n_pts)
    pts[:, 2] = rng.uniform(-4, 2, n_pts)
    pts[:, 3] = rng.uniform(0, 1, n_pts)
    n_lanes = int(rng.integers(5, 12))
    lanes = rng.normal(0, 15, size=(n_lanes, 10, 5))
    theta = float(rng.uniform(-3, 3))
    meas = {"x": float(100 + 3.0 * f + rng.normal(0, 0.1)), "y": float(-50 + 1.5 * f + rng.normal(0, 0.1)),
            "theta": theta if f != 3 else float("nan"),
            "x_command": float(130 + rng.normal(0, 1)), "y_command": float(-35 + rng.normal(0, 1)),
            "steer": 0.01 * f, "throttle": 0.5, "brake": False, "command": 4, "speed": float(rng.uniform(0, 8))}
    return {
        "rgb": rng.integers(0, 256, size=(300, 400, 3), dtype=np.uint8),
        "map": rng.integers(0, 256, size=(256, 256, 3), dtype=np.uint8),
        "points": pts, "lanes": lanes, "radar": raw_radar(7000 + f, int(rng.integers(20, 120))), "meas": meas,
    }
