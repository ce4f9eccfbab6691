"""Seeded inputs of the PID-controller golden (tools/make_goldens.py control()) and its test."""
import numpy as np


def control_inputs(n=40, seed=11):
    """Seeded (waypoints (1,4,2), velocity (1,)) sequence for the PID controller golden (shared with the test)."""
    rng = np.random.default_rng(seed)
    seq = []
    for i in range(n):
        wp = np.cumsum(rng.normal(0, 1.0, size=(4, 2)) + np.array([0.1, -1.0]), axis=0).astype(np.float32)
        if i % 9 == 4:
            wp[:] = wp[:1] * 0.01                    # nearly stationary plan -> brake branch
        v = np.float32(0.0 if i % 7 == 3 else rng.uniform(0, 9))
        seq.append((wp[None], np.array([v], dtype=np.float32)))
    return seq
