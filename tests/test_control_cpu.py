"""CPU: MMFN.control_pid (the waypoint -> steer / throttle / brake PID glue the e2e agents call, model_rad.py:697-739)
against a golden sequence produced by the reference method (tools/make_goldens.py control()); the two PID
controllers carry state across calls, so the whole 40-step trajectory is compared."""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from control_fixture import control_inputs                       # noqa: E402
from mmfn_b200.config import GlobalConfig                         # noqa: E402
from mmfn_b200.model_rad import MMFN, PIDController               # noqa: E402


def test_control_pid_sequence_matches_reference(golden_dir):
    gold = np.load(os.path.join(golden_dir, "control_pid_golden.npz"))["rows"]
    cfg = GlobalConfig()
    # control_pid only touches the config and the two controllers: drive the unbound method with a stub instead of
    # constructing the CUDA-only module
    stub = types.SimpleNamespace(
        config=cfg,
        turn_controller=PIDController(cfg.turn_KP, cfg.turn_KI, cfg.turn_KD, cfg.turn_n),
        speed_controller=PIDController(cfg.speed_KP, cfg.speed_KI, cfg.speed_KD, cfg.speed_n))
    rows = []
    for wp, v in control_inputs():
        steer, throttle, brake, meta = MMFN.control_pid(stub, torch.from_numpy(wp.copy()), torch.from_numpy(v.copy()))
        rows.append([float(steer), float(throttle), float(brake), meta["desired_speed"], meta["angle"], meta["delta"],
                     meta["aim"][0], meta["aim"][1], meta["speed"]])
    rows = np.asarray(rows, dtype=np.float64)
    assert rows.shape == gold.shape == (40, 9)
    assert np.array_equal(rows[:, 2], gold[:, 2])                 # brake decisions
    assert gold[:, 2].sum() > 0 and (gold[:, 1] > 0).sum() > 0    # the sequence exercises both branches
    assert np.allclose(rows, gold, rtol=1e-12, atol=1e-12)
