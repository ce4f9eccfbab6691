"""Epoch-level driver around TrainEngine: the B200-native counterpart of the reference `Engine` class
(run_steps/phase2_train_net.py:32-220) and of main()'s resume logic (:288-302).

  Trainer.train(loader)      one epoch of optimisation steps (Engine.train :54-122)
  Trainer.validate(loader)   eval-mode mean waypoint L1 (Engine.validate :124-183)
  Trainer.save(logdir)       recent.log / model.pth / recent_optim.pth and, on a new best validation loss,
                             best_model.pth / best_optim.pth (Engine.save :185-220) -- same file names, same JSON
                             keys, state_dicts loadable by the reference model / torch.optim.AdamW
  Trainer.resume(logdir)     the reference's resume branch (:288-302)

Batches come from a DataLoader built on data.PRE_Data + data.collate_single_cpu (or any iterable of collated
reference batches), or from data.PackedLoader (packed shards: already in engine form, histogram as uint8 counts and radar
azimuths instead of the adjacency matrix -- expanded on the GPU); collated batches are flattened with data.to_engine_batch, packed into one pinned buffer and moved with a
single H2D copy per step (engine.BatchStager).
"""
import json
import os

import torch

from . import data as data_mod
from . import ops
from .engine import BatchStager, TrainEngine, batch_inputs
from .params import is_unused


class Trainer:
    def __init__(self, model, lr=1e-4, logdir=None, pad_lanes_to=None, process_group=None):
        self.model = model
        self.engine = TrainEngine(model, lr=lr, process_group=process_group)
        self.logdir = logdir
        self.pad_lanes_to = pad_lanes_to
        self.cur_epoch, self.cur_iter = 0, 0
        self.bestval, self.bestval_epoch = 1e10, 0
        self.train_loss, self.val_loss = [], []
        self._stager = None

    # ---- batch plumbing ------------------------------------------------------------------------
    def _device_batch(self, data):
        b = data if "rgb_u8" in data else data_mod.to_engine_batch(data, self.model.config.seq_len, self.pad_lanes_to)
        if self.model.VARIANT == "img" and "map_u8" not in b:
            b["map_u8"] = data["maps"][0].to(torch.uint8)
        shapes = {k: tuple(v.shape) for k, v in b.items()}
        if self._stager is None or self._stager_shapes != shapes:        # lane count may change between batches
            self._stager, self._stager_shapes = BatchStager(b, self.model.device), shapes
        return self._stager.stage(b)

    # ---- Engine.train ----------------------------------------------------------------------------
    def train(self, loader):
        total, n = 0.0, 0
        for data in loader:
            loss = self.engine.step(self._device_batch(data))
            total += float(loss)                    # the reference reads loss.item() every step (:109)
            n += 1
            self.cur_iter += 1
        if n:
            self.train_loss.append(total / n)
        self.cur_epoch += 1
        return self.train_loss[-1] if n else None

    # ---- Engine.validate -------------------------------------------------------------------------
    @torch.no_grad()
    def validate(self, loader):
        model = self.model
        model.eval()
        total, n = 0.0, 0
        for data in loader:
            b = self._device_batch(data)
            lidar, radar_adj = batch_inputs(b)                     # histogram / adjacency from any batch form (packed too)
            lane = b["map_u8"] if model.VARIANT == "img" else b.get("lane")
            pred = model.net.forward(b["rgb_u8"], lidar, lane, b.get("lane_num"), b.get("radar"), radar_adj,
                                     b["target_point"], b["velocity"], model.seed, False)
            loss, _ = ops.l1_loss(pred, b["gt_waypoints"], want_grad=False)
            total += float(loss)
            n += 1
        model.train()
        if n:
            self.val_loss.append(total / n)
            return self.val_loss[-1]
        return None

    # ---- Engine.save / resume --------------------------------------------------------------------
    def log_table(self):
        return {"epoch": self.cur_epoch, "iter": self.cur_iter, "bestval": self.bestval,
                "bestval_epoch": self.bestval_epoch, "train_loss": self.train_loss, "val_loss": self.val_loss}

    def save(self, logdir=None):
        logdir = logdir or self.logdir
        os.makedirs(logdir, exist_ok=True)
        save_best = False
        if self.val_loss and self.val_loss[-1] <= self.bestval:
            self.bestval, self.bestval_epoch, save_best = self.val_loss[-1], self.cur_epoch, True
        msd = {k: v.detach().cpu().contiguous() for k, v in self.model.state_dict().items()}
        osd = optimizer_state_dict(self.engine)
        if save_best:
            torch.save(msd, os.path.join(logdir, "best_model.pth"))
            torch.save(osd, os.path.join(logdir, "best_optim.pth"))
        torch.save(msd, os.path.join(logdir, "model.pth"))
        torch.save(osd, os.path.join(logdir, "recent_optim.pth"))
        with open(os.path.join(logdir, "recent.log"), "w") as f:
            f.write(json.dumps(self.log_table()))
        return save_best

    def resume(self, logdir=None):
        logdir = logdir or self.logdir
        path = os.path.join(logdir, "recent.log")
        if not os.path.isfile(path):
            return False
        with open(path) as f:
            t = json.load(f)
        self.cur_epoch, self.cur_iter = t["epoch"], t.get("iter", 0)
        self.bestval, self.train_loss, self.val_loss = t["bestval"], t["train_loss"], t["val_loss"]
        self.bestval_epoch = t.get("bestval_epoch", 0)
        # the reference reloads best_* (phase2_train_net.py:288-302); a run saved before its first validation has only
        # model.pth / recent_optim.pth -- fall back to those instead of raising
        def pick(best, recent):
            p = os.path.join(logdir, best)
            return p if os.path.isfile(p) else os.path.join(logdir, recent)
        load_optimizer_state_dict(self.engine, torch.load(pick("best_optim.pth", "recent_optim.pth"), map_location="cpu"))
        self.model.load_state_dict(torch.load(pick("best_model.pth", "model.pth"), map_location="cpu"))
        return True


# ---- torch.optim.AdamW-compatible optimizer state ------------------------------------------------
def _param_keys(model):
    return [k for k, _ in model._param_items]                 # model.parameters() order == reference order


def _opt_view(engine, flat, key):
    """m / v of one parameter, in the shape torch.optim would hold (conv filters (K,C,R,S))."""
    st = engine.st
    shape, off = st.shapes[key], st.offsets[key]
    n = 1
    for s in shape:
        n *= s
    t = flat[off: off + n]
    if st.kinds[key] == "conv":
        co, c, r, s = shape
        return t.view(co, r, s, c).permute(0, 3, 1, 2)
    return t.view(shape)


def optimizer_state_dict(engine):
    """What torch.optim.AdamW(model.parameters(), lr).state_dict() holds after the same steps
    (phase2_train_net.py:256; saved at :213,:219): per-parameter step / exp_avg / exp_avg_sq, indexed by the
    position in model.parameters(); parameters that never received a gradient have no entry."""
    keys = _param_keys(engine.model)
    step = float(engine.state[0].item())
    state = {}
    if step > 0:
        for i, k in enumerate(keys):
            if is_unused(k, engine.model.VARIANT):
                continue
            state[i] = {"step": torch.tensor(step), "exp_avg": _opt_view(engine, engine.m, k).detach().cpu().contiguous(),
                        "exp_avg_sq": _opt_view(engine, engine.v, k).detach().cpu().contiguous()}
    group = {"lr": engine.lr, "betas": tuple(engine.betas), "eps": engine.eps, "weight_decay": engine.wd,
             "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
             "fused": None, "params": list(range(len(keys)))}
    return {"state": state, "param_groups": [group]}


def load_optimizer_state_dict(engine, sd):
    keys = _param_keys(engine.model)
    g = sd["param_groups"][0]
    if len(g["params"]) != len(keys):
        raise ValueError(f"optimizer state has {len(g['params'])} parameters, the model has {len(keys)}")
    engine.lr, engine.betas, engine.eps, engine.wd = g["lr"], tuple(g["betas"]), g["eps"], g["weight_decay"]
    engine.m.zero_()
    engine.v.zero_()
    step = 0.0
    for i, stt in sd["state"].items():
        k = keys[int(i)]
        if is_unused(k, engine.model.VARIANT):
            continue
        _opt_view(engine, engine.m, k).copy_(stt["exp_avg"])
        _opt_view(engine, engine.v, k).copy_(stt["exp_avg_sq"])
        step = max(step, float(stt["step"]))
    # same arithmetic as the device-side advance (head.cu: fp32 betas widened to double)
    b1, b2 = (float(torch.tensor(b, dtype=torch.float32)) for b in engine.betas)
    engine.state.copy_(torch.tensor([step, 1.0 - b1 ** step, 1.0 - b2 ** step], dtype=torch.float32))
