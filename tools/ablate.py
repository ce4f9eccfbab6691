#!/usr/bin/env python3
"""GPU developer aid: where does the captured training step spend its time?

Captures sub-schedules of the step (each fusion transformer alone, each trunk alone, VectorNet, the radar
GAT, forward only, the whole forward+backward) into their own CUDA graphs and times graph replays.  The
isolated times bound each component's share of the step's critical path (GPT_i are strictly sequential with
the trunks; the three trunks run concurrently between fusion points)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmfn_b200 import ops, synthetic                      # noqa: E402
from mmfn_b200.config import GlobalConfig                 # noqa: E402
from mmfn_b200.engine import BatchStager, TrainEngine     # noqa: E402
from mmfn_b200.model_rad import MMFN, _Aux                # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ops.set_precision(sys.argv[2] if len(sys.argv) > 2 else "tf32")      # "tf32" (configs[1]) or "bf16" (configs[2])
dev = torch.device("cuda:0")
model = MMFN(GlobalConfig(), dev)
eng = TrainEngine(model)
net = model.net
hb = synthetic.synth_batch(B)
st = BatchStager(hb, dev)
st.stage(hb)
torch.cuda.synchronize()
db = st.dev_views
results = {}


def time_graph(name, fn, reps=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            fn()
            _Aux.join_all()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    l0 = ops.lib().launches
    with torch.cuda.graph(g):
        fn()
        _Aux.join_all()
    n = ops.lib().launches - l0
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    results[name] = dict(ms=round(ms, 3), launches=n)
    print(f"{name:28s} {ms:8.3f} ms  {n:5d} launches  {1e3 * ms / max(n, 1):6.2f} us/launch", flush=True)


vel = db["velocity"]
WID = (64, 128, 256, 512)
HW = (64, 32, 16, 8)
for i in range(4):
    nmod = 3 if i < 3 else 4
    feats = [torch.randn(B, HW[i], HW[i], WID[i], device=dev) for _ in range(nmod)]
    dfe = [torch.zeros_like(f) for f in feats]
    dtok = torch.randn(B, nmod * 64, WID[i], device=dev) * 1e-3

    def gpt_fb(i=i, feats=feats, dfe=dfe, dtok=dtok):
        net.gpts[i].fwd(feats, vel, 7, True)
        net.gpts[i].bwd(dtok, dfe)

    def gpt_f(i=i, feats=feats):
        net.gpts[i].fwd(feats, vel, 7, True)
    time_graph(f"gpt{i + 1} fwd", gpt_f)
    time_graph(f"gpt{i + 1} fwd+bwd", gpt_fb)

img = ops.nchw_to_nhwc(db["rgb_u8"], net.mean, net.std)
lid = ops.nchw_to_nhwc(ops.bev_scatter(db["points"]))


def trunk(stem, layers, x0, first):
    def fb():
        x = stem.fwd(x0, True) if stem else x0
        for l in layers[first:]:
            x = l.fwd(x, True)
        d = torch.ones_like(x)
        for l in reversed(layers[first:]):
            d = l.bwd(d)
        if stem:
            stem.bwd(d)

    def f():
        x = stem.fwd(x0, True) if stem else x0
        for l in layers[first:]:
            x = l.fwd(x, True)
    return f, fb


f, fb = trunk(net.img_stem, net.img_layers, img, 0)
time_graph("image ResNet34 fwd", f)
time_graph("image ResNet34 fwd+bwd", fb)
f, fb = trunk(net.lid_stem, net.lid_layers, lid, 0)
time_graph("lidar ResNet18 fwd", f)
time_graph("lidar ResNet18 fwd+bwd", fb)
mp0 = torch.randn(B, 64, 64, 64, device=dev)
f, fb = trunk(None, net.map_layers, mp0, 1)
time_graph("map ResNet34[2:4] fwd", f)
time_graph("map ResNet34[2:4] fwd+bwd", fb)


def vn_fb():
    y = net.vectornet.fwd(db["lane"], db["lane_num"])
    net.vectornet.bwd(torch.ones_like(y))


time_graph("vectornet fwd", lambda: net.vectornet.fwd(db["lane"], db["lane_num"]))
time_graph("vectornet fwd+bwd", vn_fb)


def gat_fb():
    y = net.gat.fwd(db["radar"], db["radar_adj"], 5, True)
    net.gat.bwd(torch.ones_like(y))


time_graph("radar GAT fwd+bwd", gat_fb)


def fwd_only():
    lidar = ops.bev_scatter(db["points"])
    net.forward(db["rgb_u8"], lidar, db["lane"], db["lane_num"], db["radar"], db["radar_adj"],
                db["target_point"], db["velocity"], 0, True)


time_graph("forward only (streams)", fwd_only)
time_graph("forward+backward (streams)", lambda: eng.forward_backward(db))
net.use_streams = False
time_graph("forward+backward (1 trunk stream)", lambda: eng.forward_backward(db))
_Aux.enabled = False
time_graph("forward+backward (serial)", lambda: eng.forward_backward(db))
net.use_streams = True
_Aux.enabled = True
time_graph("adamw", lambda: eng.optimizer_step(collective=False))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(results, open("gpurun_out/ablate.json", "w"), indent=1)
