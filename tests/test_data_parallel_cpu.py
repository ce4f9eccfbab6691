"""world_size-2 gloo test (CPU) of the data-parallel wiring: disjoint shards, ONE flat-gradient
all-reduce, 1/world folded into AdamW -- checked against the multi-replica oracle of SURVEY.md 8(e):
N replicas on the N shards with hand-averaged gradients (NOT one replica on the concatenated
batch: BatchNorm statistics stay rank-local)."""
import os
import tempfile

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mmfn_b200 import parallel, synthetic
from mmfn_b200.config import GlobalConfig
from mmfn_b200.params import ParamStore, param_spec, is_unused
from oracle import bev_oracle, mmfn_oracle


def test_shard_range():
    assert parallel.shard_range(32, 0, 2) == (0, 16) and parallel.shard_range(32, 1, 2) == (16, 32)
    assert [parallel.shard_range(8, r, 8) for r in range(8)] == [(r, r + 1) for r in range(8)]
    try:
        parallel.shard_range(10, 0, 4)
        assert False
    except ValueError:
        pass
    b = {"x": torch.arange(8).view(8, 1), "y": torch.arange(16).view(8, 2)}
    s = parallel.shard_batch(b, 1, 4)
    assert s["x"].flatten().tolist() == [2, 3] and s["y"].shape == (2, 2)


def _small_step(sd, cfg, b):
    lidar = torch.from_numpy(np.stack([bev_oracle.lidar_to_histogram_features(p[:, :3].numpy()) for p in b["points"]]))
    inputs = (b["rgb_u8"].float(), lidar, b["lane"], b["lane_num"], b["radar"], b["radar_adj"],
              b["target_point"], b["velocity"])
    return mmfn_oracle.train_step(sd, cfg, dict(inputs=inputs, gt_waypoints=b["gt_waypoints"]))


def _worker(rank, world, initfile, outdir):
    dist.init_process_group("gloo", init_method=f"file://{initfile}", rank=rank, world_size=world)
    torch.set_num_threads(4)
    cfg = GlobalConfig(embd_pdrop=0.0, attn_pdrop=0.0, resid_pdrop=0.0)
    store = ParamStore(cfg, "cpu")                       # the product's flat layout, on CPU
    root = torch.nn.Module()
    store.register(root)
    sd0 = synthetic.fill_golden_weights(root.state_dict(), 42)
    # rank 1 starts from different weights: broadcast_ must make the replicas identical
    if rank == 1:
        sd0 = {k: (v + 1 if v.dtype.is_floating_point else v) for k, v in sd0.items()}
    root.load_state_dict(sd0)
    parallel.broadcast_([store.flat, store.flat_buf], src=0)
    sd = {k: v.detach().clone() for k, v in root.state_dict().items()}
    gb = synthetic.synth_batch(2 * world, n_lanes=16)    # global batch, every rank builds the same one
    mine = parallel.shard_batch(gb, rank, world)
    loss, _, grads = _small_step(sd, cfg, mine)
    # local gradients -> flat buffer (what the CUDA backward writes), then the gradient exchange exactly as
    # TrainEngine.step_graph issues it: the early bucket [n_mid, n_active) first, then the mid bucket [n_late, n_mid)
    # (both under the rest of backward on the GPU), then the late bucket [0, n_late).  Together they must equal ONE
    # all-reduce over [0, n_active).
    store.flat_grad.zero_()
    for k, g in grads.items():
        if g is not None:
            store.torch_view(k, grad=True).copy_(g)
    whole = store.flat_grad[: store.n_active].clone()
    parallel.allreduce_sum_(whole)
    parallel.allreduce_sum_(store.flat_grad[store.n_mid: store.n_active])
    parallel.allreduce_sum_(store.flat_grad[store.n_late: store.n_mid])
    parallel.allreduce_sum_(store.flat_grad[: store.n_late])
    assert torch.equal(whole, store.flat_grad[: store.n_active])
    torch.save({"flat_grad": store.flat_grad.clone(), "loss": loss, "w": store.flat[:1000].clone(),
                "local": {k: g for k, g in grads.items() if g is not None and "decoder" in k or k == "join.0.weight"}},
               os.path.join(outdir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_matches_multi_replica_oracle():
    world = 2
    with tempfile.TemporaryDirectory() as d:
        initfile = os.path.join(d, "init")
        mp.spawn(_worker, args=(world, initfile, d), nprocs=world, join=True)
        r = [torch.load(os.path.join(d, f"rank{i}.pt")) for i in range(world)]
    # every rank holds the same summed gradient and (after broadcast) the same weights
    assert torch.equal(r[0]["flat_grad"], r[1]["flat_grad"]) and torch.equal(r[0]["w"], r[1]["w"])
    # the sum equals the hand-summed per-replica gradients; the mean is what AdamW consumes (grad_scale = 1/world)
    cfg = GlobalConfig()
    store = ParamStore(cfg, "cpu")
    for k in ("join.0.weight", "decoder.weight_hh"):
        hand = r[0]["local"][k] + r[1]["local"][k]
        store.flat_grad.copy_(r[0]["flat_grad"])
        assert torch.allclose(store.torch_view(k, grad=True), hand, rtol=1e-6, atol=1e-7), k
    # shards differ, so the local losses must differ (disjoint data), and never-used parameters stay zero
    assert abs(r[0]["loss"].item() - r[1]["loss"].item()) > 1e-6
    unused = [k for k, _, kind in param_spec(cfg) if kind in ("conv", "f") and is_unused(k)]
    for k in unused[:3]:
        assert store.torch_view(k, grad=True).abs().max().item() == 0.0
    assert store.n_active + sum(store.torch_view(k).numel() for k in unused) <= store.n_total
