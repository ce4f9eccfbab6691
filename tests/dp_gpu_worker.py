"""Rank body of tests/test_gpu_parity.py::test_two_gpu_data_parallel_step_matches_hand_summed_gradients
(launched with torch.distributed.run, one rank per GPU, NCCL).  The reference contract is DistributedDataParallel
(run_steps/phase2_train_net.py:263-269): disjoint shards, rank-local BatchNorm, gradients averaged over ranks,
identical AdamW step everywhere.  Checks, on real hardware:

  1. broadcast_parameters() makes the replicas identical (rank 1 starts from perturbed weights);
  2. after the all-reduce the flat gradient buffer equals the HAND-SUMMED per-rank gradients bit for bit
     (2-rank fp32 sum is order-independent), and its mean matches the CPU oracle's replica average;
  3. parameters are bit-identical on both ranks after an eager step and after two CUDA-graph steps
     (two-bucket exchange, early bucket on the side stream).

Not collected by pytest (no test_ prefix)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mmfn_b200 import ops, parallel, synthetic  # noqa: E402
from mmfn_b200.config import GlobalConfig  # noqa: E402
from mmfn_b200.engine import BatchStager, TrainEngine  # noqa: E402
from mmfn_b200.model_rad import MMFN  # noqa: E402


def gather(t):
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t.contiguous())
    return out


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    assert world == 2
    ops.TF32 = False                                   # exact-fp32 kernels: the oracle comparison below is tight
    cfg = GlobalConfig(embd_pdrop=0.0, attn_pdrop=0.0, resid_pdrop=0.0)
    model = MMFN(cfg, dev)
    sd = synthetic.fill_golden_weights(model.state_dict(), 42)
    model.load_state_dict(sd)
    if rank == 1:
        model.store.flat.add_(1.0)
    eng = TrainEngine(model, lr=1e-4)
    eng.broadcast_parameters()
    flats = gather(model.store.flat)
    assert torch.equal(flats[0], flats[1]), "broadcast_parameters left the replicas different"
    flat0 = model.store.flat.clone()

    per = 2
    gb = synthetic.synth_batch(per * world, n_lanes=32)
    mine = parallel.shard_batch(gb, rank, world)
    db = {k: v.to(dev).contiguous() for k, v in mine.items()}
    n = model.store.n_active
    eng.forward_backward(db)
    local_grad = model.store.flat_grad[:n].clone()
    grads = gather(local_grad)
    assert not torch.equal(grads[0], grads[1]), "ranks saw the same shard"
    hand = grads[0] + grads[1]
    eng.optimizer_step()                               # ONE all-reduce over the flat gradient + fused AdamW (1/world folded in)
    torch.cuda.synchronize()
    assert torch.equal(model.store.flat_grad[:n], hand), "all-reduced gradient != hand-summed per-rank gradients"
    flats = gather(model.store.flat)
    assert torch.equal(flats[0], flats[1]), "parameters differ between ranks after the eager step"
    assert not torch.equal(flats[0][:n], flat0[:n]), "the step did not move the weights"
    assert torch.equal(flats[0][n:], flat0[n:]), "never-used parameters must stay untouched"

    if rank == 0:                                      # oracle: one replica per shard, hand-averaged gradients
        from oracle import bev_oracle, mmfn_oracle
        acc = None
        for r in range(world):
            b = parallel.shard_batch(gb, r, world)
            lidar = torch.from_numpy(np.stack([bev_oracle.lidar_to_histogram_features(p[:, :3].numpy()) for p in b["points"]]))
            inputs = (b["rgb_u8"].float(), lidar, b["lane"], b["lane_num"], b["radar"], b["radar_adj"],
                      b["target_point"], b["velocity"])
            _, _, g = mmfn_oracle.train_step({k: v.clone() for k, v in sd.items()}, cfg,
                                             dict(inputs=inputs, gt_waypoints=b["gt_waypoints"]))
            acc = g if acc is None else {k: (None if v is None else acc[k] + v) for k, v in g.items()}
        dot = n1 = n2 = 0.0
        for k, g in acc.items():
            if g is None:
                continue
            got = model.store.torch_view(k, grad=True).detach().cpu().double() / world
            ref = g.double() / world
            dot += (got * ref).sum().item(); n1 += got.pow(2).sum().item(); n2 += ref.pow(2).sum().item()
        cos = dot / (n1 ** 0.5 * n2 ** 0.5)
        assert cos > 0.9995, cos
        print(f"rank0: averaged-gradient cosine vs oracle replicas = {cos:.6f}", flush=True)

    # CUDA-graph schedule: two-bucket exchange (early bucket under the rest of backward)
    ops.TF32 = True
    stager = BatchStager(mine, dev)
    static = stager.stage(mine)
    torch.cuda.synchronize()
    eng.capture(static, warmup=1)
    for _ in range(2):
        loss = eng.step_graph()
    torch.cuda.synchronize()
    assert torch.isfinite(loss).item()
    flats = gather(model.store.flat)
    assert torch.equal(flats[0], flats[1]), "parameters differ between ranks after the CUDA-graph steps"
    dist.barrier()
    if rank == 0:
        print("DP_GPU_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
