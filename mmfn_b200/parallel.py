"""Data-parallel plumbing (one process per GPU).  The reference's only parallelism is
DistributedDataParallel + DistributedSampler (run_steps/phase2_train_net.py:227, :263-269); the
equivalent here is a disjoint contiguous shard of every global batch per rank and ONE all-reduce
over the flat gradient buffer per step.  Device-agnostic so the wiring can be
tested with the gloo backend on CPU."""
import torch
import torch.distributed as dist


def shard_range(global_batch, rank, world):
    """Samples [lo, hi) of a global batch owned by `rank`: equal contiguous shards, like
    DistributedSampler(shuffle=False) over one batch."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world}")
    per = global_batch // world
    return rank * per, (rank + 1) * per


def shard_batch(batch, rank, world):
    """Slice every tensor of a host batch (dict) along dim 0 to this rank's shard."""
    n = next(iter(batch.values())).shape[0]
    lo, hi = shard_range(n, rank, world)
    return {k: v[lo:hi] for k, v in batch.items()}


def allreduce_sum_(flat_grad, group=None):
    """ONE collective over the flat gradient buffer (in place).  The 1/world scaling is folded into
    the fused AdamW kernel (grad_scale), not applied here."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


def broadcast_(tensors, src=0, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in tensors:
            dist.broadcast(t, src, group=group)
