"""Host-side data layer of the training path.

  PRE_Data               mirror of team_code/mmfn_utils/datasets/dataloader.py:349-384 (pkl reader that adds
                         the 81x81 radar "adjacency" adj[i, j] = radar[j, 1] - radar[i, 1])
  collate_single_cpu     mirror of team_code/mmfn_utils/datasets/data_utils.py:9-67: default-collate with the
                         "vectormaps" special case (pad lanes to the batch maximum, return
                         [padded, lane_nums, max]); written for torch 2.x (the reference imports the removed
                         torch._six module)
  to_engine_batch        what Engine.train does between the loader and the model
                         (run_steps/phase2_train_net.py:66-103), producing the packed-batch dict the
                         TrainEngine / BatchStager consume

  PackedShard / PackedLoader   the packed alternative to the pickles (written by
                         preprocess.write_packed): memory-mapped arrays -> engine batch with no unpickling, no collate and
                         no per-sample Python; the histogram travels as uint8 counts and the radar adjacency is built on
                         the GPU from the float64 azimuths (TrainEngine expands both)

Samples are the dicts CARLA_Data.__getitem__ builds (dataloader.py:183-268): lists of per-timestep
tensors for fronts / lidars / maps / vectormaps / radar, tuples for waypoints / target_point, floats.
"""
import collections.abc
import os
import pickle

import numpy as np
import torch
from torch.nn.utils.rnn import pad_sequence


class PRE_Data(torch.utils.data.Dataset):
    def __init__(self, root, config, data_use="train"):
        self.seq_len, self.pred_len = config.seq_len, config.pred_len
        preload_file = os.path.join(root, f"rg_vec_mmfn_diag_pl_{self.seq_len}_{self.pred_len}_{data_use}.npy")
        if not os.path.exists(preload_file):
            files = [os.path.join(str(root), f) for f in os.listdir(root) if f.split(".")[-1] == "pkl"]
            np.save(preload_file, files)
        self.preload_dict = np.load(preload_file)

    def __len__(self):
        return len(self.preload_dict)

    def __getitem__(self, index):
        with open(self.preload_dict[index], "rb") as fd:
            data = pickle.load(fd)
        az = np.asarray(data["radar"][0])[:, 1]
        data["radar_adj"] = az[None, :] - az[:, None]          # row i: radar[:, 1] - radar[i, 1]
        return data


def collate_single_cpu(batch, now_key=""):
    """Same output structure, dtypes and values as the reference collate."""
    elem = batch[0]
    if now_key == "vectormaps" and isinstance(elem, torch.Tensor):
        lane_nums = torch.tensor([lane.shape[0] for lane in batch])
        return [pad_sequence(batch, batch_first=True), lane_nums, int(lane_nums.max().item())]
    if isinstance(elem, torch.Tensor):
        return torch.stack(batch, 0)
    if isinstance(elem, np.ndarray):
        if elem.dtype.kind in "SaUO":
            raise TypeError(f"collate_single_cpu: unsupported array dtype {elem.dtype}")
        return collate_single_cpu([torch.as_tensor(b) for b in batch], now_key)
    if isinstance(elem, np.generic):
        return torch.as_tensor(np.asarray(batch))
    if isinstance(elem, float):
        return torch.tensor(batch, dtype=torch.float64)
    if isinstance(elem, int):
        return torch.tensor(batch)
    if isinstance(elem, (str, bytes)):
        return batch
    if isinstance(elem, collections.abc.Mapping):
        return {key: collate_single_cpu([d[key] for d in batch], key) for key in elem}
    if isinstance(elem, tuple) and hasattr(elem, "_fields"):
        return type(elem)(*(collate_single_cpu(s, now_key) for s in zip(*batch)))
    if isinstance(elem, collections.abc.Sequence):
        n = len(elem)
        if not all(len(e) == n for e in batch):
            raise RuntimeError("each element in list of batch should be of equal size")
        return [collate_single_cpu(samples, now_key) for samples in zip(*batch)]
    raise TypeError(f"collate_single_cpu: batch must contain tensors, numpy arrays, numbers, dicts or lists; found {type(elem)}")


def to_engine_batch(data, seq_len=1, pad_lanes_to=None):
    """Collated reference batch -> the flat dict TrainEngine consumes (all CPU tensors, ready for
    BatchStager).  `lidar` carries the pre-computed BEV histogram stored by phase 1; when the samples hold a
    raw sweep under `points` it is passed through instead and the histogram is built on the GPU."""
    lane, lane_num, lmax = data["vectormaps"][0]
    lane = lane.to(torch.float32)
    if pad_lanes_to is not None:                              # fixed shapes keep one CUDA graph valid
        if lmax > pad_lanes_to:
            raise ValueError(f"batch has {lmax} lanes, more than pad_lanes_to={pad_lanes_to}")
        lane = torch.nn.functional.pad(lane, (0, 0, 0, 0, 0, pad_lanes_to - lane.shape[1]))
    wps = [torch.stack(data["waypoints"][i], dim=1) for i in range(seq_len, len(data["waypoints"]))]
    out = {
        "rgb_u8": data["fronts"][0].to(torch.uint8),
        "lane": lane.contiguous(),
        "lane_num": lane_num.to(torch.int32),
        "radar": data["radar"][0].to(torch.float32),
        "radar_adj": data["radar_adj"].to(torch.float32),
        "velocity": data["velocity"].to(torch.float32),
        "target_point": torch.stack(data["target_point"], dim=1).to(torch.float32),
        "gt_waypoints": torch.stack(wps, dim=1).to(torch.float32),
    }
    if "points" in data:
        out["points"] = data["points"][0].to(torch.float32)
    else:
        out["lidar"] = data["lidars"][0].to(torch.float32)
    return out


class PackedShard:
    """Reader of one shard written by preprocess.write_packed.  engine_batch(indices) returns the dict to_engine_batch
    builds from the same samples -- bit for bit -- except that `lidar` (float32 histogram) is `lidar_u8` (counts; the
    engine multiplies by 0.2f on the GPU) and `radar_adj` is `radar_az64` (the engine forms az[j] - az[i] in float64)."""

    def __init__(self, path):
        import json
        from .preprocess import PACK_ALIGN, PACK_MAGIC
        with open(path, "rb") as fd:
            if fd.read(len(PACK_MAGIC)) != PACK_MAGIC:
                raise ValueError(f"{path}: not an MMFN packed shard")
            hlen = int(np.frombuffer(fd.read(8), dtype=np.uint64)[0])
            head = json.loads(fd.read(hlen).decode())
        data0 = (len(PACK_MAGIC) + 8 + hlen + PACK_ALIGN - 1) // PACK_ALIGN * PACK_ALIGN
        self.n = int(head["n"])
        self.arrays = {name: np.memmap(path, mode="r", dtype=np.dtype(m["dtype"]), shape=tuple(m["shape"]), offset=data0 + m["offset"])
                       for name, m in head["arrays"].items()}

    def __len__(self):
        return self.n

    def engine_batch(self, indices, seq_len=1, pad_lanes_to=None):
        a = self.arrays
        idx = np.asarray(indices, dtype=np.int64)
        order = np.argsort(idx, kind="stable")                     # memmap gathers in file order, then restored
        inv = np.empty_like(order)
        inv[order] = np.arange(len(order))

        def take(name):
            return torch.from_numpy(np.ascontiguousarray(a[name][idx[order]][inv]))
        lane_num = take("lane_num")
        lmax = int(lane_num.max())
        L = lmax if pad_lanes_to is None else pad_lanes_to
        if lmax > L:
            raise ValueError(f"batch has {lmax} lanes, more than pad_lanes_to={pad_lanes_to}")
        lane_all = take("lane")
        lane = torch.zeros((len(idx), L) + tuple(lane_all.shape[2:]), dtype=torch.float32)
        k = min(L, lane_all.shape[1])
        lane[:, :k] = lane_all[:, :k]
        wps = take("waypoints")
        return {
            "rgb_u8": take("fronts"),
            "lidar_u8": take("lidar_u8"),
            "lane": lane,
            "lane_num": lane_num.to(torch.int32),
            "radar": take("radar"),
            "radar_az64": take("radar_az64"),
            "velocity": take("velocity").to(torch.float32),
            "target_point": take("target_point").to(torch.float32),
            "gt_waypoints": wps[:, seq_len:].to(torch.float32).contiguous(),
        }

    def sample(self, i):
        """Sample i back in the phase-1 pickle layout (what PRE_Data would unpickle) -- for interchange with the reference."""
        a = self.arrays
        n = int(a["lane_num"][i])
        return {
            "fronts": [torch.from_numpy(np.array(a["fronts"][i]))],
            "lidars": [np.array(a["lidar_u8"][i], dtype=np.float32) * np.float32(0.2)],
            "vectormaps": [torch.from_numpy(np.array(a["lane"][i, :n], dtype=np.float64))],
            "radar": [np.array(a["radar"][i], dtype=np.float64)],
            "maps": [torch.from_numpy(np.array(a["maps"][i]))],
            "waypoints": [tuple(w) for w in np.array(a["waypoints"][i]).tolist()],
            "target_point": tuple(np.array(a["target_point"][i]).tolist()),
            "steer": float(a["steer"][i]), "throttle": float(a["throttle"][i]), "brake": bool(a["brake"][i]),
            "command": int(a["command"][i]), "velocity": float(a["velocity"][i]),
        }


class PackedLoader:
    """Batches of one or more shards for TrainEngine / BatchStager.  Index order follows torch's DistributedSampler
    (phase2_train_net.py:265-267): a seeded permutation per epoch, padded to a multiple of the world size, rank r takes
    every world-th index; incomplete last batches are dropped (fixed shapes keep one CUDA graph valid)."""

    def __init__(self, shards, batch_size, shuffle=True, seed=0, rank=0, world=1, seq_len=1, pad_lanes_to=None):
        self.shards = [s if isinstance(s, PackedShard) else PackedShard(s) for s in shards]
        self.starts = np.cumsum([0] + [len(s) for s in self.shards])
        self.batch_size, self.shuffle, self.seed, self.rank, self.world = batch_size, shuffle, seed, rank, world
        self.seq_len, self.pad_lanes_to, self.epoch = seq_len, pad_lanes_to, 0

    def set_epoch(self, epoch):
        self.epoch = epoch

    def indices(self):
        n = int(self.starts[-1])
        if self.shuffle:
            g = torch.Generator()
            g.manual_seed(self.seed + self.epoch)
            idx = torch.randperm(n, generator=g).tolist()
        else:
            idx = list(range(n))
        total = (n + self.world - 1) // self.world * self.world
        idx += idx[: total - n]
        return idx[self.rank: total: self.world]

    def __len__(self):
        return len(self.indices()) // self.batch_size

    def __iter__(self):
        idx = self.indices()
        for b in range(len(idx) // self.batch_size):
            chunk = np.asarray(idx[b * self.batch_size: (b + 1) * self.batch_size])
            shard_of = np.searchsorted(self.starts, chunk, side="right") - 1
            if len(set(shard_of.tolist())) == 1:
                s = int(shard_of[0])
                yield self.shards[s].engine_batch(chunk - self.starts[s], self.seq_len, self.pad_lanes_to)
            else:                                                  # a batch that straddles shards: per-shard gathers, restored order
                parts, pos = [], []
                for s in sorted(set(shard_of.tolist())):
                    m = np.nonzero(shard_of == s)[0]
                    parts.append(self.shards[s].engine_batch(chunk[m] - self.starts[s], self.seq_len,
                                                             self.pad_lanes_to if self.pad_lanes_to is not None else
                                                             max(int(sh.arrays["lane"].shape[1]) for sh in self.shards)))
                    pos.append(m)
                order = np.argsort(np.concatenate(pos))
                yield {k: torch.cat([p[k] for p in parts])[order] for k in parts[0]}
