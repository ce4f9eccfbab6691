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
