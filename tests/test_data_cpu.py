"""CPU: the data layer against goldens produced by the reference's own
collate_single_cpu / PRE_Data adjacency code (tools/make_goldens.py:collate)."""
import json
import os
import pickle

import numpy as np
import torch

from mmfn_b200 import data, synthetic
from mmfn_b200.config import GlobalConfig
from oracle import bev_oracle


def describe(x):
    if isinstance(x, torch.Tensor):
        return ["tensor", list(x.shape), str(x.dtype)]
    if isinstance(x, dict):
        return {k: describe(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [describe(v) for v in x]
    return [type(x).__name__, x if isinstance(x, (int, float, bool, str)) else None]


def _samples(tmp_path=None):
    smp = [synthetic.synth_sample(i, bev_oracle.lidar_to_histogram_features) for i in range(3)]
    if tmp_path is None:
        for s in smp:
            s["radar_adj"] = synthetic.radar_adjacency(s["radar"][0])
        return smp
    for i, s in enumerate(smp):                              # through the pickle reader, like phase 2
        with open(os.path.join(tmp_path, f"{i}.pkl"), "wb") as fd:
            pickle.dump(s, fd)
    ds = data.PRE_Data(str(tmp_path), GlobalConfig(), "train")
    assert len(ds) == 3
    order = np.argsort([int(os.path.basename(p).split(".")[0]) for p in ds.preload_dict])
    return [ds[int(j)] for j in order]


def test_collate_matches_reference_structure_and_values(golden_dir, tmp_path):
    gold = np.load(os.path.join(golden_dir, "collate_golden.npz"))
    structure = json.load(open(os.path.join(golden_dir, "collate_structure.json")))
    out = data.collate_single_cpu(_samples(tmp_path))
    assert json.loads(json.dumps(describe(out))) == structure          # same tree, shapes and dtypes
    lanes, lane_nums, lmax = out["vectormaps"][0]
    assert np.array_equal(lanes.numpy(), gold["lanes"]) and np.array_equal(lane_nums.numpy(), gold["lane_nums"])
    assert lmax == int(gold["lmax"]) == int(lane_nums.max())
    assert np.array_equal(out["radar_adj"].numpy(), gold["radar_adj"])          # PRE_Data adjacency, bit-exact
    assert np.array_equal(out["radar"][0].numpy(), gold["radar"])
    assert np.array_equal(out["velocity"].numpy(), gold["velocity"])
    assert np.array_equal(torch.stack(out["target_point"], 1).numpy(), gold["target_point"])
    assert np.array_equal(torch.stack([torch.stack(w, 1) for w in out["waypoints"]], 1).numpy(), gold["waypoints"])
    assert int(out["fronts"][0].long().sum()) == int(gold["fronts_sum"])
    assert float(out["lidars"][0].double().sum()) == float(gold["lidars_sum"])
    # padded lanes are zero beyond each sample's lane count (pad_sequence semantics)
    for b, n in enumerate(lane_nums.tolist()):
        assert lanes[b, n:].abs().sum() == 0


def test_to_engine_batch_reproduces_the_train_loop_tensors():
    out = data.collate_single_cpu(_samples())
    eb = data.to_engine_batch(out, seq_len=1, pad_lanes_to=128)
    B = 3
    assert eb["rgb_u8"].shape == (B, 3, 256, 256) and eb["rgb_u8"].dtype == torch.uint8
    assert eb["lidar"].shape == (B, 2, 256, 256) and eb["lane"].shape == (B, 128, 10, 5)
    assert eb["lane_num"].dtype == torch.int32 and eb["gt_waypoints"].shape == (B, 4, 2)
    ref = synthetic.synth_batch(B)                                              # the same frames, direct path
    assert torch.equal(eb["rgb_u8"], ref["rgb_u8"]) and torch.equal(eb["lane_num"], ref["lane_num"])
    assert torch.equal(eb["lane"], ref["lane"]) and torch.allclose(eb["radar_adj"], ref["radar_adj"])
    assert torch.allclose(eb["gt_waypoints"], ref["gt_waypoints"]) and torch.allclose(eb["target_point"], ref["target_point"])
    assert torch.allclose(eb["velocity"], ref["velocity"])
    try:
        data.to_engine_batch(out, pad_lanes_to=64)
        assert False
    except ValueError:
        pass


def test_collate_error_behaviour():
    import pytest
    with pytest.raises(RuntimeError, match="equal size"):
        data.collate_single_cpu([{"a": [1, 2]}, {"a": [1]}])
    with pytest.raises(TypeError):
        data.collate_single_cpu([object(), object()])
    assert data.collate_single_cpu([1.5, 2.5]).dtype == torch.float64
    assert data.collate_single_cpu(["a", "b"]) == ["a", "b"]


def _ragged_samples(n=5):
    lanes = (70, 128, 93, 11, 128)
    smp = [synthetic.synth_sample(60 + i, bev_oracle.lidar_to_histogram_features, n_lanes=128) for i in range(n)]
    for i, s in enumerate(smp):
        s["vectormaps"] = [s["vectormaps"][0][: lanes[i % len(lanes)]]]
    return smp


def test_packed_shard_reproduces_the_pickle_loader_batch(tmp_path):
    """preprocess.write_packed -> data.PackedShard.engine_batch against pickles -> PRE_Data -> collate_single_cpu ->
    to_engine_batch on the same samples: every field the engine consumes is bit-identical; the
    histogram travels as uint8 counts (x 0.2f = the float32 histogram) and the adjacency as float64 azimuths."""
    from mmfn_b200 import preprocess as pp
    smp = _ragged_samples()
    pk = tmp_path / "pkl"
    pk.mkdir()
    for i, s in enumerate(smp):
        with open(pk / f"{i}.pkl", "wb") as fd:
            pickle.dump(s, fd)
    ds = data.PRE_Data(str(pk), GlobalConfig(), "train")
    order = np.argsort([int(os.path.basename(p).split(".")[0]) for p in ds.preload_dict])
    pick = [3, 0, 4, 0]                                             # unordered, with a repeat
    ref = data.to_engine_batch(data.collate_single_cpu([ds[int(order[j])] for j in pick]), pad_lanes_to=128)
    assert pp.write_packed(smp, str(tmp_path / "s0.mmfnpack")) == len(smp)
    shard = data.PackedShard(str(tmp_path / "s0.mmfnpack"))
    assert len(shard) == len(smp)
    got = shard.engine_batch(pick, pad_lanes_to=128)
    for k in ("rgb_u8", "lane", "lane_num", "radar", "velocity", "target_point", "gt_waypoints"):
        assert got[k].dtype == ref[k].dtype and torch.equal(got[k], ref[k]), k
    assert got["lidar_u8"].dtype == torch.uint8 and int(got["lidar_u8"].max()) <= 5
    assert torch.equal(got["lidar_u8"].float() * torch.tensor(0.2, dtype=torch.float32), ref["lidar"])
    az = got["radar_az64"]
    assert az.dtype == torch.float64
    assert torch.equal((az[:, None, :] - az[:, :, None]).to(torch.float32), ref["radar_adj"])
    # unpadded: the batch maximum, as collate_single_cpu pads
    assert shard.engine_batch([0, 3])["lane"].shape[1] == 70
    # a shard sample converts back to the pickle layout PRE_Data would read
    back = shard.sample(2)
    assert torch.equal(back["fronts"][0], smp[2]["fronts"][0]) and torch.equal(back["maps"][0], smp[2]["maps"][0])
    assert np.array_equal(back["lidars"][0], smp[2]["lidars"][0]) and back["lidars"][0].dtype == np.float32
    assert np.array_equal(np.asarray(back["waypoints"]), np.asarray(smp[2]["waypoints"]))
    assert back["target_point"] == tuple(smp[2]["target_point"]) and back["velocity"] == smp[2]["velocity"]
    assert torch.equal(back["vectormaps"][0].float(), smp[2]["vectormaps"][0].float())
    assert np.array_equal(back["radar"][0].astype(np.float32), smp[2]["radar"][0].astype(np.float32))


def test_packed_loader_follows_the_distributed_sampler(tmp_path):
    """PackedLoader shards like torch's DistributedSampler (phase2_train_net.py:265-267) and batches across shards."""
    from mmfn_b200 import preprocess as pp
    smp = _ragged_samples(5)
    pp.write_packed(smp[:3], str(tmp_path / "a.mmfnpack"))
    pp.write_packed(smp[3:], str(tmp_path / "b.mmfnpack"))
    paths = [str(tmp_path / "a.mmfnpack"), str(tmp_path / "b.mmfnpack")]
    whole = data.PackedShard(paths[0]), data.PackedShard(paths[1])
    for world in (1, 2):
        for epoch in (0, 1):
            seen = []
            for rank in range(world):
                ld = data.PackedLoader(paths, batch_size=2, shuffle=True, seed=3, rank=rank, world=world, pad_lanes_to=128)
                ld.set_epoch(epoch)
                sampler = torch.utils.data.distributed.DistributedSampler(range(5), num_replicas=world, rank=rank, shuffle=True, seed=3)
                sampler.set_epoch(epoch)
                want = list(sampler)
                assert ld.indices() == want
                batches = list(ld)
                assert len(batches) == len(ld) == len(want) // 2
                for bi, b in enumerate(batches):
                    for j, gi in enumerate(want[bi * 2: bi * 2 + 2]):
                        sh, li = (whole[0], gi) if gi < 3 else (whole[1], gi - 3)
                        one = sh.engine_batch([li], pad_lanes_to=128)
                        for k in one:
                            assert torch.equal(b[k][j], one[k][0]), (k, gi)
                seen += want
            assert set(seen) == set(range(5))
