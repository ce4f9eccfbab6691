"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs, and against the goldens produced by the unmodified reference.

Tolerances: BEV bins bit-exact; fp32 path waypoints within 1e-3 L1 of the reference (north_star);
measured error is ~1e-5, asserted at 2e-4.  Gradients: relative to each tensor's norm.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from mmfn_b200 import synthetic  # noqa: E402
from mmfn_b200.config import GlobalConfig  # noqa: E402
from oracle import bev_oracle, mmfn_oracle  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def test_bev_scatter_bit_exact_vs_oracle_and_goldens(dev, golden_dir):
    from mmfn_b200 import ops
    gold = np.load(os.path.join(golden_dir, "bev_golden.npz"))
    for key in gold.files:
        seed, n = int(key.split("_")[0][1:]), int(key.split("_n")[1])
        pts = synthetic.synth_points(seed, n)
        ref = (gold[key].astype(np.float64) / 5).astype(np.float32)
        for stride in (4, 3):
            p = torch.from_numpy(np.ascontiguousarray(pts[None, :, :stride])).to(dev)
            for strips in (0, -1, 2, 4, 8, 16):
                out = ops.bev_scatter(p, strips).cpu().numpy()[0]
                assert np.array_equal(out, ref), (key, stride, strips)
    # a ragged multi-frame batch against the oracle
    pts = np.stack([synthetic.synth_points(500 + i, 4096) for i in range(5)])
    out = ops.bev_scatter(torch.from_numpy(pts).to(dev)).cpu().numpy()
    for i in range(5):
        assert np.array_equal(out[i], bev_oracle.lidar_to_histogram_features(pts[i, :, :3]))


def test_bev_scatter_full_size_properties(dev):
    """Size-independent checks at BASELINE's full size (64 frames x 32768 points)."""
    from mmfn_b200 import ops
    pts = torch.from_numpy(np.stack([synthetic.synth_points(9000 + i) for i in range(64)])).to(dev)
    out = ops.bev_scatter(pts)
    assert torch.equal(ops.bev_scatter(pts, -1), out)          # strip kernel (64 frames) == one-visit kernels
    vals = torch.unique(out).cpu().numpy()
    assert set(np.round(vals * 5).astype(int)) <= {0, 1, 2, 3, 4, 5}
    # permutation invariance and frame independence
    perm = torch.randperm(pts.shape[1], device=dev)
    assert torch.equal(ops.bev_scatter(pts[:, perm].contiguous()), out)
    assert torch.equal(ops.bev_scatter(pts[7:8].contiguous())[0], out[7])
    # checksum against the oracle on a sample of frames
    for i in (0, 63):
        assert np.array_equal(out[i].cpu().numpy(), bev_oracle.lidar_to_histogram_features(pts[i, :, :3].cpu().numpy()))


def _setup(dev, B, tf32=False):
    """tf32=False: exact-fp32 SIMT kernels everywhere (tight tolerances).  tf32=True: the production
    path, large GEMMs/convs on the tcgen05 TF32 tensor-core kernels (north_star tolerance 1e-3)."""
    from mmfn_b200 import ops
    from mmfn_b200.model_rad import MMFN
    ops.TF32 = tf32
    cfg = GlobalConfig(embd_pdrop=0.0, attn_pdrop=0.0, resid_pdrop=0.0)
    model = MMFN(cfg, dev)
    sd = synthetic.fill_golden_weights(model.state_dict(), 42)
    model.load_state_dict(sd)
    model.train()
    b = synthetic.synth_batch(B)
    return cfg, model, sd, b


def _probe(t, n=16):
    f = t.detach().reshape(-1).double().cpu()
    idx = torch.linspace(0, f.numel() - 1, n).long()
    return np.concatenate([[f.norm().item(), f.sum().item()], f[idx].numpy()])


@pytest.mark.parametrize("tf32", [False, True], ids=["fp32-simt", "tf32-tcgen05"])
def test_train_step_matches_oracle_and_reference_goldens(dev, golden_dir, tf32):
    """Reference-style call sequence (Engine.train): model(...) -> l1 -> loss.backward()."""
    from mmfn_b200 import ops
    B = 2
    cfg, model, sd, b = _setup(dev, B, tf32)
    # measured on B200: waypoint L1 1.5e-7 (fp32) / 1.7e-4 (tf32); loss error 0 / 1e-6;
    # worst per-tensor gradient error 1.5e-2 (fp32, B=2 train-mode BN amplifies fp32 noise) / 0.37 (tf32)
    # (grad_tol 3e-2: the worst tensor moves between 1.5e-2 and 2.2e-2 with the summation order of split-K atomics)
    wp_tol, loss_tol, grad_tol, norm_tol = (2e-4, 2e-4, 3e-2, 2e-2) if not tf32 else (1e-3, 1e-3, 0.6, 0.25)
    lidar = ops.bev_scatter(b["points"].to(dev))
    lidar_ref = torch.from_numpy(np.stack([bev_oracle.lidar_to_histogram_features(p[:, :3].numpy()) for p in b["points"]]))
    assert torch.equal(lidar.cpu(), lidar_ref)
    vectormaps = [[b["lane"].to(dev)], [b["lane_num"].to(dev).float()], b["lane"].shape[1]]
    pred = model([b["rgb_u8"].to(dev).float()], [lidar], None, vectormaps, [b["radar"].to(dev)],
                 [b["radar_adj"].to(dev)], b["target_point"].to(dev), b["velocity"].to(dev))
    loss = torch.nn.functional.l1_loss(pred, b["gt_waypoints"].to(dev), reduction="none").mean()
    loss.backward()

    # oracle on the CPU with the same weights / inputs
    osd = {k: v.clone() for k, v in sd.items()}
    inputs = (b["rgb_u8"].float(), lidar_ref, b["lane"], b["lane_num"], b["radar"], b["radar_adj"],
              b["target_point"], b["velocity"])
    oloss, opred, ograds = mmfn_oracle.train_step(osd, cfg, dict(inputs=inputs, gt_waypoints=b["gt_waypoints"]))
    wp_l1 = (pred.detach().cpu() - opred).abs().mean().item()
    assert wp_l1 < wp_tol, wp_l1                     # north_star bar: 1e-3
    assert abs(loss.item() - oloss.item()) < loss_tol

    # goldens of the real reference
    gold = np.load(os.path.join(golden_dir, "mmfn_golden_b2.npz"))
    assert np.abs(pred.detach().cpu().numpy() - gold["pred_wp"]).mean() < wp_tol
    assert abs(loss.item() - float(gold["loss"])) < loss_tol

    worst, worst_key = 0.0, None
    dot = n1 = n2 = 0.0
    params = dict(model.named_parameters())
    for k, g in ograds.items():
        p = params[k]
        if g is None:
            assert p.grad is None, k
            continue
        got = p.grad.detach().cpu()
        denom = max(g.norm().item(), 1e-6)
        err = (got - g).norm().item() / denom
        if err > worst:
            worst, worst_key = err, k
        dot += (got.double() * g.double()).sum().item()
        n1 += got.double().pow(2).sum().item()
        n2 += g.double().pow(2).sum().item()
        ref = gold["grad/" + k]
        assert abs(_probe(got, 6)[0] - ref[0]) <= norm_tol * max(ref[0], 1e-6), k
    assert worst < grad_tol, (worst, worst_key)
    cosine = dot / (n1 ** 0.5 * n2 ** 0.5)           # whole-model gradient direction
    assert cosine > (0.9999 if not tf32 else 0.97), cosine

    # BatchNorm running statistics after one training step
    msd = model.state_dict()
    for k in osd:
        if k.endswith("running_mean") or k.endswith("running_var"):
            rt, at = (1e-4, 1e-5) if not tf32 else (1e-2, 2e-3)      # TF32 conv outputs feed the statistics
            assert torch.allclose(msd[k].cpu(), osd[k], rtol=rt, atol=at), k
        if k.endswith("num_batches_tracked"):
            assert int(msd[k]) == int(osd[k]), k


def test_engine_steps_match_oracle_adamw(dev):
    """Two full optimisation steps through TrainEngine (BEV scatter + fwd + bwd + fused AdamW)."""
    from mmfn_b200.engine import TrainEngine
    B = 2
    cfg, model, sd, b = _setup(dev, B, tf32=False)
    eng = TrainEngine(model, lr=1e-4)
    osd = {k: v.clone() for k, v in sd.items()}
    opt = {"t": 0, "m": {}, "v": {}}
    db = {k: v.to(dev) for k, v in b.items()}
    lidar_ref = torch.from_numpy(np.stack([bev_oracle.lidar_to_histogram_features(p[:, :3].numpy()) for p in b["points"]]))
    inputs = (b["rgb_u8"].float(), lidar_ref, b["lane"], b["lane_num"], b["radar"], b["radar_adj"],
              b["target_point"], b["velocity"])
    for step in range(2):
        loss = eng.step(db)
        oloss, _, _ = mmfn_oracle.train_step(osd, cfg, dict(inputs=inputs, gt_waypoints=b["gt_waypoints"]), opt_state=opt)
        # step 2 runs on weights that already took one ~lr*sign(g) AdamW move; elements with g ~ 0
        # may move the other way, so the second loss is compared a little looser
        assert abs(loss.item() - oloss.item()) < (2e-4 if step == 0 else 3e-3), (step, loss.item(), oloss.item())
    msd = model.state_dict()
    # after 2 AdamW steps every trained weight moved by <= ~2*lr; compare the moved weights
    for k in ("encoder.transformer4.blocks.7.mlp.2.weight", "join.0.weight", "decoder.weight_hh",
              "encoder.image_encoder.features.layer4.2.conv2.weight", "encoder.vectornet_encoder.generator.3.bias"):
        d_mine = (msd[k].cpu() - sd[k])
        d_orac = (osd[k] - sd[k])
        close = ((d_mine - d_orac).abs() < 5e-5).float().mean().item()
        assert close > 0.95, (k, close)
        assert (d_mine - d_orac).abs().max().item() < 4.1e-4, k      # at most 2 steps x 2*lr
    # untouched (never-used) parameters keep their exact values: no weight decay applied
    k = "encoder.img_map_encoder.features.layer1.0.conv1.weight"
    assert torch.equal(msd[k].cpu(), sd[k])


def test_edge_case_inputs_match_oracle(dev):
    """Degenerate frames the reference data can contain: a single valid lane next to a full lane set (masking of
    the padded lanes), a radar frame with no returns (all-zero rows -> every GAT edge masked), and a LiDAR sweep with
    no point inside the BEV grid.  Exact-fp32 path against the CPU oracle."""
    from mmfn_b200 import ops
    from mmfn_b200.engine import TrainEngine
    B = 2
    cfg, model, sd, b = _setup(dev, B, tf32=False)
    b["lane_num"] = torch.tensor([1, b["lane"].shape[1]], dtype=torch.int32)
    b["lane"][0, 1:] = 0                                    # pad_sequence zero padding (data_utils.py:42-48)
    b["radar"][0] = 0
    b["radar_adj"][0] = 0
    b["points"][1, :, 0] = 100.0                            # every point outside [-16, 16]
    eng = TrainEngine(model, lr=1e-4)
    db = {k: v.to(dev) for k, v in b.items()}
    lidar = ops.bev_scatter(db["points"])
    lidar_ref = torch.from_numpy(np.stack([bev_oracle.lidar_to_histogram_features(p[:, :3].numpy()) for p in b["points"]]))
    assert torch.equal(lidar.cpu(), lidar_ref) and lidar_ref[1].abs().sum() == 0
    loss = eng.forward_backward(db)
    inputs = (b["rgb_u8"].float(), lidar_ref, b["lane"], b["lane_num"], b["radar"], b["radar_adj"],
              b["target_point"], b["velocity"])
    oloss, opred, ograds = mmfn_oracle.train_step({k: v.clone() for k, v in sd.items()}, cfg,
                                                  dict(inputs=inputs, gt_waypoints=b["gt_waypoints"]))
    assert (eng.last_pred.cpu() - opred).abs().mean().item() < 2e-4
    assert abs(loss.item() - oloss.item()) < 2e-4
    params = dict(model.named_parameters())
    dot = n1 = n2 = 0.0
    for k, g in ograds.items():
        if g is None:
            continue
        got = model.store.torch_view(k, grad=True).detach().cpu().double()
        assert torch.isfinite(got).all(), k
        dot += (got * g.double()).sum().item(); n1 += got.pow(2).sum().item(); n2 += g.double().pow(2).sum().item()
    assert dot / (n1 ** 0.5 * n2 ** 0.5) > 0.999


def test_config2_full_size_properties(dev):
    """BASELINE configs[1] at its full size (B=16, production TF32 path, dropout off): properties that need no
    CPU oracle run -- finite outputs, bit-exact BEV for sampled frames, and batch-permutation equivariance (train-mode
    BatchNorm statistics are order-invariant, so permuting the 16 frames must permute the waypoints and leave the
    loss and the summed gradients unchanged up to accumulation order)."""
    from mmfn_b200 import ops
    from mmfn_b200.engine import TrainEngine
    B = 16
    cfg, model, sd, b = _setup(dev, B, tf32=True)
    eng = TrainEngine(model, lr=1e-4)
    db = {k: v.to(dev) for k, v in b.items()}
    lidar = ops.bev_scatter(db["points"])
    for i in (0, 15):
        assert np.array_equal(lidar[i].cpu().numpy(), bev_oracle.lidar_to_histogram_features(b["points"][i, :, :3].numpy()))
    loss = eng.forward_backward(db).item()
    pred = eng.last_pred.clone()
    grad = model.store.flat_grad[: model.store.n_active].clone()
    assert np.isfinite(loss) and torch.isfinite(pred).all() and torch.isfinite(grad).all()
    assert pred.shape == (B, 4, 2) and grad.abs().sum().item() > 0
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(3))
    dbp = {k: v[perm.to(dev)].contiguous() for k, v in db.items()}
    loss_p = eng.forward_backward(dbp).item()
    assert abs(loss_p - loss) < 1e-4 * max(1.0, abs(loss))
    assert (eng.last_pred - pred[perm.to(dev)]).abs().max().item() < 2e-3
    grad_p = model.store.flat_grad[: model.store.n_active]
    cos = torch.dot(grad_p.double(), grad.double()) / (grad_p.double().norm() * grad.double().norm())
    assert cos.item() > 0.999, cos.item()


def _report(name, payload):
    """Measured parity figures are also dropped under gpurun_out/ (scratch) so a GPU session can copy them to profiles/."""
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, name), "w") as f:
            json.dump(payload, f, indent=1)
    except OSError:
        pass


def _grad_stats(model, ograds):
    """whole-model cosine + per-tensor relative errors of store.flat_grad against the oracle gradients.
    attn.key.bias tensors are skipped: their true gradient is exactly zero (a constant added to every key shifts all
    scores of a softmax row equally), so both sides hold nothing but rounding noise."""
    dot = n1 = n2 = 0.0
    rel = {}
    for k, g in ograds.items():
        if g is None or k.endswith("attn.key.bias"):
            continue
        got = model.store.torch_view(k, grad=True).detach().cpu().double()
        gd = g.double()
        assert torch.isfinite(got).all(), k
        dot += (got * gd).sum().item(); n1 += got.pow(2).sum().item(); n2 += gd.pow(2).sum().item()
        rel[k] = ((got - gd).norm() / gd.norm().clamp_min(1e-12)).item()
    return dot / (n1 ** 0.5 * n2 ** 0.5), rel


def test_config2_b16_matches_oracle(dev):
    """BASELINE configs[1] at the BENCHMARKED size (B=16, production TF32 tensor-core path, dropout off, train-mode
    BatchNorm) against the CPU oracle on the same seeded frames: waypoints (north_star bar 1e-3 L1), loss, every
    parameter gradient.  Measured on B200 (profiles/r02_parity_b16.json): see the asserted bounds."""
    from mmfn_b200.engine import TrainEngine
    B = 16
    cfg, model, sd, b = _setup(dev, B, tf32=True)
    eng = TrainEngine(model, lr=1e-4)
    db = {k: v.to(dev) for k, v in b.items()}
    loss = eng.forward_backward(db).item()
    torch.cuda.synchronize()
    lidar_ref = torch.from_numpy(np.stack([bev_oracle.lidar_to_histogram_features(p[:, :3].numpy()) for p in b["points"]]))
    inputs = (b["rgb_u8"].float(), lidar_ref, b["lane"], b["lane_num"], b["radar"], b["radar_adj"],
              b["target_point"], b["velocity"])
    oloss, opred, ograds = mmfn_oracle.train_step({k: v.clone() for k, v in sd.items()}, cfg,
                                                  dict(inputs=inputs, gt_waypoints=b["gt_waypoints"]))
    wp_l1 = (eng.last_pred.cpu() - opred).abs().mean().item()
    wp_max = (eng.last_pred.cpu() - opred).abs().max().item()
    cosine, rel = _grad_stats(model, ograds)
    rels = sorted(rel.values())
    worst_key = max(rel, key=rel.get)
    _report("parity_b16_tf32.json", dict(B=B, path="tf32", waypoint_l1=wp_l1, waypoint_max=wp_max, loss=loss,
                                         oracle_loss=oloss.item(), grad_cosine=cosine, grad_rel_median=rels[len(rels) // 2],
                                         grad_rel_p90=rels[int(0.9 * len(rels))], grad_rel_worst=rels[-1], worst_key=worst_key))
    assert wp_l1 < 1e-3, wp_l1                                   # north_star bar
    assert abs(loss - oloss.item()) < 1e-3
    # measured on B200 (profiles/r02_parity_b16_tf32.json): waypoint L1 2.5e-4, loss error 2e-5, cosine 0.992, median
    # per-tensor error 0.14, p90 0.23 -- the unmodified reference under cuBLAS / cuDNN TF32 sits at the same distance
    # from fp32 (tools/precision_yardstick.py, profiles/r02_precision_yardstick_b16.json)
    assert cosine > 0.985, cosine
    assert rels[len(rels) // 2] < 0.2 and rels[int(0.9 * len(rels))] < 0.35, (rels[len(rels) // 2], rels[int(0.9 * len(rels))], worst_key)


def test_loss_trajectory_tracks_oracle_over_20_steps(dev):
    """20 optimisation steps (BEV scatter + forward + backward + fused AdamW) of the production TF32 path against the
    oracle's AdamW on two alternating fixed batches (B=4, dropout off).  AdamW's sign-like early steps amplify tiny
    gradient differences element-wise, so the weights drift apart slowly; the LOSS trajectories must stay together."""
    from mmfn_b200.engine import TrainEngine
    B, steps = 4, 20
    cfg, model, sd, _ = _setup(dev, B, tf32=True)
    batches = [synthetic.synth_batch(B, first_index=i * B) for i in range(2)]
    eng = TrainEngine(model, lr=1e-4)
    osd = {k: v.clone() for k, v in sd.items()}
    opt = {"t": 0, "m": {}, "v": {}}
    dbs, oins = [], []
    for b in batches:
        dbs.append({k: v.to(dev) for k, v in b.items()})
        lidar_ref = torch.from_numpy(np.stack([bev_oracle.lidar_to_histogram_features(p[:, :3].numpy()) for p in b["points"]]))
        oins.append(dict(inputs=(b["rgb_u8"].float(), lidar_ref, b["lane"], b["lane_num"], b["radar"], b["radar_adj"],
                                 b["target_point"], b["velocity"]), gt_waypoints=b["gt_waypoints"]))
    mine, orac = [], []
    for i in range(steps):
        mine.append(eng.step(dbs[i % 2]).item())
        orac.append(mmfn_oracle.train_step(osd, cfg, oins[i % 2], opt_state=opt)[0].item())
    dev_rel = [abs(a - o) / max(abs(o), 1e-6) for a, o in zip(mine, orac)]
    _report("trajectory_tf32.json", dict(B=B, steps=steps, loss_gpu=mine, loss_oracle=orac, rel_dev=dev_rel))
    # measured on B200 over several runs (profiles/r02_trajectory_tf32.json): 6e-5 at step 1, 1-3 % over the first six
    # steps, 4-6 % worst over 20; run-to-run variation comes from the atomic accumulation order (AdamW's
    # lr * sign(g)-like early steps flip for elements whose gradient is within the TF32 noise of zero)
    assert dev_rel[0] < 1e-3, dev_rel[0]
    # (worst step over 20: 2.9 - 7.9 % over the runs of the last session, alone and inside the full suite; the two
    # trajectories re-converge: 0.5 - 0.8 % apart at steps 19 / 20)
    assert max(dev_rel[:6]) < 0.05 and max(dev_rel) < 0.15, (max(dev_rel), mine, orac)
    # both optimisers make the same progress on the two batches they keep seeing
    assert sum(mine[-2:]) < sum(mine[:2]) and sum(orac[-2:]) < sum(orac[:2])
    # (same 10 % as the per-step bound: the final pair measured 0.0 - 6.0 % apart over the runs of one day)
    assert abs(sum(mine[-2:]) - sum(orac[-2:])) < 0.10 * sum(orac[-2:])


@pytest.mark.parametrize("B,attn", [(2, "bf16"), (2, "tf32"), (32, "bf16")])
def test_bf16_step_matches_oracle(dev, B, attn):
    """BASELINE configs[2] numerics (bf16 tensor-core operands for the BasicBlock convolutions and the transformer
    linears, fp32 accumulate / statistics / residual stream / master weights) against the fp32 CPU oracle, at B=2 and
    at the configuration's per-GPU batch 32.  The reference itself under torch.autocast(bfloat16) drifts from its fp32
    output by 1.6e-3 mean / 4.9e-3 max waypoint L1: the bound asserted here is 5e-3, the
    measured figures are written to gpurun_out/parity_bf16_b*.json (committed under profiles/)."""
    from mmfn_b200 import ops
    from mmfn_b200.engine import TrainEngine
    try:
        cfg, model, sd, b = _setup(dev, B, tf32=True)
        ops.set_precision("bf16")
        ops.BF16_ATTN = attn == "bf16"             # attention core on the bf16 kernel, or the TF32 kernel on fp32 qkv
        eng = TrainEngine(model, lr=1e-4)
        db = {k: v.to(dev) for k, v in b.items()}
        loss = eng.forward_backward(db).item()
        torch.cuda.synchronize()
        lidar_ref = torch.from_numpy(np.stack([bev_oracle.lidar_to_histogram_features(p[:, :3].numpy()) for p in b["points"]]))
        inputs = (b["rgb_u8"].float(), lidar_ref, b["lane"], b["lane_num"], b["radar"], b["radar_adj"],
                  b["target_point"], b["velocity"])
        oloss, opred, ograds = mmfn_oracle.train_step({k: v.clone() for k, v in sd.items()}, cfg,
                                                      dict(inputs=inputs, gt_waypoints=b["gt_waypoints"]))
        wp_l1 = (eng.last_pred.cpu() - opred).abs().mean().item()
        wp_max = (eng.last_pred.cpu() - opred).abs().max().item()
        cosine, rel = _grad_stats(model, ograds)
        rels = sorted(rel.values())
        _report(f"parity_bf16_b{B}_attn{attn}.json", dict(B=B, path="bf16", attention_core=attn, waypoint_l1=wp_l1, waypoint_max=wp_max, loss=loss,
                                              oracle_loss=oloss.item(), grad_cosine=cosine, grad_rel_median=rels[len(rels) // 2],
                                              grad_rel_p90=rels[int(0.9 * len(rels))], grad_rel_worst=rels[-1],
                                              worst_key=max(rel, key=rel.get)))
        assert np.isfinite(loss) and torch.isfinite(eng.last_pred).all()
        assert wp_l1 < 5e-3, wp_l1
        assert abs(loss - oloss.item()) < 5e-3, (loss, oloss.item())
        assert cosine > 0.9, cosine
    finally:
        ops.set_precision("tf32")
        ops.BF16_ATTN = True


def test_bf16_graph_steps_keep_shadow_in_sync_and_train(dev):
    """The captured bf16 step: AdamW refreshes the bf16 weight shadow in the same pass (no convert kernel in the graph),
    load_state_dict refreshes it too, and the loss trajectory follows the TF32 path's on the same batch."""
    from mmfn_b200 import ops
    from mmfn_b200.engine import BatchStager, TrainEngine
    B = 4
    traj = {}
    try:
        for mode in ("tf32", "bf16"):
            cfg, model, sd, b = _setup(dev, B, tf32=True)
            ops.set_precision(mode)
            eng = TrainEngine(model, lr=1e-4)
            stager = BatchStager(b, dev)
            db = stager.stage(b)
            torch.cuda.synchronize()
            eng.capture(db, warmup=1)
            model.load_state_dict(sd)
            traj[mode] = [eng.step_graph().item() for _ in range(6)]
            torch.cuda.synchronize()
            if mode == "bf16":
                st = model.store
                assert torch.equal(st.flat16[: st.n_active].float(), st.flat[: st.n_active].to(torch.bfloat16).float())
        _report("trajectory_bf16_vs_tf32.json", traj)
        for a, c in zip(traj["tf32"], traj["bf16"]):      # measured: <= 3.4 % apart over six steps
            assert abs(a - c) < 0.08 * max(1.0, abs(a)), traj
        assert traj["bf16"][-1] < traj["bf16"][0]
    finally:
        ops.set_precision("tf32")


def test_loader_to_engine_path_matches_oracle(dev, tmp_path):
    """SURVEY 8(f) rank 1, end to end on the GPU: phase-1 pickles on disk -> PRE_Data (adds the radar adjacency) ->
    torch DataLoader with collate_single_cpu (ragged lane sets padded) -> to_engine_batch -> BatchStager (ONE packed H2D
    copy) -> TrainEngine step, against the oracle fed with the same collated tensors the way Engine.train builds them
    (phase2_train_net.py:66-103).  Exact-fp32 kernels."""
    import pickle
    from mmfn_b200 import data as mdata, ops
    from mmfn_b200.engine import BatchStager, TrainEngine
    B = 3
    cfg, model, sd, _ = _setup(dev, B, tf32=False)
    lanes = (70, 128, 93)
    for i in range(B):
        smp = synthetic.synth_sample(40 + i, bev_oracle.lidar_to_histogram_features, n_lanes=128)
        smp["vectormaps"] = [smp["vectormaps"][0][: lanes[i]]]
        with open(tmp_path / f"{i}.pkl", "wb") as f:
            pickle.dump(smp, f)
    ds = mdata.PRE_Data(str(tmp_path), cfg)
    loader = torch.utils.data.DataLoader(ds, batch_size=B, shuffle=False, num_workers=0, collate_fn=mdata.collate_single_cpu)
    data = next(iter(loader))
    eb = mdata.to_engine_batch(data, seq_len=cfg.seq_len, pad_lanes_to=128)
    assert eb["lane"].shape == (B, 128, 10, 5) and len(set(eb["lane_num"].tolist())) > 1 and "lidar" in eb   # ragged lane sets
    stager = BatchStager(eb, dev)
    db = stager.stage(eb)
    eng = TrainEngine(model, lr=1e-4)
    loss = eng.forward_backward(db).item()
    # the reference's own tensors (Engine.train): fronts / lidars lists, vectormaps = [[lane], [lane_num], Lmax]
    lane, lane_num, _ = data["vectormaps"][0]
    lane_p = torch.nn.functional.pad(lane.float(), (0, 0, 0, 0, 0, 128 - lane.shape[1]))
    inputs = (data["fronts"][0].float(), data["lidars"][0].float(), lane_p, lane_num, data["radar"][0].float(),
              data["radar_adj"].float(), torch.stack(data["target_point"], 1).float(), data["velocity"].float())
    gt = torch.stack([torch.stack(data["waypoints"][i], 1) for i in range(cfg.seq_len, len(data["waypoints"]))], 1).float()
    oloss, opred, ograds = mmfn_oracle.train_step({k: v.clone() for k, v in sd.items()}, cfg, dict(inputs=inputs, gt_waypoints=gt))
    assert (eng.last_pred.cpu() - opred).abs().mean().item() < 2e-4
    assert abs(loss - oloss.item()) < 2e-4
    cosine, _ = _grad_stats(model, ograds)
    assert cosine > 0.9995, cosine


def test_gpu_phase1_preprocessing_is_bit_exact(dev, golden_dir):
    """SURVEY 8(f) rank 2 on the GPU (csrc/loader.cu + csrc/bev.cu): y flip + float64 ego transform of raw sweeps
    (dataloader.py:229-239, :311-334), histogram (:271-293) and its uint8 packing, radar adjacency (:379-384) -- every
    output bit-identical to the CPU mirror (mmfn_b200/preprocess.py, itself pinned to the reference by
    tests/test_preprocess_cpu.py) on the seeded raw frames of tests/preprocess_fixture.py."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import preprocess_fixture as fx
    from mmfn_b200 import ops, preprocess as pp
    gold = np.load(os.path.join(golden_dir, "preprocess_golden.npz"))
    # (1) general two-frame transforms (r1 != r2) from the reference-generated fixture
    pts, poses = gold["tf_points"].astype(np.float32), gold["tf_poses"]
    raw = np.concatenate([pts, np.zeros((pts.shape[0], 1), np.float32)], 1)               # x y z intensity
    for p in poses:
        flipped = np.array(pts, dtype=np.float64)
        flipped[:, 1] *= -1
        want = pp.transform_points_2d(flipped, *p).astype(np.float32)
        got = ops.lidar_ego_transform(torch.from_numpy(raw[None]).to(dev), torch.from_numpy(pp.ego_pose_params(*p)[None]).to(dev))
        assert np.array_equal(got[0].cpu().numpy(), want)
    # (2) whole frames: ragged sweeps in one batch, NaN heading, histogram counts against the CPU mirror + oracle
    frames = [fx.raw_frame(f) for f in range(1, fx.N_FRAMES + 1)]
    for f, fr in enumerate(frames):
        fr["points"] = fr["points"][: 4096 - 301 * f]                                    # different sweep lengths
    xs, ys, th = ([fr["meas"][k] for fr in frames] for k in ("x", "y", "theta"))
    counts = pp.lidar_frames_to_bev_gpu([fr["points"] for fr in frames], xs, ys, th, dev)
    assert counts.dtype == torch.uint8 and tuple(counts.shape) == (len(frames), 2, 256, 256)
    for f, fr in enumerate(frames):
        t = 0.0 if np.isnan(th[f]) else th[f]
        p64 = np.array(fr["points"][:, :3], dtype=np.float64)
        p64[:, 1] *= -1
        ego = pp.transform_points_2d(p64, np.pi / 2 - t, -xs[f], -ys[f], np.pi / 2 - t, -xs[f], -ys[f]).astype(np.float32)
        hist = bev_oracle.lidar_to_histogram_features(ego)
        assert np.array_equal(counts[f].cpu().numpy(), np.rint(hist * 5).astype(np.uint8)), f
        assert np.array_equal(ops.bev_unpack_u8(counts[f: f + 1])[0].cpu().numpy(), hist), f
    h = ops.bev_unpack_u8(counts)
    assert torch.equal(ops.bev_pack_u8(h), counts)
    # (3) radar "adjacency" from float64 azimuths
    radar = np.stack([pp.radar_to_size(fr["radar"], (81, 5)) for fr in frames])
    adj = ops.radar_adjacency(torch.from_numpy(np.ascontiguousarray(radar[:, :, 1])).to(dev))
    assert np.array_equal(adj.cpu().numpy(), synthetic.radar_adjacency(radar).astype(np.float32))


def test_packed_loader_to_engine_path_matches_pickle_loader(dev, tmp_path):
    """packed shard -> PackedLoader -> BatchStager (ONE H2D copy of the packed batch: histogram as
    uint8 counts, no adjacency matrix) -> TrainEngine, against pickles -> PRE_Data -> collate -> to_engine_batch ->
    BatchStager -> TrainEngine on the same samples: the tensors the network receives are bit-identical, so are the
    prediction and the loss; the packed batch moves 2.3x fewer bytes."""
    import pickle
    from mmfn_b200 import data as mdata, ops, preprocess as pp
    from mmfn_b200.engine import BatchStager, TrainEngine
    B = 3
    cfg, model, sd, _ = _setup(dev, B, tf32=False)
    lanes = (70, 128, 93)
    smp = []
    for i in range(B):
        s = synthetic.synth_sample(40 + i, bev_oracle.lidar_to_histogram_features, n_lanes=128)
        s["vectormaps"] = [s["vectormaps"][0][: lanes[i]]]
        smp.append(s)
        with open(tmp_path / f"{i}.pkl", "wb") as f:
            pickle.dump(s, f)
    ds = mdata.PRE_Data(str(tmp_path), cfg)
    order = np.argsort([int(os.path.basename(p).split(".")[0]) for p in ds.preload_dict])
    eb = mdata.to_engine_batch(mdata.collate_single_cpu([ds[int(j)] for j in order]), seq_len=cfg.seq_len, pad_lanes_to=128)
    pp.write_packed(smp, str(tmp_path / "s.mmfnpack"))
    loader = mdata.PackedLoader([str(tmp_path / "s.mmfnpack")], batch_size=B, shuffle=False, pad_lanes_to=128)
    pb = next(iter(loader))
    st_ref, st_pk = BatchStager(eb, dev), BatchStager(pb, dev)
    assert st_pk.nbytes * 2 < st_ref.nbytes
    db_ref, db_pk = st_ref.stage(eb), st_pk.stage(pb)
    assert torch.equal(ops.bev_unpack_u8(db_pk["lidar_u8"]), db_ref["lidar"])
    assert torch.equal(ops.radar_adjacency(db_pk["radar_az64"]), db_ref["radar_adj"])
    for k in ("rgb_u8", "lane", "lane_num", "radar", "velocity", "target_point", "gt_waypoints"):
        assert torch.equal(db_pk[k], db_ref[k]), k
    # Trainer.validate (Engine.validate, phase2_train_net.py:124-183) takes both batch forms
    from mmfn_b200.trainer import Trainer
    tr = Trainer(model, lr=1e-4, pad_lanes_to=128)
    v_ref, v_pk = tr.validate([eb]), tr.validate(loader)
    assert v_ref > 0 and abs(v_ref - v_pk) <= 1e-6 * v_ref          # same tensors in; atomics order in the reductions
    eng = tr.engine
    loss_ref = eng.forward_backward(db_ref).item()
    pred_ref = eng.last_pred.clone()
    loss_pk = eng.forward_backward(db_pk).item()
    # same device tensors in, same kernels: equal up to the order of the fp64 BatchNorm-statistics atomics
    assert (eng.last_pred - pred_ref).abs().max().item() <= 1e-6 and abs(loss_pk - loss_ref) <= 1e-6


def test_torchvision_resnet34_weights_load_into_both_trunks(dev):
    """ImageCNN = models.resnet34(pretrained=True) minus fc (model_rad.py:22-23): a torchvision state_dict must land in
    encoder.image_encoder.features.* and encoder.img_map_encoder.features.* (KRSC storage behind the (K,C,R,S) view)."""
    import torchvision
    from mmfn_b200.model_rad import MMFN
    tv = torchvision.models.resnet34(weights=None)
    tsd = {k: torch.randn_like(v) if v.dtype.is_floating_point else v + 3 for k, v in tv.state_dict().items()}
    model = MMFN(GlobalConfig(), dev)
    before = model.state_dict()["encoder.lidar_encoder._model.layer1.0.conv1.weight"].clone()
    loaded = model.load_torchvision_resnet34(tsd)
    assert len(loaded) == 2 * (len(tsd) - 2)
    msd = model.state_dict()
    for trunk in ("image_encoder", "img_map_encoder"):
        for k in ("conv1.weight", "layer2.0.downsample.0.weight", "layer4.2.bn2.running_var", "layer3.5.conv2.weight",
                  "bn1.num_batches_tracked"):
            assert torch.equal(msd[f"encoder.{trunk}.features.{k}"].cpu(), tsd[k]), (trunk, k)
    w = model.store.p("encoder.image_encoder.features.layer1.0.conv1.weight")          # KRSC as the kernels read it
    assert torch.equal(w.cpu(), tsd["layer1.0.conv1.weight"].permute(0, 2, 3, 1))
    assert torch.equal(msd["encoder.lidar_encoder._model.layer1.0.conv1.weight"], before)   # other trunks untouched
    bad = dict(tsd)
    bad["conv1.weight"] = torch.zeros(64, 3, 3, 3)
    with pytest.raises(Exception):
        model.load_torchvision_resnet34(bad)


def test_two_gpu_data_parallel_step_matches_hand_summed_gradients(dev):
    """On-hardware check of the DDP contract (phase2_train_net.py:263-269) over NCCL, world size 2: the all-reduced
    gradient equals the hand-summed per-rank gradients (and their mean the oracle's replica average), parameters
    are bit-identical on both ranks after eager and CUDA-graph steps.  Needs 2 GPUs (skipped otherwise)."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    worker = os.path.join(root, "tests", "dp_gpu_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", worker],
                       capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DP_GPU_OK" in r.stdout, r.stdout[-2000:]


def test_state_dict_interchange(dev):
    from mmfn_b200.model_rad import MMFN
    keys = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "state_dict_keys.json")))
    model = MMFN(GlobalConfig(), dev)
    sd = model.state_dict()
    assert list(sd.keys()) == list(keys.keys())
    for k, (shape, dtype) in keys.items():
        assert list(sd[k].shape) == shape and str(sd[k].dtype) == dtype, k


def test_cuda_graph_replay_matches_eager_schedule(dev):
    """The captured forward+backward graph replays to the same losses as launching kernel by kernel,
    and dropout masks change between replays (device-resident seed offset)."""
    from mmfn_b200.engine import BatchStager, TrainEngine
    from mmfn_b200.model_rad import MMFN
    B = 2
    losses = {}
    for mode in ("eager", "graph"):
        cfg, model, sd, b = _setup(dev, B, tf32=True)
        eng = TrainEngine(model, lr=1e-4)
        stager = BatchStager(b, dev)
        db = stager.stage(b)
        torch.cuda.synchronize()
        if mode == "graph":
            model.load_state_dict(sd)                      # capture() ran warm-up steps: restart from the same weights
            eng.capture(db, warmup=1)
            model.load_state_dict(sd)
            eng.m.zero_(); eng.v.zero_(); eng.state.zero_()
        losses[mode] = [(eng.step_graph() if mode == "graph" else eng.step(db)).item() for _ in range(3)]
    # step 1 runs identical arithmetic up to atomic-accumulation order; later steps start from weights that
    # already differ by that noise times the AdamW sign step, so only a loose bound is meaningful
    for i, (a, g) in enumerate(zip(losses["eager"], losses["graph"])):
        assert abs(a - g) < (1e-3 if i == 0 else 3e-2), losses
    assert losses["graph"][2] < losses["graph"][0]        # and the replayed schedule does train
    # with the reference dropout (p = 0.1) two replays on identical weights/inputs must differ
    model = MMFN(GlobalConfig(), dev)
    model.load_state_dict(sd)
    eng = TrainEngine(model, lr=0.0, weight_decay=0.0)
    eng.capture(db, warmup=1)
    l1, l2 = eng.step_graph().item(), eng.step_graph().item()
    assert l1 != l2


def test_eval_forward_matches_oracle(dev):
    """Inference path (BN running stats, no dropout) -- what the e2e agents call."""
    B = 1
    cfg, model, sd, b = _setup(dev, B, tf32=False)
    model.eval()
    lidar_ref = torch.from_numpy(np.stack([bev_oracle.lidar_to_histogram_features(p[:, :3].numpy()) for p in b["points"]]))
    vectormaps = [[b["lane"].to(dev)], [b["lane_num"].to(dev).float()], b["lane"].shape[1]]
    with torch.no_grad():
        pred = model([b["rgb_u8"].to(dev).float()], [lidar_ref.to(dev)], None, vectormaps, [b["radar"].to(dev)],
                     [b["radar_adj"].to(dev)], b["target_point"].to(dev), b["velocity"].to(dev))
        opred = mmfn_oracle.forward({k: v.clone() for k, v in sd.items()}, cfg, b["rgb_u8"].float(), lidar_ref, b["lane"],
                                    b["lane_num"], b["radar"], b["radar_adj"], b["target_point"], b["velocity"], train=False)
    # golden running stats are deliberately far from the batch statistics, so eval-mode outputs
    # are O(1e3); compare relative to their magnitude
    rel = ((pred.cpu() - opred).abs().mean() / opred.abs().mean()).item()
    assert rel < 1e-5, rel
    steer, throttle, brake, meta = model.control_pid(pred, b["velocity"].to(dev))
    assert -1.0 <= steer <= 1.0 and 0.0 <= throttle <= 0.75


def _vectornet_setup(dev, B, L, P, tf32):
    from mmfn_b200 import ops
    from mmfn_b200.model_rad import MMFN
    ops.TF32 = tf32
    cfg = GlobalConfig(embd_pdrop=0.0, attn_pdrop=0.0, resid_pdrop=0.0)
    model = MMFN(cfg, dev)
    sd = synthetic.fill_golden_weights(model.state_dict(), 42)
    model.load_state_dict(sd)
    b = synthetic.synth_batch(B, first_index=300, n_lanes=L, n_nodes=P, n_points=0)
    return model, sd, b


@pytest.mark.parametrize("tf32", [False, True], ids=["fp32-simt", "tf32-tcgen05"])
def test_config5_vectornet_only_matches_oracle(dev, tf32):
    """BASELINE configs[4]: VectornetEncoder alone (256 polylines x 19 vector nodes), forward + backward,
    against the oracle's autograd on a batch the CPU finishes in seconds (B=4)."""
    from mmfn_b200.model_rad import _Aux
    B, L, P = 4, 256, 20
    model, sd, b = _vectornet_setup(dev, B, L, P, tf32)
    vn = model.net.vectornet
    model.store.flat_grad.zero_()
    out = vn.fwd(b["lane"].to(dev), b["lane_num"].to(dev))                  # (B, 64, 64, 64) NHWC
    dmap = torch.randn(B, 64, 64, 64, generator=torch.Generator().manual_seed(5))   # NCHW, like the oracle output
    vn.bwd(dmap.permute(0, 2, 3, 1).contiguous().to(dev))
    _Aux.join_all()
    torch.cuda.synchronize()

    pre = "encoder.vectornet_encoder."
    osd = {k: v.clone().requires_grad_(k.startswith(pre)) for k, v in sd.items() if k.startswith(pre)}
    oout = mmfn_oracle.vectornet(b["lane"], b["lane_num"], mmfn_oracle.Params(osd, pre))
    oout.backward(dmap)
    err = (out.permute(0, 3, 1, 2).cpu() - oout.detach()).abs().max().item() / oout.detach().abs().max().item()
    assert err < (1e-5 if not tf32 else 2e-3), err
    worst = 0.0
    for k, v in osd.items():
        if v.grad is None:                     # pos_emb MLP sees a constant zero input but still gets bias grads
            continue
        got = model.store.torch_view(k, grad=True)
        rel = (got.detach().cpu() - v.grad).norm().item() / max(v.grad.norm().item(), 1e-6)
        worst = max(worst, rel)
        assert rel < (1e-3 if not tf32 else 3e-2), (k, rel)     # measured 2.7e-4 (fp32: split-K atomics over 20k vectors)


@pytest.mark.parametrize("L,P,B", [(128, 10, 3), (256, 20, 2), (7, 10, 3), (5, 20, 1), (256, 20, 9)], ids=["ref-10-node", "bench-20-node", "ragged-10", "ragged-20", "many-tiles"])
def test_tensor_core_subgraph_forward_matches_simt_kernel(dev, L, P, B):
    """csrc/vectornet.cu subgraph_mma_fwd_kernel (TF32 mma.sync tiles, 256-row tiles of whole polylines) against the exact
    fp32 lane-per-node kernel on the same inputs: values within TF32 rounding (3e-3 of each tensor's scale), and the
    tensors it leaves for the backward are consistent among themselves EXACTLY: x[:, 64:] is the max over the polyline's
    nodes of x[:, :64], arg points at a node that attains it, tok = [m_2 | m_2]."""
    from mmfn_b200 import ops
    model, sd, b = _vectornet_setup(dev, B, L, P, True)
    vn = model.net.vectornet
    lane = b["lane"].to(dev)
    layers = [(lin.w, lin.b, ln.g, ln.b) for lin, ln in vn.sub]
    G, V = B * L, P - 1
    try:
        ops.TF32 = False
        ref = ops.subgraph_fused_fwd(lane, layers)
        ops.TF32 = True
        got = ops.subgraph_fused_fwd(lane, layers)
    finally:
        ops.TF32 = True

    def close(a, r, tol=3e-3):
        err, scale = (a - r).abs().max().item(), max(r.abs().max().item(), 1.0)
        assert err <= tol * scale, (err, scale)
    assert torch.equal(got["vec"], ref["vec"])
    for i in range(3):
        close(got["y"][i], ref["y"][i]); close(got["mean"][i], ref["mean"][i]); close(got["rstd"][i], ref["rstd"][i], 2e-2)
        assert (got["arg"][i] == ref["arg"][i]).float().mean().item() > 0.97          # near-ties may resolve differently
    close(got["x1"], ref["x1"]); close(got["x2"], ref["x2"]); close(got["tok"], ref["tok"])
    for x, arg in ((got["x1"], got["arg"][0]), (got["x2"], got["arg"][1])):
        h, m = x[:, :64].view(G, V, 64), x[:, 64:].view(G, V, 64)
        assert torch.equal(m, h.max(dim=1, keepdim=True).values.expand(-1, V, -1))
        assert torch.equal(h.gather(1, arg.long().view(G, 1, 64)).squeeze(1), m[:, 0])
    assert torch.equal(got["tok"][:, :64], got["tok"][:, 64:])
    assert torch.equal(got["argf"][:, :64], got["arg"][2]) and int(got["argf"][:, 64:].abs().max()) == 0


@pytest.mark.parametrize("L,P", [(128, 10), (256, 20), (7, 10), (5, 20)], ids=["ref-10-node", "bench-20-node", "ragged-10", "ragged-20"])
def test_fused_subgraph_forward_matches_unfused_kernels(dev, L, P):
    """csrc/vectornet.cu (one launch: vectorise + 3 x [Linear, LayerNorm, ReLU, max-pool, concat] + final max) against the
    per-layer kernels it replaces, exact-fp32 GEMMs on both sides: every tensor the backward consumes, the arg-max
    routing tables, and the parameter gradients of a full VectorNet forward+backward.  Polyline counts that
    are not a multiple of the polylines-per-warp packing (3 at 9 vectors, 1 at 19) exercise the ragged tail."""
    from mmfn_b200 import ops
    from mmfn_b200.model_rad import _Aux
    B = 3
    model, sd, b = _vectornet_setup(dev, B, L, P, False)
    vn = model.net.vectornet
    lane, num = b["lane"].to(dev), b["lane_num"].to(dev)
    dmap = torch.randn(B, 64, 64, 64, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
    res = {}
    try:
        for fused in (False, True):
            ops.FUSE_SUBGRAPH = fused
            model.store.flat_grad.zero_()
            out = vn.fwd(lane, num).clone()
            saved = dict(x=[lin.x.clone() for lin, _ in vn.sub], y=[ln.x.clone() for _, ln in vn.sub],
                         mean=[ln.mean.clone() for _, ln in vn.sub], rstd=[ln.rstd.clone() for _, ln in vn.sub],
                         arg=[a.clone() for a in vn.args], argf=vn.argf.clone())
            vn.bwd(dmap)
            _Aux.join_all()
            torch.cuda.synchronize()
            res[fused] = (out, saved, model.store.flat_grad.clone())
    finally:
        ops.FUSE_SUBGRAPH = True
    (o0, s0, g0), (o1, s1, g1) = res[False], res[True]
    for i in range(3):
        # identical routing except where two nodes tie to within the fp32 summation-order noise of the two GEMM orders
        assert (s0["arg"][i] != s1["arg"][i]).float().mean().item() < 1e-3, i
        for k in ("x", "y", "mean", "rstd"):
            assert torch.allclose(s0[k][i], s1[k][i], rtol=2e-5, atol=2e-5), (k, i, (s0[k][i] - s1[k][i]).abs().max().item())
    assert (s0["argf"] != s1["argf"]).float().mean().item() < 1e-3
    assert torch.allclose(o0, o1, rtol=1e-4, atol=1e-5)
    assert (g0 - g1).norm().item() <= 1e-4 * g0.norm().item()


def test_config5_vectornet_full_size_properties(dev):
    """Size-independent checks at BASELINE's size (B=128, 256 polylines x 19 nodes): samples are independent,
    padded lanes do not influence the result, gradients are finite."""
    from mmfn_b200.model_rad import _Aux
    B, L, P = 128, 256, 20
    model, sd, b = _vectornet_setup(dev, B, L, P, False)     # exact-fp32 kernels: the comparisons below are tight
    vn = model.net.vectornet
    lane, num = b["lane"].to(dev), b["lane_num"].to(dev)
    out = vn.fwd(lane, num).clone()
    vn.bwd(torch.ones_like(out))
    _Aux.join_all()
    assert torch.isfinite(out).all() and torch.isfinite(model.store.flat_grad).all()
    one = vn.fwd(lane[17:18].contiguous(), num[17:18].contiguous())
    assert torch.allclose(one[0], out[17], rtol=1e-4, atol=1e-5)
    # garbage in the padded lanes (index >= lane_num) must not change anything
    dirty = lane.clone()
    for i in range(B):
        dirty[i, int(num[i]):] = 123.0
    out2 = vn.fwd(dirty, num)
    assert torch.allclose(out2, out, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("tf32", [False, True], ids=["fp32-simt", "tf32-tcgen05"])
def test_config4_transfuser_matches_oracle_and_reference_goldens(dev, golden_dir, tf32):
    """BASELINE configs[3]: RGB+LiDAR only (map / radar branches off) = benchmarks/transfuser/model.py topology."""
    from mmfn_b200 import ops
    from mmfn_b200.transfuser import TransFuser
    from oracle import transfuser_oracle
    ops.TF32 = tf32
    cfg = GlobalConfig(embd_pdrop=0.0, attn_pdrop=0.0, resid_pdrop=0.0)
    model = TransFuser(cfg, dev)
    keys = json.load(open(os.path.join(golden_dir, "transfuser_state_dict_keys.json")))
    assert list(model.state_dict().keys()) == list(keys.keys())
    sd = synthetic.fill_golden_weights(model.state_dict(), 42)
    model.load_state_dict(sd)
    model.train()
    b = synthetic.synth_batch(2)
    wp_tol, loss_tol, grad_tol = (2e-4, 2e-4, 2e-2) if not tf32 else (1e-3, 1e-3, 0.6)
    lidar = ops.bev_scatter(b["points"].to(dev))
    pred = model([b["rgb_u8"].to(dev).float()], [lidar], b["target_point"].to(dev), b["velocity"].to(dev))
    loss = torch.nn.functional.l1_loss(pred, b["gt_waypoints"].to(dev), reduction="none").mean()
    loss.backward()
    gold = np.load(os.path.join(golden_dir, "transfuser_golden_b2.npz"))
    assert np.abs(pred.detach().cpu().numpy() - gold["pred_wp"]).mean() < wp_tol
    assert abs(loss.item() - float(gold["loss"])) < loss_tol
    osd = {k: v.clone() for k, v in sd.items()}
    batch = dict(inputs=(b["rgb_u8"].float(), lidar.cpu(), b["target_point"], b["velocity"]), gt_waypoints=b["gt_waypoints"])
    oloss, opred, ograds = transfuser_oracle.train_step(osd, cfg, batch)
    assert (pred.detach().cpu() - opred).abs().mean().item() < wp_tol
    dot = n1 = n2 = worst = 0.0
    for k, p in model.named_parameters():
        g, got = ograds[k], p.grad.detach().cpu()
        worst = max(worst, (got - g).norm().item() / max(g.norm().item(), 1e-6))
        dot += (got.double() * g.double()).sum().item()
        n1 += got.double().pow(2).sum().item()
        n2 += g.double().pow(2).sum().item()
    assert worst < grad_tol, worst
    assert dot / (n1 ** 0.5 * n2 ** 0.5) > (0.9999 if not tf32 else 0.97)
    # the engine path (BEV scatter + fwd + bwd + AdamW, CUDA graph capturable) runs on the same variant
    from mmfn_b200.engine import TrainEngine
    eng = TrainEngine(model, lr=1e-4)
    db = {k: b[k].to(dev) for k in ("rgb_u8", "points", "velocity", "target_point", "gt_waypoints")}
    model.load_state_dict(sd)
    l0 = eng.step(db).item()
    assert abs(l0 - oloss.item()) < loss_tol


@pytest.mark.parametrize("variant", ["vec", "img"])
def test_model_vec_and_model_img_variants_match_oracle_and_goldens(dev, golden_dir, variant):
    """The other two reference entry points (train.yaml:13-15): mmfn_b200.model_vec:MMFN and
    mmfn_b200.model_img:MMFN, production (TF32 tensor-core) path, reference-style call + loss.backward()."""
    import functools
    import importlib
    from mmfn_b200 import ops
    ops.TF32 = True
    cfg = GlobalConfig(embd_pdrop=0.0, attn_pdrop=0.0, resid_pdrop=0.0)
    model = importlib.import_module(f"mmfn_b200.model_{variant}").MMFN(cfg, dev)
    keys = json.load(open(os.path.join(golden_dir, f"{variant}_state_dict_keys.json")))
    assert list(model.state_dict().keys()) == list(keys.keys())
    sd = synthetic.fill_golden_weights(model.state_dict(), 42)
    model.load_state_dict(sd)
    model.train()
    B = 2
    b = synthetic.synth_batch(B)
    maps = synthetic.synth_map_images(B)
    lidar = ops.bev_scatter(b["points"].to(dev))
    vectormaps = [[b["lane"].to(dev)], [b["lane_num"].to(dev).float()], b["lane"].shape[1]]
    pred = model([b["rgb_u8"].to(dev).float()], [lidar], [maps.to(dev).float()], vectormaps, [b["radar"].to(dev)],
                 [b["radar_adj"].to(dev)], b["target_point"].to(dev), b["velocity"].to(dev))
    loss = torch.nn.functional.l1_loss(pred, b["gt_waypoints"].to(dev), reduction="none").mean()
    loss.backward()
    gold = np.load(os.path.join(golden_dir, f"{variant}_golden_b2.npz"))
    assert np.abs(pred.detach().cpu().numpy() - gold["pred_wp"]).mean() < 1e-3          # north_star bar
    assert abs(loss.item() - float(gold["loss"])) < 1e-3
    lane = maps.float() if variant == "img" else b["lane"]
    inputs = (b["rgb_u8"].float(), lidar.cpu(), lane, b["lane_num"], b["radar"], b["radar_adj"], b["target_point"], b["velocity"])
    oloss, opred, ograds = mmfn_oracle.train_step({k: v.clone() for k, v in sd.items()}, cfg,
                                                  dict(inputs=inputs, gt_waypoints=b["gt_waypoints"]),
                                                  forward_fn=functools.partial(mmfn_oracle.forward, variant=variant))
    dot = n1 = n2 = 0.0
    for k, p in model.named_parameters():
        g = ograds[k]
        if g is None:
            assert p.grad is None, k
            continue
        got = p.grad.detach().cpu()
        dot += (got.double() * g.double()).sum().item()
        n1 += got.double().pow(2).sum().item()
        n2 += g.double().pow(2).sum().item()
    assert dot / (n1 ** 0.5 * n2 ** 0.5) > 0.97
    # engine step on the same variant (packed batch -> BEV scatter -> fwd -> bwd -> AdamW)
    from mmfn_b200.engine import TrainEngine
    model.load_state_dict(sd)
    eng = TrainEngine(model, lr=1e-4)
    db = {k: v.to(dev) for k, v in b.items()}
    db["map_u8"] = maps.to(dev)
    assert abs(eng.step(db).item() - oloss.item()) < 1e-3


def test_inference_graph_matches_eager_eval(dev):
    """Batch-1 e2e-agent path: CUDA-graph replay of the eval forward == eager eval forward; raw points in."""
    import time
    from mmfn_b200.infer import InferenceEngine
    cfg, model, sd, b = _setup(dev, 1, tf32=True)
    frames = [synthetic.synth_batch(1, first_index=40 + i) for i in range(3)]
    eager = InferenceEngine(model, frames[0], use_graph=False)
    graph = InferenceEngine(model, frames[0], use_graph=True)
    for f in frames:
        pe = eager(f).clone()
        pg = graph(f).clone()
        # eval outputs are O(1e3) with the deliberately off-batch golden running stats; split-K atomics reorder sums
        assert torch.allclose(pe, pg, rtol=1e-3, atol=1e-2), (pe, pg)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(20):
        out = graph(frames[i % 3])
        out.cpu()
    ms = (time.perf_counter() - t0) / 20 * 1e3
    print(f"batch-1 inference, graph replay incl. H2D + read-back: {ms:.2f} ms/frame")
    steer, throttle, brake, meta = graph.run_step(frames[0])
    assert -1.0 <= steer <= 1.0 and 0.0 <= throttle <= 0.75


def test_trainer_validate_save_resume_roundtrip(dev, tmp_path):
    """Engine.validate / save / resume equivalents: reference file names, torch.optim.AdamW-format state."""
    from mmfn_b200.model_rad import MMFN
    from mmfn_b200.trainer import Trainer, optimizer_state_dict
    cfg, model, sd, b = _setup(dev, 2, tf32=True)
    tr = Trainer(model, lr=1e-4, logdir=str(tmp_path))
    batches = [synthetic.synth_batch(2, first_index=2 * i) for i in range(2)]
    l_train = tr.train(batches)
    l_val = tr.validate(batches[:1])
    assert l_train > 0 and l_val > 0 and model.training
    assert tr.save() is True
    for name in ("recent.log", "model.pth", "recent_optim.pth", "best_model.pth", "best_optim.pth"):
        assert (tmp_path / name).exists(), name
    log = json.load(open(tmp_path / "recent.log"))
    assert set(log) == {"epoch", "iter", "bestval", "bestval_epoch", "train_loss", "val_loss"} and log["iter"] == 2
    # the optimizer file loads into a real torch.optim.AdamW over the reference-ordered parameter list
    osd = torch.load(tmp_path / "best_optim.pth")
    cpu_params = [torch.nn.Parameter(p.detach().cpu().clone()) for p in model.parameters()]
    opt = torch.optim.AdamW(cpu_params, lr=1e-4)
    opt.load_state_dict(osd)
    n_state = len(opt.state_dict()["state"])
    assert n_state == len(cpu_params) - 21                      # the 21 never-used map-ResNet tensors have no state
    st0 = opt.state_dict()["state"][0]
    assert st0["exp_avg"].shape == cpu_params[0].shape and float(st0["step"]) == 2.0
    # resume into a fresh model + trainer: identical weights, moments and counters, and training continues identically
    model2 = MMFN(cfg, dev)
    tr2 = Trainer(model2, lr=1e-4, logdir=str(tmp_path))
    assert tr2.resume() is True
    assert tr2.cur_epoch == 1 and tr2.cur_iter == 2 and tr2.bestval == tr.bestval
    assert torch.equal(model2.store.flat[: model2.store.n_active], model.store.flat[: model.store.n_active])
    assert torch.equal(tr2.engine.m, tr.engine.m) and torch.equal(tr2.engine.v, tr.engine.v)
    assert torch.equal(tr2.engine.state, tr.engine.state)
    assert torch.equal(model2.store.flat_buf, model.store.flat_buf)


@pytest.mark.parametrize("prec", ["bf16", "tf32"])
@pytest.mark.parametrize("site,nmod", [(0, 3), (1, 3), (0, 2), (1, 2)], ids=["C64-T192", "C128-T192", "C64-T128", "C128-T128"])
@pytest.mark.parametrize("drop", [0.0, 0.1], ids=["nodrop", "drop"])
def test_whole_gpt_forward_kernel_matches_per_op_path(dev, prec, site, nmod, drop):
    """csrc/gpt_small.cu (all 8 blocks of a narrow fusion transformer in one cluster launch) against the per-op chain
    (LayerNorm, tcgen05 GEMMs, fused attention, ... = 7 launches per block) in the same precision: block outputs, every
    tensor saved for the backward, identical dropout masks (same counter-hash streams), and -- because the backward is the
    per-op one in both cases -- the parameter gradients of a full forward + backward."""
    from mmfn_b200 import ops
    from mmfn_b200.config import GlobalConfig
    from mmfn_b200.model_rad import MMFN, _Aux
    from mmfn_b200.transfuser import TransFuser
    if prec == "tf32" and site == 1:
        pytest.skip("n_embd 128 is fused in the bf16 configuration only (fp32 operand buffers do not fit)")
    ops.set_precision(prec)
    try:
        cfg = GlobalConfig(embd_pdrop=drop, attn_pdrop=drop, resid_pdrop=drop)
        model = (MMFN if nmod == 3 else TransFuser)(cfg, dev)
        sd = synthetic.fill_golden_weights(model.state_dict(), 42)
        model.load_state_dict(sd)
        if prec == "bf16":
            model.store.sync_shadow()
        gpt = model.net.gpts[site]
        B, C, T = 3, gpt.C, gpt.T
        gen = torch.Generator(device=dev).manual_seed(11)
        feats = [torch.randn(B, 16, 16, C, device=dev, generator=gen) for _ in range(nmod)]
        vel = torch.randn(B, 1, device=dev, generator=gen)
        dtok = torch.randn(B, T, C, device=dev, generator=gen)
        res = {}
        for fused in (0, 2):                       # 2: also n_embd 128 (off by default: no faster than the per-op chain)
            ops.FUSE_GPT = fused
            model.store.flat_grad.zero_()
            out = gpt.fwd(feats, vel, 1234, True).clone()
            saved = []
            for blk in gpt.blocks:
                saved.append(dict(x=blk.ln1.x.float().clone(), mean1=blk.ln1.mean.clone(), rstd1=blk.ln1.rstd.clone(),
                                  h1=blk.qkv.x.float().clone(), qkv=blk.qkv_out.float().clone(), P=blk.P.float().clone(),
                                  Pd=blk.Pd.float().clone(), y=blk.proj.x.float().clone(), x1=blk.ln2.x.float().clone(),
                                  h2=blk.fc1.x.float().clone(), a=blk.fc1.y.float().clone()))
            dfeats = [torch.zeros_like(f) for f in feats]
            gpt.bwd(dtok.clone(), dfeats)
            _Aux.join_all()
            torch.cuda.synchronize()
            res[fused] = (out, saved, model.store.flat_grad.clone(), [d.clone() for d in dfeats])
    finally:
        ops.FUSE_GPT = 1
        ops.set_precision("tf32")
    (o0, s0, g0, d0), (o1, s1, g1, d1) = res[0], res[2]
    tol = 3e-2 if prec == "bf16" else 4e-3          # relative to each tensor's scale; bf16 rounding compounds over 8 blocks
    for l, (a, b) in enumerate(zip(s0, s1)):
        for k in a:
            scale = a[k].abs().max().item() + 1e-6
            if k in ("P", "Pd"):
                # same dropout mask: an element is zero in one exactly where it is zero in the other (up to values that
                # round to zero), and the kept values agree
                assert ((a[k] == 0) != (b[k] == 0)).float().mean().item() < 1e-3, (l, k)
            err = (a[k] - b[k]).abs().max().item() / scale
            assert err < tol * (1 + l), (l, k, err)
    assert (o0 - o1).abs().max().item() / o0.abs().max().item() < tol * 8
    rel = (g0 - g1).norm().item() / g0.norm().item()
    assert rel < (0.1 if prec == "bf16" else 0.02), rel
    for a, b in zip(d0, d1):
        assert (a - b).norm().item() / a.norm().item() < (0.1 if prec == "bf16" else 0.02)
