#!/usr/bin/env python3
"""MMFN training-step benchmark (BASELINE.json metric: training samples/s).

  python bench.py --gpus N --steps K --warmup W            # native arm (this repo's CUDA path)
  python bench.py --impl reference --gpus N ...            # reference arm: CPU port of the reference step
  python bench.py --impl torch-eager [--dtype bf16] ...    # the UNMODIFIED reference nn.Module (baseline/_ref, see
                                                           # tools/install_ref.py) under stock PyTorch on the same GPU
  python bench.py --dtype bf16 --batch 32                  # BASELINE configs[2] (bf16 tensor-core operands)
  python bench.py --workload rgb_lidar --batch 64          # BASELINE configs[3] (RGB+LiDAR only)
  python bench.py --workload vectornet --batch 128         # BASELINE configs[4] (VectornetEncoder only)

N=1 default workload = BASELINE.json configs[1]: full MMFN (RGB + LiDAR + vector map + radar), forward +
backward + AdamW, per-GPU batch 16, fp32 storage / TF32 tensor-core math, dropout 0.1 as in the reference config.
For N>1 the driver launches this file under torch.distributed.run; per-GPU batch stays fixed (weak scaling) and
the only data-path collective is the NCCL all-reduce of the flat gradient buffer.

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM; `e2e`: same step through the
public API from pinned host buffers (one packed H2D copy + D2H loss read inside the timed region).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic FLOPs per sample, forward + backward, 2*MAC, matmul/conv/bmm only
FLOP_PER_SAMPLE = {"full": 117.85e9, "rgb_lidar": 69.28e9, "vectornet": 0.858e9}
DEFAULT_BATCH = {"full": 16, "rgb_lidar": 64, "vectornet": 128}
WORKLOAD_DESC = {
    "full": "full MMFN (RGB+LiDAR+map+radar) fwd+bwd+AdamW",
    "rgb_lidar": "RGB+LiDAR only (TransFuser topology, map/radar branches off) fwd+bwd+AdamW",
    "vectornet": "VectornetEncoder only (256 polylines x 19 vector nodes) fwd+bwd",
}
CONFIG_TAG = {("full", "tf32"): "BASELINE configs[1]", ("full", "bf16"): "BASELINE configs[2]",
              ("rgb_lidar", "tf32"): "BASELINE configs[3]", ("rgb_lidar", "bf16"): "BASELINE configs[3], bf16 operands",
              ("vectornet", "tf32"): "BASELINE configs[4]", ("vectornet", "bf16"): "BASELINE configs[4]"}


def dtype_note(dtype):
    return ("fp32 storage, tcgen05 kind::tf32 multiply (10-bit mantissa), fp32 accumulate" if dtype == "tf32" else
            "bf16 tensor-core operands (activations feeding MMAs + bf16 weight shadow), fp32 accumulate, fp32 master "
            "weights / statistics / residual stream / optimizer")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


def measure_matmul_peak(dev, tf32):
    """cuBLAS matmul throughput on this box (8192^3; burst = best of 10, sustained = 1 s back to back): the
    denominator for kind::tf32 kernels, which MEASURED_PEAKS.json does not hold (it has HBM copy + bf16 only)."""
    import torch
    n = 8192
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    dt = torch.float32 if tf32 else torch.bfloat16
    a, b = torch.randn(n, n, device=dev, dtype=dt), torch.randn(n, n, device=dev, dtype=dt)
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    reps = max(10, int(1000.0 / best))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        a @ b
    e1.record()
    torch.cuda.synchronize()
    sust = e0.elapsed_time(e1) / reps
    torch.backends.cuda.matmul.allow_tf32 = old
    fl = 2.0 * n ** 3
    return fl / (best * 1e-3) / 1e12, fl / (sust * 1e-3) / 1e12


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                self.samples.append((float(f[0]), float(f[1]), f[2:]))
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for _, _, fl in self.samples for n, v in zip(names, fl) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(s[0] for s in self.samples), "sm_max_mhz": self.samples[0][1],
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------- CPU arm
def _oracle_inputs(b, bev_oracle):
    import numpy as np
    import torch
    lidar = torch.from_numpy(np.stack([bev_oracle.lidar_to_histogram_features(p[:, :3].numpy()) for p in b["points"]]))
    return (b["rgb_u8"].float(), lidar, b["lane"], b["lane_num"], b["radar"], b["radar_adj"],
            b["target_point"], b["velocity"])


def cpu_port_throughput(batch, steps, warmup, threads, workload="full"):
    """The reference's training step restated on the CPU (oracle/mmfn_oracle.py, kind "port": the
    Python reference itself cannot travel to the GPU box).  Returns (samples/s, s/step)."""
    import torch
    from mmfn_b200 import synthetic
    from mmfn_b200.config import GlobalConfig
    from mmfn_b200.params import param_spec
    from oracle import bev_oracle, mmfn_oracle
    torch.set_num_threads(threads)
    cfg = GlobalConfig()
    variant = "transfuser" if workload == "rgb_lidar" else "rad"
    shapes = {k: torch.empty(s, dtype=torch.int64 if kind == "nbt" else torch.float32) for k, s, kind in param_spec(cfg, variant)}
    sd = synthetic.fill_golden_weights(shapes, 42)
    b = synthetic.synth_batch(batch)
    opt = {"t": 0, "m": {}, "v": {}}
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        inputs = _oracle_inputs(b, bev_oracle)
        if workload == "rgb_lidar":
            from oracle import transfuser_oracle
            transfuser_oracle.train_step(sd, cfg, dict(inputs=(inputs[0], inputs[1], inputs[6], inputs[7]),
                                                       gt_waypoints=b["gt_waypoints"]), opt_state=opt)
        else:
            mmfn_oracle.train_step(sd, cfg, dict(inputs=inputs, gt_waypoints=b["gt_waypoints"]), opt_state=opt)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return batch * len(times) / sum(times), sum(times) / len(times)


def cpu_vectornet_throughput(batch, steps, warmup, threads):
    """VectornetEncoder forward + backward (oracle/mmfn_oracle.vectornet + autograd) on the host cores."""
    import torch
    from mmfn_b200 import synthetic
    from mmfn_b200.config import GlobalConfig
    from mmfn_b200.params import param_spec
    from oracle import mmfn_oracle
    torch.set_num_threads(threads)
    cfg = GlobalConfig()
    pre = "encoder.vectornet_encoder."
    shapes = {k: torch.empty(s) for k, s, kind in param_spec(cfg) if k.startswith(pre)}
    sd = synthetic.fill_golden_weights(shapes, 42)
    b = synthetic.synth_batch(batch, first_index=300, n_lanes=256, n_nodes=20, n_points=0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        p = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        out = mmfn_oracle.vectornet(b["lane"], b["lane_num"], mmfn_oracle.Params(p, pre))
        out.square().mean().backward()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return batch * len(times) / sum(times), sum(times) / len(times)


def cpu_extras(threads):
    """SURVEY.md section 8(d) CPU figures beside the headline one: 1-thread step, B=1 forward + loss (configs[0]),
    lidar_to_histogram_features ms/frame (the numpy path the GPU BEV scatter replaces)."""
    import torch
    from mmfn_b200 import synthetic
    from mmfn_b200.config import GlobalConfig
    from mmfn_b200.params import param_spec
    from oracle import bev_oracle, mmfn_oracle
    out = {}
    sps1, _ = cpu_port_throughput(2, 1, 1, 1)
    out["one_thread"] = {"value": sps1, "unit": "samples/s", "cores": 1, "sample": "1 step of batch 2 after 1 warm-up, torch.set_num_threads(1)"}
    torch.set_num_threads(threads)
    cfg = GlobalConfig()
    shapes = {k: torch.empty(s, dtype=torch.int64 if kind == "nbt" else torch.float32) for k, s, kind in param_spec(cfg)}
    sd = synthetic.fill_golden_weights(shapes, 42)
    b = synthetic.synth_batch(1)
    ts = []
    for i in range(4):
        t0 = time.perf_counter()
        with torch.no_grad():
            pred = mmfn_oracle.forward(dict(sd), cfg, *_oracle_inputs(b, bev_oracle), train=True)
            mmfn_oracle.l1_loss(pred, b["gt_waypoints"])
        if i:
            ts.append(time.perf_counter() - t0)
    out["b1_forward_loss"] = {"value": 1.0 / statistics.median(ts), "unit": "samples/s", "ms": 1e3 * statistics.median(ts), "cores": threads,
                              "sample": "BASELINE configs[0]: B=1 forward + L1 loss (no_grad, train-mode BN), median of 3 after 1 warm-up"}
    pts = b["points"][0, :, :3].numpy()
    t0 = time.perf_counter()
    for _ in range(10):
        bev_oracle.lidar_to_histogram_features(pts)
    out["bev_histogram"] = {"ms_per_frame": 1e2 * (time.perf_counter() - t0), "cores": 1,
                            "sample": "numpy histogramdd restatement of lidar_to_histogram_features, 32768 points, mean of 10"}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    wl = args.workload
    if wl == "vectornet":
        sample_b = 8
        sps, sec = cpu_vectornet_throughput(sample_b, args.steps, args.warmup, threads)
        what = "oracle/mmfn_oracle.vectornet forward + autograd backward"
    else:
        sample_b = 4
        sps, sec = cpu_port_throughput(sample_b, args.steps, args.warmup, threads, wl)
        what = "oracle train_step incl. numpy BEV histogram"
    line = {
        "impl": "reference", "metric": "MMFN training samples/sec", "value": sps, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{WORKLOAD_DESC[wl]} ({CONFIG_TAG[(wl, 'tf32')]})",
                   "per_gpu_batch": args.batch, "sample": f"batch {sample_b} per step on host CPU, true fp32"},
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps of batch {sample_b} ({what})"},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------- torch-eager arm
def run_torch_eager(args):
    """The practical competitor on this box: the UNMODIFIED reference nn.Module
    (baseline/_ref/team_code/...; installed by tools/install_ref.py, git-ignored) on the B200 under stock PyTorch
    eager -- cuDNN / cuBLAS sm_100 kernels, autograd, torch.optim.AdamW -- driven by a restatement of the
    Engine.train loop body (run_steps/phase2_train_net.py:60-110; the script itself needs hydra + data on disk).
    --dtype tf32: allow_tf32 for matmul and cuDNN; --dtype bf16: torch.autocast(bfloat16) around forward + loss.
    None of this repo's kernels run on this arm; the BEV histogram is an input (the reference builds it offline)."""
    import types
    import numpy as np
    import torch
    import torch.nn.functional as F
    if int(os.environ.get("RANK", "0")) != 0:
        return
    ref = os.path.join(ROOT, "baseline", "_ref", "team_code")
    if not os.path.isdir(ref):
        print(json.dumps({"impl": "torch-eager", "unavailable": "baseline/_ref is empty: run tools/install_ref.py in the build container"}))
        return
    sys.path.insert(0, ref)
    sys.modules.setdefault("torch._six", types.SimpleNamespace(string_classes=(str, bytes)))
    import torchvision
    _orig34 = torchvision.models.resnet34
    torchvision.models.resnet34 = lambda pretrained=False, **k: _orig34(weights=None, **k)   # no network: random init
    from mmfn_b200 import synthetic
    from oracle import bev_oracle                       # numpy BEV histogram = the reference's offline preprocessing
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    wl, B = args.workload, args.batch
    if wl == "rgb_lidar":
        from benchmarks.transfuser.model import TransFuser
        from benchmarks.transfuser.config import GlobalConfig
        model = TransFuser(GlobalConfig(), dev).to(dev)
    else:
        from mmfn_utils.models import model_rad
        from mmfn_utils.datasets.config import GlobalConfig
        model = model_rad.MMFN(GlobalConfig(), dev).to(dev)
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    nbuf = 2
    host, resident = [], []
    for i in range(nbuf):
        b = synthetic.synth_batch(B, first_index=i * B)
        lidar = torch.from_numpy(np.stack([bev_oracle.lidar_to_histogram_features(p[:, :3].numpy()) for p in b["points"]]))
        hb = dict(fronts=b["rgb_u8"].float(), lidars=lidar, lane=b["lane"], lane_num=b["lane_num"].float(), radar=b["radar"],
                  radar_adj=b["radar_adj"], target_point=b["target_point"], velocity=b["velocity"], gt=b["gt_waypoints"])
        hb = {k: v.pin_memory() for k, v in hb.items()}
        host.append(hb)
        resident.append({k: v.to(dev) for k, v in hb.items()})
    autocast = args.dtype == "bf16"

    def step(d):
        for p in model.parameters():                                    # phase2_train_net.py:63-64
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            if wl == "rgb_lidar":
                pred = model([d["fronts"]], [d["lidars"]], d["target_point"], d["velocity"])
            else:
                vm = [[d["lane"]], [d["lane_num"]], d["lane"].shape[1]]
                pred = model([d["fronts"]], [d["lidars"]], None, vm, [d["radar"]], [d["radar_adj"]], d["target_point"], d["velocity"])
            loss = F.l1_loss(pred.float(), d["gt"], reduction="none").mean()    # :104
        loss.backward()                                                 # :108 (anomaly mode of :107 left OFF: generous to the reference)
        lv = loss.item()                                                # :109 host sync, as the reference
        opt.step()                                                      # :110
        return lv

    def timed(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    for i in range(max(3, args.warmup)):
        step(resident[i % nbuf])
    sampler = ClockSampler(0)
    sampler.start()
    ms = timed(lambda i: step(resident[i % nbuf]), args.steps)
    sampler.stop_flag = True
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())

    def step_e2e(i):                                                    # :78-97: the per-step .to(device) copies
        step({k: v.to(dev, non_blocking=True) for k, v in host[i % nbuf].items()})
    ms_e2e = timed(step_e2e, args.steps)
    sps, sps_e2e = B * args.steps / (ms * 1e-3), B * args.steps / (ms_e2e * 1e-3)
    peaks = measured_peaks()
    peak = peaks["tf_sust"] if autocast else measure_matmul_peak(dev, True)[1]
    line = {
        "impl": "torch-eager", "metric": "MMFN training samples/sec", "value": sps, "unit": "samples/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16-autocast" if autocast else "tf32", "data": "synthetic",
        "config": {"workload": f"{WORKLOAD_DESC[wl]} -- UNMODIFIED reference nn.Module, stock torch {torch.__version__} eager "
                               f"(cuDNN {torch.backends.cudnn.version()}), torch.optim.AdamW", "per_gpu_batch": B,
                   "dropout": 0.1, "anomaly_mode": False, "bev": "pre-built input (reference builds it offline)"},
        "model_tflops": sps * FLOP_PER_SAMPLE[wl] / 1e12,
        "roofline": {"bound": "tensor", "step": {"achieved": sps * FLOP_PER_SAMPLE[wl] / 1e12, "peak": peak, "unit": "TFLOP/s",
                                                 "frac": sps * FLOP_PER_SAMPLE[wl] / 1e12 / peak}},
        "e2e": {"value": sps_e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": 0, "clocks": sampler.summary(),
    }
    print(json.dumps(line), flush=True)


# ncu --set full captures of the headline shapes (tools/ncu_targets.py -> profiles/*_ncu_full_kernels.json)
NCU_FILES = ("r02_ncu_full_kernels.json", "r01_ncu_full_kernels.json")
NCU_KERNEL_OF = {
    ("conv2d_fwd_tf32", 16, 16, 256, 256, 3, 16): ("conv3x3_patch_kernel<64, 0, 1, 32>", "(4, 32, 1)"),
    ("conv2d_fwd_tf32", 16, 64, 64, 64, 3, 64): ("conv3x3_patch_kernel<64, 0, 0, 32>", "(1, 512, 1)"),
    ("gemm_tf32", 4096, 2048, 512, 1): ("tc_gemm_persist_kernel<GemmOp<0, 0, 128, 32>, 128, 1, 0>", "(148, 1, 1)"),
    ("conv2d_fwd_tf32", 16, 8, 512, 512, 3, 8): ("tc_kernel<ConvFwdOp<64, 0, 32>, 64, 4, 0, 0>", "(8, 8, 4)"),
    ("conv2d_fwd_bf16", 32, 16, 256, 256, 3, 16): ("conv3x3_patch_kernel<128, 0, 1, 64>", "(2, 64, 1)"),
    ("conv2d_fwd_bf16", 32, 64, 64, 64, 3, 64): ("conv3x3_patch_kernel<64, 0, 0, 64>", "(1, 1024, 1)"),
    ("gemm_bf16", 8192, 2048, 512, 1): ("tc_gemm_persist_kernel<GemmOp<0, 0, 128, 64>, 128, 1, 1>", "(148, 1, 1)"),
    # round-1 names (profiles/r01_ncu_full_kernels.json)
    ("r01", "conv2d_fwd_tf32", 16, 16, 256, 256, 3, 16): ("conv3x3_patch_kernel<64, 0, 1>", "(4, 32, 1)"),
}


def ncu_traffic(dom):
    """dram__bytes_read + dram__bytes_write per launch of the dominant shape from the committed ncu capture."""
    want = NCU_KERNEL_OF.get(tuple(dom["key"]))
    if not want:
        return None
    for fname in NCU_FILES:
        path = os.path.join(ROOT, "profiles", fname)
        if not os.path.exists(path):
            continue
        for rec in json.load(open(path)):
            if rec["kernel"] == want[0] and rec.get("grid") == want[1]:
                return rec.get("traffic_bytes")
    return None


def _graph_chain_us(fn, chain=40, replays=10):
    """average duration of `fn` inside a CUDA-graph chain of back-to-back launches (CUDA events on the replay stream)"""
    import torch
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(chain):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / (replays * chain)


def dominant_kernel_chain(ops, L, dev, prof_steps, bf16):
    """Pick the tensor-core launch shape with the largest eager total and time it as a graph chain."""
    import torch
    conv_fwd, conv_dg = ("conv2d_fwd_bf16", "conv2d_dgrad_bf16") if bf16 else ("conv2d_fwd_tf32", "conv2d_dgrad_tf32")
    gemm_fn = "gemm_bf16" if bf16 else "gemm_tf32"
    dt = torch.bfloat16 if bf16 else torch.float32
    es = 2.0 if bf16 else 4.0
    # the stride-1 3x3 data gradient runs the SAME kernel as the forward convolution (mirrored taps): count them together
    merged = {}
    for key, n, ms, flops in L.last_shapes:
        if key[0] == conv_dg and key[3] == key[4] and key[2] == key[6]:
            key = (conv_fwd,) + tuple(key[1:])
        if key[0] not in (conv_fwd, gemm_fn) or (key[0] == gemm_fn and key[4] != 1):
            continue
        m = merged.setdefault(tuple(key), [0, 0.0, flops])
        m[0] += n
        m[1] += ms
    key, (n, _, flops) = max(merged.items(), key=lambda kv: kv[1][1])
    if key[0] == conv_fwd:
        _, N, H, C, Co, R, Ho = key
        stride = H // Ho
        x = torch.randn(N, H, H, C, device=dev).to(dt)
        w = (torch.randn(Co, R, R, C, device=dev) * 0.05).to(dt)
        fn = lambda: ops.conv2d_fwd(x, w, stride, R // 2)
        name = f"{conv_fwd} N{N} {H}x{H} C{C}->{Co} {R}x{R} s{stride} (implicit GEMM {N * Ho * Ho}x{Co}x{R * R * C})"
        nbytes = es * (N * H * H * C + Co * R * R * C) + 4.0 * N * Ho * Ho * Co
    else:
        _, M, Nn, K, _ = key
        A, Bm = torch.randn(M, K, device=dev).to(dt), (torch.randn(Nn, K, device=dev) * 0.02).to(dt)
        Cm = torch.empty(M, Nn, device=dev)
        fn = lambda: ops.gemm(A, Bm, Cm)
        name = f"{gemm_fn} {M}x{Nn}x{K}"
        nbytes = es * (M * K + Nn * K) + 4.0 * M * Nn
    us = _graph_chain_us(fn)
    L.next_work = None
    return dict(key=list(key), name=name, n=n // prof_steps, flops=flops, bytes=nbytes, us=us)


# ------------------------------------------------------------------------------------- native arm
def _dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "the native arm needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, local, dev


def _timed(fn, steps, world, dev):
    """K steps bracketed by barrier + synchronize on both sides, CUDA events, max over ranks -> ms"""
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item()


def measure_training(args, workload, dtype, B, steps, warmup, want_profile=True):
    """One native measurement: returns a dict with the timed numbers (+ eager per-call profile on rank 0)."""
    import torch
    from mmfn_b200 import ops, synthetic
    from mmfn_b200._lib import lib
    from mmfn_b200.config import GlobalConfig
    from mmfn_b200.engine import BatchStager, TrainEngine
    world, rank, local, dev = _dist_setup()
    ops.set_precision(dtype)
    cfg = GlobalConfig()                                   # reference config: dropout 0.1 everywhere
    if workload == "rgb_lidar":
        from mmfn_b200.transfuser import TransFuser
        model = TransFuser(cfg, dev)
    else:
        from mmfn_b200.model_rad import MMFN
        model = MMFN(cfg, dev)
    eng = TrainEngine(model, lr=1e-4)
    eng.broadcast_parameters()

    # distinct synthetic batches per rank (disjoint shards, like DistributedSampler)
    nbuf = 2
    keep = ("rgb_u8", "points", "velocity", "target_point", "gt_waypoints") if workload == "rgb_lidar" else None
    host_batches = []
    for i in range(nbuf):
        hb = synthetic.synth_batch(B, first_index=(rank * nbuf + i) * B)
        host_batches.append({k: v for k, v in hb.items() if keep is None or k in keep})
    stager = BatchStager(host_batches[0], dev)
    packed = []                                            # the same batches, packed, resident in HBM
    for hb in host_batches:
        stager.stage(hb)
        torch.cuda.synchronize()
        packed.append(stager.dev.clone())
    use_graph = not args.no_graph
    if use_graph:
        eng.capture(stager.dev_views)                      # fwd+bwd schedule -> CUDA graphs

    def run_step():
        return eng.step_graph() if use_graph else eng.step(stager.dev_views)

    def step_resident(i):
        stager.dev.copy_(packed[i % nbuf])                 # device-to-device: inputs already in HBM
        run_step()

    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    # e2e feed: the loader side (collate + pin, DataLoader workers in the reference) leaves packed batches in
    # pinned host memory; every step moves its batch pinned-host -> device with ONE copy (issued on a copy
    # stream while the previous step computes), runs the step and reads the loss back.
    for slot, hb in enumerate(host_batches):
        stager.pack(hb, slot)

    def step_e2e(i):
        stager.commit()                                     # batch i (H2D issued during step i-1) -> graph input buffer
        loss = run_step()
        stager.prefetch((i + 1) % nbuf)                     # H2D of batch i+1 overlaps this step's kernels
        loss_host.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()           # the caller reads the loss every step (phase2:109)

    for i in range(warmup):
        step_resident(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib().launches
    ms = _timed(step_resident, steps, world, dev)
    launches = lib().launches - l0
    sampler.stop_flag = True
    stager.prefetch(0)
    for i in range(2):
        step_e2e(i)
    ms_e2e = _timed(lambda i: step_e2e(i + 2), steps, world, dev)
    res = dict(world=world, rank=rank, dev=dev, B=B, ms=ms, ms_e2e=ms_e2e, launches=launches, steps=steps, warmup=warmup,
               h2d=stager.nbytes, clocks=sampler.summary() if rank == 0 else None, use_graph=use_graph,
               loss=float(loss_host.item()))
    # roofline leg: per-call CUDA-event timing of every C-ABI launch over `prof_steps` live steps (rank 0)
    if rank == 0 and want_profile:
        prof_steps = min(steps, 3)
        torch.cuda.synchronize()
        lib().start_profile()
        for i in range(prof_steps):
            stager.dev.copy_(packed[i % nbuf])
            eng.forward_backward(stager.dev_views)         # eager schedule: one event pair per C-ABI call
            eng.optimizer_step(collective=False)           # rank 0 only: must not enter a collective here
        torch.cuda.synchronize()
        prof = lib().stop_profile()
        for d in prof.values():
            d["ms_per_step"] = d["ms"] / prof_steps
        res.update(prof=prof, prof_steps=prof_steps, classes=lib().last_classes,
                   top_shapes=[dict(call=list(k), n=n // prof_steps, ms_per_step=round(t / prof_steps, 3),
                                    tflops=round(fl * n / (t * 1e-3) / 1e12, 1) if t > 0 else None)
                               for k, n, t, fl in lib().last_shapes[:14]])
        res["dom"] = dominant_kernel_chain(ops, lib(), dev, prof_steps, dtype == "bf16")
    # free this configuration's graphs / buffers before a possible second measurement
    eng._graph = None
    del eng, model, stager, packed
    torch.cuda.empty_cache()
    return res


def build_roofline(res, workload, dtype, peaks):
    """Honest roofline: ONE dominant launch shape (graph-chain timing) + the whole step + per-class time shares."""
    dom, prof, steps = res["dom"], res["prof"], res["prof_steps"]
    step_ms = res["ms"] / res["steps"]
    if dtype == "bf16":
        peak, peak_src = peaks["tf_sust"], peaks["src"] + " (MEASURED_PEAKS.json sustained bf16 matmul)"
    else:
        peak, peak_src = res["tf32_peak"][1], "measured live: torch.matmul fp32 8192^3 with allow_tf32, sustained 1 s (burst %.0f)" % res["tf32_peak"][0]
    achieved = dom["flops"] / (dom["us"] * 1e-6) / 1e12
    total_ms = sum(d["ms"] for d in prof.values())
    sps = res["world"] * res["B"] * res["steps"] / (res["ms"] * 1e-3)
    model_tf = sps / res["world"] * FLOP_PER_SAMPLE[workload] / 1e12          # per GPU
    classes = {}
    for name, c in sorted(res["classes"].items(), key=lambda kv: -kv[1]["ms"]):
        e = {"launches_per_step": c["calls"] // steps, "eager_ms_per_step": round(c["ms"] / steps, 3),
             "time_share": round(c["ms"] / total_ms, 3)}
        if name != "hbm" and c["ms"] > 0:
            tf = c["flops"] / (c["ms"] * 1e-3) / 1e12
            e.update(bound="tensor", achieved_tflops=round(tf, 1), frac=round(tf / peak, 3))
        else:
            e.update(bound="hbm")
        classes[name] = e
    return {"bound": "tensor", "kernel": dom["name"], "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "traffic": ncu_traffic(dom), "algorithmic_bytes": dom["bytes"],
            "algorithmic_flops": dom["flops"], "peak_source": peak_src,
            "avg_launch_ms": dom["us"] * 1e-3, "launches_per_step": dom["n"],
            "share_of_step": dom["n"] * dom["us"] * 1e-3 / step_ms,
            "timing": "CUDA events around a CUDA-graph chain of 40 back-to-back launches of this shape, 10 replays",
            "step": {"achieved": model_tf, "peak": peak, "unit": "TFLOP/s", "frac": model_tf / peak,
                     "note": "whole step per GPU: SURVEY 8(d) model FLOPs / measured ms_per_step -- no single kernel dominates"},
            "classes": classes,
            "classes_note": "eager per-call CUDA-event totals (include launch gaps): shares rank the classes, the graph step overlaps them",
            "top_shapes": res["top_shapes"],
            "eager_ms_per_step_by_entry_point": {k: round(v["ms_per_step"], 3) for k, v in
                                                 sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:8]},
            "eager_total_ms_per_step": total_ms / steps}


def run_native(args):
    import torch.distributed as dist
    if args.workload == "vectornet":
        return run_vectornet(args)
    wl, dtype, B = args.workload, args.dtype, args.batch
    res = measure_training(args, wl, dtype, B, args.steps, args.warmup)
    world, rank, dev = res["world"], res["rank"], res["dev"]
    extra = None
    if args.extra:
        # BASELINE configs[2] (bf16 operands, per-GPU batch 32) measured in the same launch and attached to the line, so
        # that the driver's 1/2/4/8-GPU runs also carry the configuration north_star quotes the 8-GPU target on.
        # Every rank takes the same path (collectives inside); a failure here must not cost the headline line.
        try:
            ex = measure_training(args, "full", "bf16", 32, max(3, args.steps // 2), 3, want_profile=False)
        except Exception as exc:                                      # noqa: BLE001
            ex = None
            extra = {"config": "BASELINE configs[2]", "error": repr(exc)[:300]}
        if rank == 0 and ex is not None:
            sps2 = world * 32 * ex["steps"] / (ex["ms"] * 1e-3)
            extra = {"config": "BASELINE configs[2]: full MMFN, bf16 tensor-core operands, per-GPU batch 32, dp%d" % world,
                     "value": sps2, "unit": "samples/s", "ms_per_step": ex["ms"] / ex["steps"], "steps": ex["steps"], "dtype": "bf16",
                     "e2e": {"value": world * 32 * ex["steps"] / (ex["ms_e2e"] * 1e-3), "unit": "samples/s",
                             "h2d_bytes_per_step": ex["h2d"], "d2h_bytes_per_step": 4},
                     "gpu_launches": ex["launches"], "loss": ex["loss"]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    if dtype != "bf16":
        res["tf32_peak"] = measure_matmul_peak(dev, True)
    roofline = build_roofline(res, wl, dtype, peaks)
    sps = world * B * res["steps"] / (res["ms"] * 1e-3)
    sps_e2e = world * B * res["steps"] / (res["ms_e2e"] * 1e-3)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        c_sps, _ = cpu_port_throughput(4, 3, 1, threads, wl)
        cpu = {"value": c_sps, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": "3 steps of batch 4 after 1 warm-up (oracle train_step incl. numpy BEV histogram), true fp32"}
        if wl == "full":
            cpu.update(cpu_extras(threads))
    frame = "256x256 crop of 400x300 RGB + 32768-pt LiDAR" + ("" if wl == "rgb_lidar" else " + 128 lanes x 10 nodes + 81 radar pts")
    line = {
        "metric": "MMFN training samples/sec", "value": sps, "unit": "samples/s", "n_gpus": world,
        "steps": res["steps"], "warmup": res["warmup"], "ms_per_step": res["ms"] / res["steps"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {"workload": f"{WORKLOAD_DESC[wl]}, {dtype} ({CONFIG_TAG[(wl, dtype)]})", "arithmetic": dtype_note(dtype),
                   "per_gpu_batch": B, "global_batch": B * world, "frame": frame,
                   "dropout": 0.1, "parallelism": f"dp{world}", "cuda_graph": res["use_graph"],
                   "l2": "per-step working set (activations + 420 MB params) >> 126 MB L2; inputs rotate between 2 batches"},
        "model_tflops": sps * FLOP_PER_SAMPLE[wl] / 1e12,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": {"value": sps_e2e, "unit": "samples/s", "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": 4,
                "ms_per_step": res["ms_e2e"] / res["steps"]},
        "gpu_launches": res["launches"],
        "clocks": res["clocks"],
    }
    if extra is not None:
        line["configs2_bf16_b32"] = extra
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_vectornet(args):
    """BASELINE configs[4]: VectornetEncoder alone (model_rad.py:249-417), 256 polylines x 19 vector nodes, forward +
    backward (parameter gradients included), per-GPU batch 128.  HBM-bound: roofline against SURVEY 8(d)'s algorithmic
    bytes -- Subgraph B*L*V*7*4 in + B*L*128*4 out, generator 64*262144*4 weights + B*262144*4 out."""
    import torch
    import torch.distributed as dist
    from mmfn_b200 import ops, synthetic
    from mmfn_b200._lib import lib
    from mmfn_b200.config import GlobalConfig
    from mmfn_b200.model_rad import MMFN, _Aux
    world, rank, local, dev = _dist_setup()
    ops.set_precision("tf32")
    B, L, P = args.batch, 256, 20
    V = P - 1
    model = MMFN(GlobalConfig(), dev)
    vn = model.net.vectornet
    nbuf = 2
    host = [synthetic.synth_batch(B, first_index=300 + (rank * nbuf + i) * B, n_lanes=L, n_nodes=P, n_points=0) for i in range(nbuf)]
    pinned = [(h["lane"].pin_memory(), h["lane_num"].pin_memory()) for h in host]
    resident = [(a.to(dev), b.to(dev)) for a, b in pinned]
    lane_s, num_s = torch.empty_like(resident[0][0]), torch.empty_like(resident[0][1])
    dmap = torch.randn(B, 64, 64, 64, device=dev) * 1e-3
    check = torch.zeros((), device=dev)

    def body():
        model.store.flat_grad.zero_()
        out = vn.fwd(lane_s, num_s)
        check.copy_(out.view(-1)[:1024].sum())
        vn.bwd(dmap)
        _Aux.join_all()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            body()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    use_graph = not args.no_graph
    per_step = None
    if use_graph:
        l0 = lib().launches
        g = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream(priority=-1)
        cap.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cap):
            g.capture_begin()
            body()
            g.capture_end()
        torch.cuda.current_stream().wait_stream(cap)
        torch.cuda.synchronize()
        per_step = lib().launches - l0
    run = (lambda: g.replay()) if use_graph else body

    def step_resident(i):
        lane_s.copy_(resident[i % nbuf][0]); num_s.copy_(resident[i % nbuf][1])
        run()
    check_host = torch.empty((), dtype=torch.float32).pin_memory()

    def step_e2e(i):
        lane_s.copy_(pinned[i % nbuf][0], non_blocking=True); num_s.copy_(pinned[i % nbuf][1], non_blocking=True)
        run()
        check_host.copy_(check, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib().launches
    ms = _timed(step_resident, args.steps, world, dev)
    launches = (per_step * args.steps) if use_graph else lib().launches - l0
    sampler.stop_flag = True
    ms_e2e = _timed(step_e2e, args.steps, world, dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # per-part timing (eager, per-call events) for the two HBM rooflines SURVEY 8(d) names
    lane_s.copy_(resident[0][0]); num_s.copy_(resident[0][1])
    torch.cuda.synchronize()
    lib().start_profile()
    for _ in range(3):
        body()
    torch.cuda.synchronize()
    prof = lib().stop_profile()
    peaks = measured_peaks()
    G = B * L
    sub_bytes = G * V * 7 * 4.0 + G * 128 * 4.0
    gen_bytes = 64 * 262144 * 4.0 + B * 262144 * 4.0
    # the generator forward GEMM (B x 262144 x 64) timed alone as a graph chain
    x64 = torch.randn(B, 64, device=dev)
    yb = torch.empty(B, 262144, device=dev)
    us_gen = _graph_chain_us(lambda: ops.gemm(x64, vn.g3.w, yb, bias=vn.g3.b), chain=10)
    sub_fn = getattr(vn, "subgraph_fwd", None)
    us_sub = _graph_chain_us(lambda: sub_fn(lane_s), chain=10) if sub_fn is not None else None
    sps = world * B * args.steps / (ms * 1e-3)
    roofline = {"bound": "hbm", "kernel": "generator Linear(64 -> 262144) forward GEMM (B x 262144 x 64)",
                "achieved": gen_bytes / (us_gen * 1e-6) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                "frac": gen_bytes / (us_gen * 1e-6) / 1e9 / peaks["hbm"], "traffic": None,
                "algorithmic_bytes": gen_bytes, "avg_launch_ms": us_gen * 1e-3, "peak_source": peaks["src"],
                "timing": "CUDA events around a CUDA-graph chain of 10 back-to-back launches, 10 replays",
                "step": {"achieved": sps / world * FLOP_PER_SAMPLE["vectornet"] / 1e12, "unit": "TFLOP/s",
                         "note": "0.858 GFLOP/sample: the path is HBM-bound (67 MB generator weights + gradients + 134 MB map out/in per pass)"},
                "eager_ms_per_step_by_entry_point": {k: round(v["ms"] / 3, 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:10]}}
    if us_sub is not None:
        roofline["subgraph"] = {"kernel": "polyline Subgraph forward (lane -> vectors -> 3 x [Linear, LN, ReLU, max, concat] -> max)",
                                "bound": "hbm", "achieved": sub_bytes / (us_sub * 1e-6) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                                "frac": sub_bytes / (us_sub * 1e-6) / 1e9 / peaks["hbm"], "algorithmic_bytes": sub_bytes,
                                "avg_launch_ms": us_sub * 1e-3}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        c_sps, _ = cpu_vectornet_throughput(8, 3, 1, threads)
        cpu = {"value": c_sps, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": "3 passes of batch 8 after 1 warm-up (oracle/mmfn_oracle.vectornet forward + autograd backward)"}
    line = {
        "metric": "MMFN training samples/sec", "value": sps, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32", "data": "synthetic",
        "config": {"workload": f"{WORKLOAD_DESC['vectornet']} ({CONFIG_TAG[('vectornet', 'tf32')]})", "per_gpu_batch": B,
                   "global_batch": B * world, "polylines": L, "vector_nodes": V, "parallelism": f"dp{world} (replicas, no exchange: no optimizer in this microbench)",
                   "cuda_graph": use_graph, "l2": "per-pass working set (67 MB generator weights + 2 x 134 MB map tensors + subgraph activations) > 126 MB L2"},
        "model_tflops": sps * FLOP_PER_SAMPLE["vectornet"] / 1e12,
        "roofline": roofline, "cpu_baseline": cpu,
        "e2e": {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": "samples/s",
                "h2d_bytes_per_step": pinned[0][0].numel() * 4 + pinned[0][1].numel() * 4, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "clocks": sampler.summary(),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference", "torch-eager"])
    ap.add_argument("--workload", default="full", choices=["full", "rgb_lidar", "vectornet"])
    ap.add_argument("--dtype", default="tf32", choices=["tf32", "bf16"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the BASELINE config's: 16 / 32 bf16 / 64 / 128)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", dest="extra", action="store_false",
                    help="skip the attached BASELINE configs[2] (bf16, batch 32) measurement of the default run")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.batch <= 0:
        args.batch = 32 if (args.workload == "full" and args.dtype == "bf16") else DEFAULT_BATCH[args.workload]
    if args.workload != "full" or args.dtype == "bf16":
        args.extra = False
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:      # convenience: self-launch one rank per GPU
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                                   "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:])
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch-eager":
        run_torch_eager(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
