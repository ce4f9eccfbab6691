#!/usr/bin/env python3
"""MMFN training-step benchmark (BASELINE.json metric: training samples/s).

  python bench.py --gpus N --steps K --warmup W            # native arm (this repo's CUDA path)
  python bench.py --impl reference --gpus N ...            # reference arm: CPU port of the reference step

N=1 workload = BASELINE.json configs[1]: full MMFN (RGB + LiDAR + vector map + radar), forward +
backward + AdamW, per-GPU batch 16, fp32, dropout 0.1 as in the reference config.  For N>1 the
driver launches this file under torch.distributed.run; per-GPU batch stays 16 (weak scaling) and
the only data-path collective is ONE NCCL all-reduce of the flat gradient buffer per step.

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

FLOP_PER_SAMPLE_FWD_BWD = 117.85e9      # SURVEY.md section 8(d), 2*MAC, matmul/conv/bmm only
PER_GPU_BATCH = 16


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


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
def cpu_port_throughput(batch, steps, warmup, threads):
    """The reference's training step restated on the CPU (oracle/mmfn_oracle.py, kind "port": the
    Python reference itself cannot travel to the GPU box).  Returns samples/s."""
    import numpy as np
    import torch
    from mmfn_b200 import synthetic
    from mmfn_b200.config import GlobalConfig
    from mmfn_b200.params import param_spec
    from oracle import bev_oracle, mmfn_oracle
    torch.set_num_threads(threads)
    cfg = GlobalConfig()
    shapes = {k: torch.empty(s, dtype=torch.int64 if kind == "nbt" else torch.float32) for k, s, kind in param_spec(cfg)}
    sd = synthetic.fill_golden_weights(shapes, 42)
    b = synthetic.synth_batch(batch)
    opt = {"t": 0, "m": {}, "v": {}}
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        lidar = torch.from_numpy(np.stack([bev_oracle.lidar_to_histogram_features(p[:, :3].numpy()) for p in b["points"]]))
        inputs = (b["rgb_u8"].float(), lidar, b["lane"], b["lane_num"], b["radar"], b["radar_adj"],
                  b["target_point"], b["velocity"])
        mmfn_oracle.train_step(sd, cfg, dict(inputs=inputs, gt_waypoints=b["gt_waypoints"]), opt_state=opt)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return batch * len(times) / sum(times), sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample_b = 4
    sps, sec = cpu_port_throughput(sample_b, args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": "MMFN training samples/sec", "value": sps, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "full MMFN (RGB+LiDAR+map+radar) fwd+bwd+AdamW, fp32 (BASELINE configs[1])",
                   "per_gpu_batch": PER_GPU_BATCH, "sample": f"batch {sample_b} per step on host CPU"},
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps of batch {sample_b} (oracle/mmfn_oracle.train_step incl. numpy BEV histogram)"},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ncu --set full captures of the headline shapes (tools/ncu_targets.py -> profiles/r01_ncu_full_kernels.json)
NCU_KERNEL_OF = {
    ("conv2d_fwd_tf32", 16, 16, 256, 256, 3, 16): ("conv3x3_patch_kernel<64, 0, 1>", "(4, 32, 1)"),
    ("conv2d_fwd_tf32", 16, 64, 64, 64, 3, 64): ("conv3x3_patch_kernel<64, 0, 0>", "(1, 512, 1)"),
    ("gemm_tf32", 4096, 2048, 512, 1): ("tc_kernel<GemmOp<0, 0, 128>, 128, 3, 1>", "(16, 32, 1)"),
}


def ncu_traffic(dom):
    """dram__bytes_read + dram__bytes_write per launch of the dominant shape from the committed ncu capture."""
    path = os.path.join(ROOT, "profiles", "r01_ncu_full_kernels.json")
    want = NCU_KERNEL_OF.get(tuple(dom["key"]))
    if not want or not os.path.exists(path):
        return None
    for rec in json.load(open(path)):
        if rec["kernel"] == want[0] and rec.get("grid") == want[1]:
            return rec.get("traffic_bytes")
    return None


def dominant_kernel_chain(ops, L, dev, prof_steps):
    """Pick the tensor-core launch shape with the largest eager total and time it as a graph chain."""
    import torch
    # the stride-1 3x3 data gradient runs the SAME kernel as the forward convolution (mirrored taps): count them together
    merged = {}
    for key, n, ms, flops in L.last_shapes:
        if key[0] == "conv2d_dgrad_tf32" and key[3] == key[4] and key[2] == key[6]:
            key = ("conv2d_fwd_tf32",) + tuple(key[1:])
        if key[0] not in ("conv2d_fwd_tf32", "gemm_tf32") or (key[0] == "gemm_tf32" and key[4] != 1):
            continue
        m = merged.setdefault(tuple(key), [0, 0.0, flops])
        m[0] += n
        m[1] += ms
    key, (n, _, flops) = max(merged.items(), key=lambda kv: kv[1][1])
    if key[0] == "conv2d_fwd_tf32":
        _, N, H, C, Co, R, Ho = key
        stride = H // Ho
        x = torch.randn(N, H, H, C, device=dev)
        w = torch.randn(Co, R, R, C, device=dev) * 0.05
        fn = lambda: ops.conv2d_fwd(x, w, stride, R // 2)
        name = f"conv2d_fwd_tf32 N{N} {H}x{H} C{C}->{Co} {R}x{R} s{stride} (implicit GEMM {N * Ho * Ho}x{Co}x{R * R * C})"
        nbytes = 4.0 * (N * H * H * C + Co * R * R * C + N * Ho * Ho * Co)
    else:
        _, M, Nn, K, _ = key
        A, Bm, Cm = torch.randn(M, K, device=dev), torch.randn(Nn, K, device=dev) * 0.02, torch.empty(M, Nn, device=dev)
        fn = lambda: ops.gemm(A, Bm, Cm)
        name = f"gemm_tf32 {M}x{Nn}x{K}"
        nbytes = 4.0 * (M * K + Nn * K + M * Nn)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    chain = 40
    with torch.cuda.graph(g):
        for _ in range(chain):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    L.next_work = None
    return dict(key=list(key), name=name, n=n // prof_steps, flops=flops, bytes=nbytes,
                us=1e3 * e0.elapsed_time(e1) / (10 * chain))


# ------------------------------------------------------------------------------------- native arm
def run_native(args):
    import torch
    import torch.distributed as dist
    from mmfn_b200 import ops, synthetic
    from mmfn_b200._lib import lib
    from mmfn_b200.config import GlobalConfig
    from mmfn_b200.engine import BatchStager, TrainEngine
    from mmfn_b200.model_rad import MMFN

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "the native arm needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    cfg = GlobalConfig()                                   # reference config: dropout 0.1 everywhere
    model = MMFN(cfg, dev)
    eng = TrainEngine(model, lr=1e-4)
    eng.broadcast_parameters()

    # distinct synthetic batches per rank (disjoint shards, like DistributedSampler)
    nbuf = 2
    host_batches = [synthetic.synth_batch(B, first_index=(rank * nbuf + i) * B) for i in range(nbuf)]
    stager = BatchStager(host_batches[0], dev)
    packed = []                                            # the same batches, packed, resident in HBM
    for hb in host_batches:
        stager.stage(hb)
        torch.cuda.synchronize()
        packed.append(stager.dev.clone())
    use_graph = not args.no_graph
    if use_graph:
        eng.capture(stager.dev_views)                      # fwd+bwd schedule -> one CUDA graph

    def run_step():
        return eng.step_graph() if use_graph else eng.step(stager.dev_views)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
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

    def step_resident(i):
        stager.dev.copy_(packed[i % nbuf])                 # device-to-device: inputs already in HBM
        run_step()

    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    # e2e feed: the loader side (collate + pin, DataLoader workers in the reference) leaves packed batches in
    # pinned host memory; every step moves its 12.4 MB batch pinned-host -> device with ONE copy (issued on a copy
    # stream while the previous step computes), runs the step and reads the loss back.
    for slot, hb in enumerate(host_batches):
        stager.pack(hb, slot)

    def step_e2e(i):
        stager.commit()                                     # batch i (H2D issued during step i-1) -> graph input buffer
        loss = run_step()
        stager.prefetch((i + 1) % nbuf)                     # H2D of batch i+1 overlaps this step's kernels
        loss_host.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()           # the caller reads the loss every step (phase2:109)

    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib().launches
    ms = timed(step_resident, args.steps)
    launches = lib().launches - l0
    sampler.stop_flag = True
    stager.prefetch(0)
    for i in range(2):
        step_e2e(i)
    ms_e2e = timed(lambda i: step_e2e(i + 2), args.steps)

    # roofline leg: per-call CUDA-event timing of every C-ABI launch over `prof_steps` live steps
    prof = None
    if rank == 0:
        prof_steps = min(args.steps, 3)
        torch.cuda.synchronize()
        lib().start_profile()
        for i in range(prof_steps):
            stager.dev.copy_(packed[i % nbuf])
            eng.forward_backward(stager.dev_views)         # eager schedule: one event pair per C-ABI call
            eng.optimizer_step(collective=False)           # rank 0 only: must not enter a collective here
        torch.cuda.synchronize()
        prof = lib().stop_profile()
        top_shapes = [dict(call=list(k), n=n // prof_steps, ms_per_step=round(ms / prof_steps, 3),
                           tflops=round(fl * n / (ms * 1e-3) / 1e12, 1) if ms > 0 else None)
                      for k, n, ms, fl in lib().last_shapes[:14]]
        for d in prof.values():
            d["ms_per_step"] = d["ms"] / prof_steps
    barrier()

    dom = None
    if rank == 0:
        dom = dominant_kernel_chain(ops, lib(), dev, prof_steps)
    barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    total_ms = sum(d["ms"] for d in prof.values())
    step_ms = ms / args.steps
    # The roofline entry describes ONE concrete launch shape: the tensor-core shape with the largest eager total.
    # Its duration is re-measured live as the average over a CUDA-graph chain of back-to-back launches (per-call
    # events in the eager leg above include the host launch gap, so they only rank shapes).
    peak = peaks["tf_sust"] / 2.0                        # fp32 storage, TF32 multiply: half the measured bf16 peak
    achieved = dom["flops"] / (dom["us"] * 1e-6) / 1e12
    roofline = {"bound": "tensor", "kernel": dom["name"], "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": ncu_traffic(dom), "algorithmic_bytes": dom["bytes"],
                "algorithmic_flops": dom["flops"],
                "peak_source": peaks["src"] + " (sustained bf16 / 2 for TF32-class fp32 math)",
                "avg_launch_ms": dom["us"] * 1e-3, "launches_per_step": dom["n"],
                "share_of_step": dom["n"] * dom["us"] * 1e-3 / step_ms,
                "timing": "CUDA events around a CUDA-graph chain of 40 back-to-back launches of this shape, 10 replays",
                "note": "fp32 operands make every large TF32 GEMM/conv L2->SM-bandwidth-bound (32 KB per 128x128x32 k-block); see DESIGN.md section 7",
                "top_shapes": top_shapes,
                "eager_ms_per_step_by_entry_point": {k: round(v["ms_per_step"], 3) for k, v in
                                                     sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:8]},
                "eager_total_ms_per_step": total_ms / prof_steps}
    sps = world * B * args.steps / (ms * 1e-3)
    sps_e2e = world * B * args.steps / (ms_e2e * 1e-3)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        c_sps, c_sec = cpu_port_throughput(4, 3, 1, threads)
        cpu = {"value": c_sps, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": "3 steps of batch 4 after 1 warm-up (oracle/mmfn_oracle.train_step incl. numpy BEV histogram)"}
    line = {
        "metric": "MMFN training samples/sec", "value": sps, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "full MMFN (RGB+LiDAR+map+radar) fwd+bwd+AdamW, fp32 (BASELINE configs[1])",
                   "per_gpu_batch": B, "global_batch": B * world, "frame": "256x256 crop of 400x300 RGB + 32768-pt LiDAR + 128 lanes x 10 nodes + 81 radar pts",
                   "dropout": 0.1, "parallelism": f"dp{world}", "cuda_graph": use_graph, "l2": "per-step working set (activations + 420 MB params) >> 126 MB L2; inputs rotate between 2 batches"},
        "model_tflops": sps * FLOP_PER_SAMPLE_FWD_BWD / 1e12,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": {"value": sps_e2e, "unit": "samples/s", "h2d_bytes_per_step": stager.nbytes, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "clocks": sampler.summary(),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:      # convenience: self-launch one rank per GPU
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                                   "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:])
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
