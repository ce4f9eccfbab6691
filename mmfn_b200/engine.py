"""Training-step engine: the B200-native equivalent of Engine.train's loop body
(run_steps/phase2_train_net.py:60-110) plus the optimizer/DDP wiring of main() (:256-269).

One step = [one packed H2D copy] -> BEV scatter -> forward -> L1 loss -> backward ->
[ONE all-reduce over the flat gradient buffer] -> ONE fused AdamW launch over the flat
parameter buffer.  The forward/backward schedule (~1900 kernel launches) is captured once into
a CUDA graph and replayed, so the step is not bound by host launch latency; dropout masks still
change every step because every dropout site adds a device-resident offset to its seed.
Data parallelism is one process per GPU (torch.distributed, NCCL); every rank keeps a full
replica, BatchNorm statistics stay rank-local as in the reference (no SyncBN).
"""
import torch
import torch.distributed as dist

from . import ops, parallel
from ._lib import lib

_FIELDS = (
    ("rgb_u8", torch.uint8), ("points", torch.float32), ("lane", torch.float32), ("lane_num", torch.int32),
    ("radar", torch.float32), ("radar_adj", torch.float32), ("velocity", torch.float32),
    ("target_point", torch.float32), ("gt_waypoints", torch.float32), ("lidar", torch.float32), ("map_u8", torch.uint8),
    # packed loader format (data.PackedShard): histogram as uint8 counts, radar azimuths in float64 instead of the
    # 81 x 81 adjacency -- both expanded on the GPU (ops.bev_unpack_u8 / ops.radar_adjacency)
    ("lidar_u8", torch.uint8), ("radar_az64", torch.float64),
)


class BatchStager:
    """Packs a host batch (dict of CPU tensors, see synthetic.synth_batch) into ONE pinned buffer and
    moves it with ONE async H2D copy into a FIXED device buffer (graph-replay friendly); the reference
    issues ~25 separate .to(device) copies per step (phase2_train_net.py:78-97)."""

    def __init__(self, example, device):
        self.device = torch.device(device)
        self.layout, off = [], 0
        for name, dtype in _FIELDS:
            if name not in example:                         # either raw `points` or a pre-built `lidar` histogram
                continue
            t = example[name]
            nbytes = t.numel() * dtype.itemsize
            self.layout.append((name, dtype, tuple(t.shape), off, nbytes))
            off += (nbytes + 255) // 256 * 256
        self.nbytes = off
        self.host = [torch.empty(off, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.dev = torch.empty(off, dtype=torch.uint8, device=self.device)
        self.dev_views = self._views(self.dev)
        self.flip = 0
        self._copy_stream = self._ready = self._consumed = None

    def _views(self, buf):
        return {name: buf[off: off + nbytes].view(dtype).view(shape) for name, dtype, shape, off, nbytes in self.layout}

    def pack(self, batch, slot):
        """Collate-side work (what a DataLoader worker + pin_memory thread do in the reference): host dict ->
        pinned buffer `slot`."""
        hv = self._views(self.host[slot])
        for name, dtype, _, _, _ in self.layout:
            hv[name].copy_(batch[name].to(dtype))

    def stage(self, batch):
        """host dict -> device dict (views of the fixed device buffer); asynchronous on the current stream."""
        self.flip ^= 1
        self.pack(batch, self.flip)
        self.dev.copy_(self.host[self.flip], non_blocking=True)
        return self.dev_views

    # ---- double-buffered input feed: the H2D copy of step i+1 rides a copy stream under step i's kernels ----
    def prefetch(self, slot):
        """Start the ONE packed H2D copy of pinned buffer `slot` into the device staging buffer (copy stream)."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self.dev_next = torch.empty_like(self.dev)
        cs = self._copy_stream
        if self._consumed is not None:
            cs.wait_event(self._consumed)              # the previous commit() has finished reading dev_next
        with torch.cuda.stream(cs):
            self.dev_next.copy_(self.host[slot], non_blocking=True)
            self._ready = torch.cuda.Event()
            self._ready.record(cs)

    def commit(self):
        """Make the prefetched batch the current one: device-to-device copy into the fixed (graph-captured) buffer."""
        main = torch.cuda.current_stream()
        main.wait_event(self._ready)
        self.dev.copy_(self.dev_next, non_blocking=True)
        self._consumed = torch.cuda.Event()
        self._consumed.record(main)
        return self.dev_views


def batch_inputs(b, lazy=False):
    """(lidar histogram, radar adjacency) of a device batch in any of its forms: a pre-built `lidar` histogram, the
    packed loader's `lidar_u8` counts, or a raw `points` sweep; `radar_adj` or the packed loader's float64 azimuths.
    lazy: zero-argument callables for whatever needs a kernel, so that the network issues it inside the branch that
    consumes it (off the image trunk's stream)."""
    if "lidar" in b:
        lidar = b["lidar"]
    elif "lidar_u8" in b:
        lidar = lambda: ops.bev_unpack_u8(b["lidar_u8"])
    else:
        lidar = lambda: ops.bev_scatter(b["points"])
    radar_adj = b.get("radar_adj")
    if radar_adj is None and "radar_az64" in b:
        radar_adj = lambda: ops.radar_adjacency(b["radar_az64"])
    if not lazy:
        lidar = lidar() if callable(lidar) else lidar
        radar_adj = radar_adj() if callable(radar_adj) else radar_adj
    return lidar, radar_adj


class TrainEngine:
    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, process_group=None):
        self.model, self.net, self.st = model, model.net, model.store
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        n = self.st.n_active
        dev = self.st.device
        self.m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.v = torch.zeros(n, device=dev, dtype=torch.float32)
        self.state = torch.zeros(3, device=dev, dtype=torch.float32)
        self.p_active = self.st.flat[:n]
        self.g_active = self.st.flat_grad[:n]
        self.p16_active = self.st.flat16[:n]         # bf16 weight shadow, refreshed by the AdamW kernel (bf16 configuration)
        self.st.sync_shadow()
        self.last_pred = None
        # device-resident dropout offset, bumped once per step inside the (captured) schedule.  ONE cell per
        # device, owned by the library binding (never freed, never rebound by a second engine); the data-parallel
        # rank is folded into its initial value so replicas draw different masks.
        rank = dist.get_rank(process_group) if self.world > 1 else 0
        self.rng = lib().rng_tensor(dev, rank)
        self._graph = None

    def broadcast_parameters(self, src=0):
        """What DistributedDataParallel's constructor does (phase2_train_net.py:269)."""
        parallel.broadcast_([self.st.flat, self.st.flat_buf], src, self.pg)

    # ---- eager schedule ----------------------------------------------------------------------
    def forward_backward(self, b):
        """b: device batch. Returns the loss (0-d device tensor); gradients land in store.flat_grad."""
        model = self.model
        model.train()
        if ops.BF16 and not torch.cuda.is_current_stream_capturing():
            self.st.sync_shadow()                      # eager callers may have touched the fp32 masters; replays rely on AdamW
        # the gradient buffer is first touched by backward: its zeroing rides a side stream under the first fusion
        # transformer's forward (net.idle_hook) instead of heading the step
        self.net.idle_hook = self.st.flat_grad.zero_
        self.rng.add_(1000003)
        self.st.flat_nbt.add_(model._nbt_step())
        # input kernels of the LiDAR / radar branches run inside those branches (callables), not ahead of the image trunk
        lidar, radar_adj = batch_inputs(b, lazy=True)
        image = b["rgb_u8"] if "rgb_u8" in b else b["image"]
        # the RGB+LiDAR-only variant (transfuser.TransFuser) has no lane / radar inputs
        lane = b["map_u8"] if model.VARIANT == "img" else b.get("lane")     # model_img: rasterised map image
        pred = self.net.forward(image, lidar, lane, b.get("lane_num"), b.get("radar"), radar_adj,
                                b["target_point"], b["velocity"], model.seed, True)
        self.net.idle_hook = None
        loss, dpred = ops.l1_loss(pred, b["gt_waypoints"])
        self.net.backward(dpred)
        self.last_pred = pred
        return loss

    def optimizer_step(self, collective=True):
        if collective:
            parallel.allreduce_sum_(self.g_active, self.pg)                       # ONE collective per step
        ops.adamw_step_(self.p_active, self.g_active, self.m, self.v, self.state, self.lr, self.betas[0],
                        self.betas[1], self.eps, self.wd, grad_scale=1.0 / self.world,
                        p16=self.p16_active if ops.BF16 else None)

    def step(self, device_batch):
        loss = self.forward_backward(device_batch)
        self.optimizer_step()
        return loss

    # ---- CUDA-graph schedule -----------------------------------------------------------------
    def capture(self, static_batch, warmup=2):
        """Capture forward+backward on `static_batch` (fixed device tensors, e.g. BatchStager.dev_views) as THREE graphs
        split where a gradient bucket is final: the early bucket (head, transformer4, radar encoder, layer4 of every
        trunk: [n_mid, n_active) of the flat buffer) after the last fusion stage, the mid bucket (layer3, transformer3,
        transformer2: [n_late, n_mid)) after transformer2's backward.  Their all-reduce and AdamW updates run on a side
        stream under the following graph; only the late bucket [0, n_late) is exchanged after backward.  Later steps
        refill the static tensors in place and call step_graph()."""
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            # sets kernel attributes, warms the allocator.  forward+backward only: the optimizer kernels are not part
            # of the captured graphs, and a warm-up must not move weights, m / v or the step count.  BatchNorm running
            # statistics / num_batches_tracked are training state too: restored after the warm-up.
            buf0, nbt0 = self.st.flat_buf.clone(), self.st.flat_nbt.clone()
            for _ in range(warmup):
                self.forward_backward(static_batch)
            self.st.flat_buf.copy_(buf0)
            self.st.flat_nbt.copy_(nbt0)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        graphs = [torch.cuda.CUDAGraph() for _ in range(3)]
        pool = torch.cuda.graph_pool_handle()
        l0 = lib().launches
        cap = torch.cuda.Stream(priority=-1)     # critical chain: above the leaf (weight-gradient) streams
        cap.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cap):
            graphs[0].capture_begin(pool=pool)
            state = {"i": 0}

            def split(which):
                assert which == ("early", "mid")[state["i"]], which
                graphs[state["i"]].capture_end()
                state["i"] += 1
                graphs[state["i"]].capture_begin(pool=pool)
            self.net.mid_hook = split
            try:
                self._graph_loss = self.forward_backward(static_batch)
                self._graph_pred = self.last_pred
            finally:
                self.net.mid_hook = None
            assert state["i"] == 2, "the schedule did not report both bucket boundaries"
            graphs[2].capture_end()
        torch.cuda.current_stream().wait_stream(cap)
        torch.cuda.synchronize()
        self.graph_launches = lib().launches - l0
        self._graph = tuple(graphs)
        self._comm = torch.cuda.Stream()
        return self._graph

    def _bucket_update(self, lo, hi):
        """all-reduce (data parallel) + AdamW of the flat range [lo, hi) on the current stream"""
        if hi <= lo:
            return
        g = self.g_active[lo:hi]
        parallel.allreduce_sum_(g, self.pg)
        ops.adamw_apply_(self.p_active[lo:hi], g, self.m[lo:hi], self.v[lo:hi], self.state, self.lr, self.betas[0],
                         self.betas[1], self.eps, self.wd, grad_scale=1.0 / self.world,
                         p16=self.p16_active[lo:hi] if ops.BF16 else None)

    def step_graph(self):
        """Replay the captured step.  Order on the device: graph A (zero grads, BEV, forward, backward down to the last
        fusion stage) -> [side stream: early bucket all-reduce + AdamW] || graph B (backward through transformer2) ->
        [side stream: mid bucket] || graph C (rest of backward) -> late bucket all-reduce + AdamW.  Returns the loss."""
        ga, gb, gc = self._graph
        main = torch.cuda.current_stream()
        st = self.st
        ops.adamw_advance_(self.state, self.betas[0], self.betas[1])
        ga.replay()
        ev = torch.cuda.Event()
        ev.record(main)
        self._comm.wait_event(ev)
        with torch.cuda.stream(self._comm):
            self._bucket_update(st.n_mid, st.n_active)
        gb.replay()
        ev2 = torch.cuda.Event()
        ev2.record(main)
        self._comm.wait_event(ev2)
        with torch.cuda.stream(self._comm):
            self._bucket_update(st.n_late, st.n_mid)
            done = torch.cuda.Event()
            done.record(self._comm)
        gc.replay()
        self._bucket_update(0, st.n_late)
        main.wait_event(done)
        lib().launches += self.graph_launches
        self.last_pred = self._graph_pred
        return self._graph_loss
