"""B200-native MMFN (RGB + LiDAR-BEV + VectorNet map + radar GAT -> 4 fusion transformers -> GRU
waypoint head).  Drop-in for team_code/mmfn_utils/models/model_rad.py:MMFN (:639-739):

    MMFN(config, device)
    forward(image_list, lidar_list, maps_list, vectormaps_list, radar_list, radar_adj,
            target_point, velocity) -> (B, pred_len, 2)
    control_pid(waypoints, velocity)

with the reference's state_dict keys/shapes, selectable through the reference plugin hook
(train_agent.entry_point = "mmfn_b200.model_rad:MMFN", run_steps/utils.py:68-72).

Nothing here uses torch autograd or ATen math: forward and backward are explicit schedules of
libmmfn_b200.so kernels over NHWC fp32 activations; torch provides memory, streams and the
nn.Module/Parameter bookkeeping.  `loss.backward()` still works for callers that want it (one
autograd.Function around the whole network), but the fast path is engine.TrainEngine.
"""
import math
import os
from collections import deque

import numpy as np
import torch
from torch import nn

from . import ops
from .params import ParamStore, RESNET18, RESNET34, WIDTHS, is_unused

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


# --------------------------------------------------------------------------- building blocks
class _Aux:
    """Weight-gradient side streams.  In backward only the data-gradient chain is sequential; every
    weight/bias gradient (wgrad conv, dW GEMM, bias column sum, LayerNorm parameter sums) is a LEAF of the
    dependency graph.  Leaves are issued round-robin on a small pool of auxiliary streams paired with the current
    one, so they overlap the dgrad chain and each other (parallel graph branches once captured) -- one FIFO
    stream would chain ~65 us of leaf work per transformer layer behind a ~50 us critical path.  Work the
    critical path re-joins soon (fork(): attention dK / dV, filter flips) gets its own stream pair so it never
    queues behind leaves.  Inputs are kept alive until join_all() -- no allocator reuse hazards."""
    enabled = True
    N_LEAF, N_FORK = int(os.environ.get("MMFN_AUX_LEAF", "4")), 2
    pools, used, keep = {}, {}, []
    epoch = 0            # bumped by join_all(): fork events of earlier epochs are already ordered before the caller

    @classmethod
    def _pool(cls, prim):
        pool = cls.pools.get(prim.cuda_stream)
        if pool is None:
            # leaves at the lowest priority: their CTAs only fill SMs the data-gradient chain leaves idle
            mk = lambda n, pr: [torch.cuda.Stream(device=prim.device, priority=pr) for _ in range(n)]
            pool = cls.pools[prim.cuda_stream] = {"leaf": mk(cls.N_LEAF, 0), "fork": mk(cls.N_FORK, -1), "i": 0, "j": 0}
        return pool

    @classmethod
    def _issue(cls, aux, prim, fn, tensors):
        ev = torch.cuda.Event()
        ev.record(prim)
        aux.wait_event(ev)
        cls.keep.extend(tensors)
        with torch.cuda.stream(aux):
            fn()
        cls.used[aux.cuda_stream] = aux

    deferred = None      # list while a trunk phase defers its leaves to the following transformer phase

    @classmethod
    def defer_begin(cls):
        if cls.enabled:
            cls.deferred = []

    @classmethod
    def defer_flush(cls):
        """Issue the deferred leaves from the current stream (after the trunk phase has re-joined it): they now overlap
        the latency-bound fusion-transformer backward that follows instead of competing with the trunks' own chain."""
        jobs, cls.deferred = cls.deferred, None
        for fn, tensors in jobs or []:
            cls.run(fn, *tensors)

    @classmethod
    def run(cls, fn, *tensors):
        if not cls.enabled:
            fn()
            return
        if cls.deferred is not None:
            cls.deferred.append((fn, tensors))
            return
        prim = torch.cuda.current_stream()
        pool = cls._pool(prim)
        aux = pool["leaf"][pool["i"] % cls.N_LEAF]
        pool["i"] += 1
        cls._issue(aux, prim, fn, tensors)

    @classmethod
    def fork(cls, fn, *tensors):
        """Like run(), on the fork streams, and returns an event the caller's stream can wait on (local fork/join)."""
        if not cls.enabled:
            fn()
            return None
        prim = torch.cuda.current_stream()
        pool = cls._pool(prim)
        aux = pool["fork"][pool["j"] % cls.N_FORK]
        pool["j"] += 1
        cls._issue(aux, prim, fn, tensors)
        ev = torch.cuda.Event()
        ev.record(aux)
        return (ev, cls.epoch)

    @classmethod
    def wait(cls, *events):
        for e in events:
            if e is not None and e[1] == cls.epoch:       # older epochs were joined (possibly in an earlier graph capture)
                torch.cuda.current_stream().wait_event(e[0])

    @classmethod
    def join_all(cls):
        main = torch.cuda.current_stream()
        for aux in cls.used.values():
            ev = torch.cuda.Event()
            ev.record(aux)
            main.wait_event(ev)
        cls.used.clear()
        cls.keep.clear()
        cls.epoch += 1


class ConvBN:
    """conv (bias-free) -> BatchNorm2d [-> + residual] [-> ReLU]; torchvision BasicBlock pieces."""

    def __init__(self, st, conv_key, bn_key, stride, pad, dgrad=True):
        self.dgrad = dgrad
        self.w, self.dw = st.p(conv_key), st.g(conv_key)
        self.w16 = st.p16(conv_key)                 # bf16 shadow of the filters (bf16 configuration)
        self.gam, self.dgam = st.p(bn_key + ".weight"), st.g(bn_key + ".weight")
        self.bet, self.dbet = st.p(bn_key + ".bias"), st.g(bn_key + ".bias")
        self.rm, self.rv = st.buf(bn_key + ".running_mean"), st.buf(bn_key + ".running_var")
        self.stride, self.pad = stride, pad

    def fwd(self, x, res=None, relu=True, train=True, no_twin=False, only16=False):
        """only16 (a hint): the caller's only reader of y is a bf16 convolution -- where the BatchNorm-apply kernel runs
        separately and the backward takes its ReLU mask from z, y is written as bf16 only."""
        self.x, self.relu = x, relu
        self.col, self.direct = None, False
        self.has_res = res is not None
        # bf16 configuration: the convolution reads the bf16 twin its producer wrote (x.h) and the bf16 filter shadow;
        # its fp32 output z feeds the batch statistics; BatchNorm-apply writes y in fp32 AND its bf16 twin
        Ho, Wo = ops.conv_out_hw(x.shape[1], x.shape[2], self.w.shape[1], self.w.shape[2], self.stride, self.pad)
        self.bf = ops.twin(x) is not None and ops.bf16_conv_ok(x.shape[3], self.w.shape[0], Ho, Wo)
        xin, win = (x.h, self.w16) if self.bf else (x, self.w)
        if train and ops.conv_bn_fusable(xin, win, self.stride, self.pad):
            # batch statistics from the convolution's own epilogue (+ last-CTA finalize): no reduction pass over z
            self.z, self.mean, self.rstd = ops.conv2d_fwd_bn(xin, win, self.stride, self.pad, self.rm, self.rv)
            only16 = (only16 and ops.BF16 and ops.BF16_ONLY_INNER and ops.BN_MASK_FROM_Z and relu and res is None
                      and not no_twin)
            self.y = ops.bn_apply(self.z, self.gam, self.bet, self.mean, self.rstd, res=res, relu=relu,
                                  want16=ops.BF16 and not no_twin, only16=only16)
            return self.y
        if self.bf:
            self.z = ops.conv2d_fwd(x.h, self.w16, self.stride, self.pad)
        elif ops.stem_conv_direct_ok(x, self.w, self.stride, self.pad):
            # stems: direct tensor-core kernel on a shared-memory input patch (no column matrix)
            self.z, self.direct = ops.conv2d_stem7_fwd(x, self.w), True
        elif ops.stem_uses_im2col(x, self.w):
            # stems: im2col + one dense tensor-core GEMM; the column matrix is kept for the weight gradient
            self.z, self.col, self.w_pad = ops.conv2d_fwd_im2col(x, self.w, self.stride, self.pad, getattr(self, "w_pad", None),
                                                                 bf16=ops.BF16)
        else:
            self.z = ops.conv2d_fwd(x, self.w, self.stride, self.pad)
        bn = ops.bn_train_fwd if train else ops.bn_eval_fwd
        self.y, self.mean, self.rstd = bn(self.z, self.gam, self.bet, self.rm, self.rv, res=res, relu=relu,
                                          want16=ops.BF16 and not no_twin)
        return self.y

    def bwd(self, dy, need_dx=True, want_dres=False, dx_res=None):
        # ReLU mask: recomputed from z when the ReLU followed the BatchNorm directly, else read from y (bf16 twin if any)
        from_z = self.relu and not self.has_res and ops.BN_MASK_FROM_Z
        ymask = None if (from_z or not self.relu) else (ops.twin(self.y) if ops.twin(self.y) is not None and ops.BN_MASK_FROM_Z else self.y)
        dz, dres = ops.bn_train_bwd(dy, self.z, ymask, self.mean, self.rstd, self.gam,
                                    self.dgam, self.dbet, want_dres, relu_beta=self.bet if from_z else None,
                                    out_bf16=self.bf or (self.col is not None and self.col.dtype == torch.bfloat16)
                                    or (self.direct and ops.BF16))
        x, col = self.x, self.col
        if self.bf:
            x16 = x.h
            _Aux.run(lambda: ops.conv2d_wgrad_(dz, x16, self.dw, self.stride, self.pad), dz, x16)
        elif self.direct:
            _Aux.run(lambda: ops.conv2d_stem7_wgrad_(dz, x, self.dw), dz, x)
        elif col is not None:
            _Aux.run(lambda: ops.conv2d_wgrad_im2col_(dz, col, self.dw), dz, col)
        else:
            _Aux.run(lambda: ops.conv2d_wgrad_(dz, x, self.dw, self.stride, self.pad), dz, x)
        dx = None
        if need_dx:
            dx = ops.conv2d_dgrad(dz, self.w16 if self.bf else self.w, self.x.shape, self.stride, self.pad, res=dx_res)
        self.x = self.z = self.y = self.col = None
        return dx, dres


class BasicBlock:
    def __init__(self, st, prefix, stride, downsample):
        self.c1 = ConvBN(st, prefix + ".conv1.weight", prefix + ".bn1", stride, 1)
        self.c2 = ConvBN(st, prefix + ".conv2.weight", prefix + ".bn2", 1, 1)
        self.ds = ConvBN(st, prefix + ".downsample.0.weight", prefix + ".downsample.1", stride, 0) if downsample else None

    def fwd(self, x, train):
        # a is read by conv2 only: bf16-only when conv2 takes the bf16 tensor-core path (same spatial size, stride 1)
        Ho, Wo = ops.conv_out_hw(x.shape[1], x.shape[2], 3, 3, self.c1.stride, 1)
        a = self.c1.fwd(x, relu=True, train=train,
                        only16=train and ops.BF16 and ops.bf16_conv_ok(self.c1.w.shape[0], self.c2.w.shape[0], Ho, Wo))
        idn = self.ds.fwd(x, relu=False, train=train, no_twin=True) if self.ds else x     # residual only: fp32 suffices
        return self.c2.fwd(a, res=idn, relu=True, train=train)

    def bwd(self, dout):
        da, didn = self.c2.bwd(dout, want_dres=True)
        if self.ds:
            didn, _ = self.ds.bwd(didn)
        dx, _ = self.c1.bwd(da, dx_res=didn)
        return dx


class ResLayer:
    def __init__(self, st, prefix, nblocks, stride):
        self.blocks = [BasicBlock(st, f"{prefix}.{i}", stride if i == 0 else 1, i == 0 and stride != 1)
                       for i in range(nblocks)]

    def fwd(self, x, train):
        for b in self.blocks:
            x = b.fwd(x, train)
        return x

    def bwd(self, d):
        for b in reversed(self.blocks):
            d = b.bwd(d)
        return d


class Stem:
    """conv7x7/2 -> BN -> ReLU -> MaxPool3x3/2 on the (already NHWC) network input; no dgrad.  Training runs the tail
    (BN -> ReLU -> pool) as fused kernels (ops.stem_bn_relu_maxpool_*): the pre-pool activation is never written."""

    def __init__(self, st, prefix):
        self.cb = ConvBN(st, prefix + ".conv1.weight", prefix + ".bn1", 2, 3, dgrad=False)
        self.fused = False

    def fwd(self, x, train):
        cb = self.cb
        self.fused = train and ops.FUSE_STEM_TAIL
        if self.fused:
            cb.x, cb.relu, cb.bf, cb.col, cb.y, cb.direct = x, True, False, None, None, False
            if ops.stem_conv_direct_ok(x, cb.w, cb.stride, cb.pad):
                cb.z, cb.direct = ops.conv2d_stem7_fwd(x, cb.w), True
            elif ops.stem_uses_im2col(x, cb.w):
                cb.z, cb.col, cb.w_pad = ops.conv2d_fwd_im2col(x, cb.w, cb.stride, cb.pad, getattr(cb, "w_pad", None), bf16=ops.BF16)
            else:
                cb.z = ops.conv2d_fwd(x, cb.w, cb.stride, cb.pad)
            out, self.idx, cb.mean, cb.rstd = ops.stem_bn_relu_maxpool_fwd(cb.z, cb.gam, cb.bet, cb.rm, cb.rv, want16=ops.BF16)
            return out
        y = cb.fwd(x, relu=True, train=train, no_twin=True)
        self.in_shape = y.shape
        out, self.idx = ops.maxpool_fwd(y, want16=ops.BF16)
        return out

    def bwd(self, d):
        cb = self.cb
        if self.fused:
            col = cb.col
            dz = ops.stem_bn_relu_maxpool_bwd(d, self.idx, cb.z, cb.mean, cb.rstd, cb.gam, cb.dgam, cb.dbet,
                                              out_bf16=(col is not None and col.dtype == torch.bfloat16) or (cb.direct and ops.BF16))
            x = cb.x
            if cb.direct:
                _Aux.run(lambda: ops.conv2d_stem7_wgrad_(dz, x, cb.dw), dz, x)
            elif col is not None:
                _Aux.run(lambda: ops.conv2d_wgrad_im2col_(dz, col, cb.dw), dz, col)
            else:
                _Aux.run(lambda: ops.conv2d_wgrad_(dz, x, cb.dw, cb.stride, cb.pad), dz, x)
            cb.x = cb.z = cb.col = self.idx = None
            return
        dy = ops.maxpool_bwd(d, self.idx, self.in_shape)
        cb.bwd(dy, need_dx=False)
        self.idx = None


class Linear:
    def __init__(self, st, prefix, bias=True, keys=None):
        if keys is None:
            self.w, self.dw = st.p(prefix + ".weight"), st.g(prefix + ".weight")
            self.b, self.db = (st.p(prefix + ".bias"), st.g(prefix + ".bias")) if bias else (None, None)
            self.w16 = st.p16(prefix + ".weight")
        else:       # several adjacent reference Linears fused into one GEMM
            self.w, self.dw = st.fused([k + ".weight" for k in keys]), st.fused([k + ".weight" for k in keys], True)
            self.b, self.db = st.fused([k + ".bias" for k in keys]), st.fused([k + ".bias" for k in keys], True)
            self.w16 = st.fused16([k + ".weight" for k in keys])

    def _w(self, like):
        """the weight copy matching an operand's element type: bf16 shadow for bf16 activations, fp32 master otherwise"""
        return self.w16 if like.dtype == torch.bfloat16 else self.w

    def fwd(self, x, out=None, act=0, res=None, drop_p=0.0, seed=0, out_bf16=False):
        """x (M,K) -> (M,N) = drop(act(x W^T + b)) + res.  A bf16 x multiplies the bf16 weight shadow; out_bf16 writes
        the result as bf16 (tensors that only feed further GEMMs)."""
        self.x, self.act = x, act
        y = out if out is not None else torch.empty((x.shape[0], self.w.shape[0]), device=x.device,
                                                    dtype=torch.bfloat16 if out_bf16 else torch.float32)
        ops.gemm(x, self._w(x), y, bias=self.b, res=res, act=act, drop_p=drop_p, seed=seed)
        self.y = y if act else None
        return y

    def bwd(self, dy, need_dx=True, dx_mask=None, masked=False, dx_bf16=False):
        """dy: gradient of this layer's output (dropout mask already applied by the caller).  The
        ReLU derivative is applied here from the saved output unless `masked`.  dx_mask fuses the
        ReLU mask of the PRODUCER of x into the dgrad GEMM epilogue.  bf16 dy (bf16 configuration): the weight
        gradient multiplies dy^T by the saved bf16 x, the data gradient dy by the bf16 weight shadow."""
        if self.act == 1 and not masked:
            dy = ops.relu_bwd(dy, self.y)
        x = self.x

        def wgrad():
            if self.db is not None:
                ops.colsum_(dy, self.db)
            if (tuple(self.dw.shape) == (64, 7) and dy.dtype == x.dtype == torch.float32 and dy.shape[0] >= 4096
                    and dy.is_contiguous() and x.is_contiguous()):
                ops.wgrad_n64_k7_(dy, x, self.dw)          # polyline input layer: K = 7 is no shape for a GEMM tile
            else:
                ops.gemm(dy.t(), x.t(), self.dw, accum=1)
        _Aux.run(wgrad, dy, x)
        dx = None
        if need_dx:
            dx = torch.empty((dy.shape[0], self.w.shape[1]), device=dy.device, dtype=torch.bfloat16 if dx_bf16 else torch.float32)
            ops.gemm(dy, self._w(dy).t(), dx, mask=dx_mask)
        self.x = self.y = None
        return dx


class LayerNorm:
    def __init__(self, st, prefix, act=0):
        self.g, self.dg = st.p(prefix + ".weight"), st.g(prefix + ".weight")
        self.b, self.db = st.p(prefix + ".bias"), st.g(prefix + ".bias")
        self.act = act

    def fwd(self, x, out=None, out_bf16=False):
        self.x = x
        y, self.mean, self.rstd = ops.layernorm_fwd(x, self.g, self.b, act=self.act, out=out, out_bf16=out_bf16)
        return y

    def bwd(self, dy, dres=None, drop=None, drop_bf16=False):
        """dx (and, with drop=(p, seed), also dx * dropout mask; drop_bf16: that copy in bf16).  The parameter-gradient
        reduction is a leaf of the backward graph: it runs on the auxiliary stream."""
        x, mean, rstd = self.x, self.mean, self.rstd
        _Aux.run(lambda: ops.layernorm_bwd(dy, x, self.g, self.b, mean, rstd, self.dg, self.db, act=self.act, parts=2),
                 dy, x, mean, rstd)
        out = ops.layernorm_bwd(dy, x, self.g, self.b, mean, rstd, None, None, act=self.act, dres=dres, parts=1, drop=drop,
                                drop_bf16=drop_bf16)
        self.x = None
        return out


class Block:
    """Pre-LN transformer block (model_rad.py:112-133) with fused QKV GEMM."""

    def __init__(self, st, prefix, C, n_head, attn_p, resid_p):
        self.C, self.nh, self.hs = C, n_head, C // n_head
        self.attn_p, self.resid_p = attn_p, resid_p
        self.ln1, self.ln2 = LayerNorm(st, prefix + ".ln1"), LayerNorm(st, prefix + ".ln2")
        self.qkv = Linear(st, None, keys=[f"{prefix}.attn.{n}" for n in ("key", "query", "value")])
        self.proj = Linear(st, prefix + ".attn.proj")
        self.fc1, self.fc2 = Linear(st, prefix + ".mlp.0"), Linear(st, prefix + ".mlp.2")

    def _heads(self, t2d, B, T, col0):
        """(B*T, 3C) column slice -> (B, nh, T, hs) view"""
        return t2d[:, col0: col0 + self.C].view(B, T, self.nh, self.hs).permute(0, 2, 1, 3)

    def fwd(self, x, B, T, seed, train):
        C, nh, hs = self.C, self.nh, self.hs
        ap, rp = (self.attn_p, self.resid_p) if train else (0.0, 0.0)
        self.B, self.T, self.seed, self.ap, self.rp = B, T, seed, ap, rp
        # bf16 configuration: tensors that only feed the linears (LayerNorm outputs, attention output, MLP hidden) are
        # bf16 and multiply the bf16 weight shadow; the residual stream x, the qkv buffer and the attention core
        # (TF32 tcgen05 kernel, fp32 softmax) stay fp32
        bf = self.bf = ops.BF16 and C % 64 == 0
        self.bfa = bf and ops.attention_bf16_ok(T, C, nh) and ops.BF16_ATTN    # attention core in bf16 too (attn_bf16.cu)
        h1 = self.ln1.fwd(x, out_bf16=bf)
        qkv = self.qkv.fwd(h1, out_bf16=self.bfa)                    # columns [key | query | value]
        if self.bfa:
            # bf16 operands, K / V resident in shared memory, single-exp softmax; P / Pd saved as bf16
            y, self.P, self.Pd, _ = ops.attention_fwd_bf16(qkv, B, T, C, nh, ap, seed)
        elif ops.attention_fwd_ok(T, C, nh):
            # S = QK^T -> softmax -> dropout -> PV in ONE tcgen05 kernel; only P (and Pd) reach HBM
            y, self.P, self.Pd = ops.attention_fwd(qkv, B, T, C, nh, ap, seed, y_bf16=bf)
        else:
            k, q, v = (self._heads(qkv, B, T, i * C) for i in range(3))
            S = torch.empty((B, nh, T, T), device=x.device, dtype=torch.float32)
            ops.gemm(q, k, S)
            self.P, self.Pd = ops.softmax_fwd(S, 1.0 / math.sqrt(hs), ap, seed)
            y = torch.empty((B * T, C), device=x.device, dtype=torch.float32)
            ops.gemm(self.Pd, v.transpose(-1, -2), y.view(B, T, nh, hs).permute(0, 2, 1, 3))
            if bf:
                y = ops.to_bf16(y)
        self.qkv_out = qkv
        x1 = self.proj.fwd(y, res=x, drop_p=rp, seed=seed + 1)
        h2 = self.ln2.fwd(x1, out_bf16=bf)
        a = self.fc1.fwd(h2, act=1, out_bf16=bf)
        return self.fc2.fwd(a, res=x1, drop_p=rp, seed=seed + 2)

    def params12(self, bf):
        """operand-typed weights, fp32 biases and LayerNorm parameters in the order mmfn_gpt_small_fwd's table wants"""
        w = (lambda lin: lin.w16 if bf else lin.w)
        return (w(self.qkv), w(self.proj), w(self.fc1), w(self.fc2), self.qkv.b, self.proj.b, self.fc1.b, self.fc2.b,
                self.ln1.g, self.ln1.b, self.ln2.g, self.ln2.b)

    def adopt(self, o, l, x_in, B, T, seed, ap, rp, bf):
        """take over block l's saved tensors from the whole-GPT forward kernel: exactly the state fwd() leaves behind"""
        self.B, self.T, self.seed, self.ap, self.rp = B, T, seed, ap, rp
        self.bf = self.bfa = bf
        self.ln1.x, self.ln1.mean, self.ln1.rstd = x_in, o["mean1"][l], o["rstd1"][l]
        self.qkv.x, self.qkv.act, self.qkv.y = o["h1"][l], 0, None
        self.qkv_out, self.P, self.Pd = o["qkv"][l], o["P"][l], o["Pd"][l]
        self.proj.x, self.proj.act, self.proj.y = o["y"][l], 0, None
        self.ln2.x, self.ln2.mean, self.ln2.rstd = o["x1"][l], o["mean2"][l], o["rstd2"][l]
        self.fc1.x, self.fc1.act, self.fc1.y = o["h2"][l], 1, o["a"][l]
        self.fc2.x, self.fc2.act, self.fc2.y = o["a"][l], 0, None

    def bwd(self, dx2, dz, drop_prev=None):
        """dx2: gradient of the block output; dz = dx2 * dropout mask of the fc2 branch (produced by the
        LayerNorm backward that made dx2; bf16 in the bf16 configuration).  drop_prev=(p, seed) asks for the same pair
        for the block below."""
        bf = self.bf
        a = self.fc1.y
        da = self.fc2.bwd(dz, dx_mask=a, dx_bf16=bf)                  # ReLU mask fused into the dgrad GEMM
        dh2 = self.fc1.bwd(da, masked=True)
        dx1, dzp = self.ln2.bwd(dh2, dres=dx2, drop=(self.rp, self.seed + 1), drop_bf16=bf)
        dy = self.proj.bwd(dzp, dx_bf16=self.bfa)                  # operand of the attention-gradient products (bf16 / TF32)
        dqkv = self.attn_bwd(dy)
        dh1 = self.qkv.bwd(dqkv)
        self.P = self.Pd = self.qkv_out = None
        return self.ln1.bwd(dh1, dres=dx1, drop=drop_prev if drop_prev is not None else (0.0, 0),
                            drop_bf16=bf and drop_prev is not None)

    def attn_bwd(self, dy):
        """dy: gradient of the attention output (B*T, C) -> dqkv (B*T, 3C) [key | query | value]"""
        B, T, C, nh, hs = self.B, self.T, self.C, self.nh, self.hs
        bf = self.bf
        y_att = self.proj.x                                        # attention output saved by the projection
        qkv = self.qkv_out
        want = torch.bfloat16 if bf else torch.float32              # operand type of the qkv linear's backward
        if dy.dtype == self.P.dtype == qkv.dtype == want and ops.attention_bwd_small_ok(T, C, nh, self.P):
            # narrow heads: dPd, softmax backward, dQ, dK and dV in ONE launch per block (csrc/attn_bwd_small.cu)
            return ops.attention_bwd_small(qkv, dy, self.P, self.Pd, B, T, C, nh)
        k, q, v = (self._heads(qkv, B, T, i * C) for i in range(3))
        dqkv = torch.empty(qkv.shape, device=qkv.device, dtype=torch.bfloat16 if bf else torch.float32)
        dk, dq, dv = (self._heads(dqkv, B, T, i * C) for i in range(3))
        dyh = dy.view(B, T, nh, hs).permute(0, 2, 1, 3)
        Pd = self.Pd
        # critical chain: dPd -> dS -> dQ; dV and dK are leaves until the qkv backward: fork streams
        ev_v = _Aux.fork(lambda: ops.gemm(Pd.transpose(-1, -2), dyh.transpose(-1, -2), dv), Pd, dy, dqkv)
        if ops.attention_fwd_ok(T, C, nh) and ops.FUSED_ATTN_BWD and not bf:
            # ONE tcgen05 kernel: dPd stays in TMEM, dS tiles feed the dQ MMA and are stored for the dK GEMM
            dS = ops.attention_bwd_dq(qkv, dy, y_att, self.P, dqkv, B, T, C, nh, self.ap, self.seed)
            ev_k = _Aux.fork(lambda: ops.gemm(dS.transpose(-1, -2), q.transpose(-1, -2), dk), dS, qkv, dqkv)
        else:
            dPd = torch.empty((B, nh, T, T), device=dy.device, dtype=torch.float32)
            ops.gemm(dyh, v, dPd)
            dS = ops.softmax_bwd(self.P, dPd, 1.0 / math.sqrt(hs), self.ap, self.seed)
            ev_k = _Aux.fork(lambda: ops.gemm(dS.transpose(-1, -2), q.transpose(-1, -2), dk), dS, qkv, dqkv)
            ops.gemm(dS, k.transpose(-1, -2), dq)
        _Aux.wait(ev_v, ev_k)
        return dqkv

    def leaves_rows_b(self, o):
        """parameter gradients of the MLP, LayerNorm2 and the projection from the row kernel's products (side streams)"""
        self.fc2.bwd(o["dz"], need_dx=False)
        self.fc1.bwd(o["da"], need_dx=False, masked=True)
        ln = self.ln2
        _Aux.run(lambda dy=o["dh2"], x=ln.x, m=ln.mean, r=ln.rstd: ops.layernorm_bwd(dy, x, ln.g, ln.b, m, r, ln.dg, ln.db, act=0, parts=2),
                 o["dh2"], ln.x, ln.mean, ln.rstd)
        ln.x = None
        self.proj.bwd(o["dzp"], need_dx=False)

    def leaves_rows_a(self, dqkv, dh1):
        """parameter gradients of the qkv linear and LayerNorm1"""
        self.qkv.bwd(dqkv, need_dx=False)
        ln = self.ln1
        _Aux.run(lambda dy=dh1, x=ln.x, m=ln.mean, r=ln.rstd: ops.layernorm_bwd(dy, x, ln.g, ln.b, m, r, ln.dg, ln.db, act=0, parts=2),
                 dh1, ln.x, ln.mean, ln.rstd)
        ln.x = None
        self.P = self.Pd = self.qkv_out = None


class FusionGPT:
    """GPT / RadarGPT (model_rad.py:136-247, :887-1000): pool to 8x8 anchors, add pos/vel
    embeddings, n_layer blocks, ln_f.  Tokens are modality-major, (row, col) raster order."""

    def __init__(self, st, prefix, C, nmod, cfg, site):
        self.C, self.nmod, self.T = C, nmod, nmod * 64
        self.pos, self.dpos = st.p(prefix + ".pos_emb")[0], st.g(prefix + ".pos_emb")[0]
        self.vw, self.dvw = st.p(prefix + ".vel_emb.weight")[:, 0], st.g(prefix + ".vel_emb.weight")[:, 0]
        self.vb, self.dvb = st.p(prefix + ".vel_emb.bias"), st.g(prefix + ".vel_emb.bias")
        self.blocks = [Block(st, f"{prefix}.blocks.{i}", C, cfg.n_head, cfg.attn_pdrop, cfg.resid_pdrop)
                       for i in range(cfg.n_layer)]
        self.ln_f = LayerNorm(st, prefix + ".ln_f")
        self.embd_p, self.site = cfg.embd_pdrop, site
        self.fused, self.wT, self.tab_dev = False, None, None

    def fwd(self, feats, velocity, seed, train):
        B = feats[0].shape[0]
        self.shape, self.vel = feats[0].shape, velocity
        self.ep = self.embd_p if train else 0.0
        self.seed = seed + self.site * 100
        x = ops.tokens_fwd(feats, self.pos, self.vw, self.vb, velocity, self.ep, self.seed).view(B * self.T, self.C)
        blocks = self.blocks
        self.fused = ops.gpt_small_ok(self.C, self.T, blocks[0].nh, len(blocks), B)
        if self.fused:
            # all blocks in ONE launch (csrc/gpt_small.cu); each block adopts the tensors its backward reads
            bf = ops.BF16
            ap, rp = (blocks[0].attn_p, blocks[0].resid_p) if train else (0.0, 0.0)
            o = ops.gpt_small_fwd(x, B, self.T, self.C, blocks[0].nh, [blk.params12(bf) for blk in blocks], ap, rp, self.seed)
            for i, blk in enumerate(blocks):
                blk.adopt(o, i, x if i == 0 else o["xout"][i - 1], B, self.T, self.seed + 3 * i + 1, ap, rp, bf)
            x = o["xout"][len(blocks) - 1]
        else:
            for i, blk in enumerate(blocks):
                x = blk.fwd(x, B, self.T, self.seed + 3 * i + 1, train)
        return self.ln_f.fwd(x).view(B, self.T, self.C)

    def bwd(self, dtok_out, dfeats):
        """dtok_out (B,T,C); dfeats: per-modality feature gradients, accumulated in place."""
        blocks = self.blocks
        if self.fused and ops.FUSE_GPT_BWD:
            return self._bwd_fused_rows(dtok_out, dfeats)
        d, dz = self.ln_f.bwd(dtok_out.view(-1, self.C), drop=(blocks[-1].rp, blocks[-1].seed + 2), drop_bf16=blocks[-1].bf)
        for i in range(len(blocks) - 1, -1, -1):
            below = (blocks[i - 1].rp, blocks[i - 1].seed + 2) if i > 0 else None
            d, dz = blocks[i].bwd(d, dz, below)
        ops.tokens_bwd_(d, dfeats, self.shape, self.vel, self.dpos, self.dvw, self.dvb, self.ep, self.seed)


    def _bwd_fused_rows(self, dtok_out, dfeats):
        """Backward of a GPT whose forward ran in the whole-GPT kernel: between two attention backwards everything is
        row-local and runs in ONE launch (mmfn_gpt_small_bwd_rows: finish block i + 1, start block i); parameter
        gradients are leaves on the side streams, fed with the row kernel's products."""
        blocks, C = self.blocks, self.C
        n, M, bf = len(blocks), dtok_out.shape[0] * self.T, blocks[0].bf
        dt = torch.bfloat16 if bf else torch.float32
        if self.wT is None or self.wT.dtype != dt:
            self.wT = torch.empty((n, 12 * C * C), device=dtok_out.device, dtype=dt)
            self.tab_dev = torch.tensor([[t.data_ptr() for t in blk.params12(bf)] for blk in blocks], dtype=torch.int64, device=dtok_out.device)
        ev = _Aux.fork(lambda: ops.gpt_small_transpose(self.tab_dev, n, C, bf, self.wT), self.wT)
        d = self.ln_f.bwd(dtok_out.view(-1, C))
        _Aux.wait(ev)
        part_a = None
        for i in range(n - 1, -1, -1):
            blk = blocks[i]
            part_b = dict(dx2=d if part_a is None else None, a=blk.fc1.y, x1=blk.ln2.x, mean=blk.ln2.mean, rstd=blk.ln2.rstd,
                          gamma=blk.ln2.g, wT=self.wT[i], p=blk.rp, seed_mlp=blk.seed + 2, seed_proj=blk.seed + 1)
            o = ops.gpt_small_bwd_rows(M, C, bf, part_a, part_b)
            if part_a is not None:
                blocks[i + 1].leaves_rows_a(part_a["dqkv"], o["dh1"])
            blk.leaves_rows_b(o)
            dqkv = blk.attn_bwd(o["dy"])
            part_a = dict(dqkv=dqkv, dx1=o["dx1"], x=blk.ln1.x, mean=blk.ln1.mean, rstd=blk.ln1.rstd, gamma=blk.ln1.g, wT=self.wT[i])
        o = ops.gpt_small_bwd_rows(M, C, bf, part_a, None)
        blocks[0].leaves_rows_a(part_a["dqkv"], o["dh1"])
        ops.tokens_bwd_(o["dx"], dfeats, self.shape, self.vel, self.dpos, self.dvw, self.dvb, self.ep, self.seed)


class VectorNet:
    """VectornetEncoder (model_rad.py:327-417).  Lane-to-lane attention, agent fusion and the
    generator are evaluated for lane token 0 only -- the only one consumed (:413)."""

    def __init__(self, st, prefix):
        sg = prefix + ".lane_subgraph.layers.mlp_"
        self.sub = [(Linear(st, f"{sg}{i}.mlp.0"), LayerNorm(st, f"{sg}{i}.mlp.1", act=1)) for i in range(3)]
        self.qkv = Linear(st, prefix + ".L2L.to_qkv", bias=False)
        self.to_out = Linear(st, prefix + ".L2L.to_out.0")
        self.pos0, self.pos_ln, self.pos3 = Linear(st, prefix + ".pos_emb.0"), LayerNorm(st, prefix + ".pos_emb.1", act=2), Linear(st, prefix + ".pos_emb.3")
        self.af0, self.af_ln, self.af3 = Linear(st, prefix + ".agent_fusion.0"), LayerNorm(st, prefix + ".agent_fusion.1", act=2), Linear(st, prefix + ".agent_fusion.3")
        self.g0, self.g_ln, self.g3 = Linear(st, prefix + ".generator.0"), LayerNorm(st, prefix + ".generator.1", act=2), Linear(st, prefix + ".generator.3")

    def fwd(self, lane, lane_num):
        B, L, P, _ = lane.shape
        G, V = B * L, P - 1
        self.G, self.V, self.B, self.L, self.lane_num = G, V, B, L, lane_num
        if ops.FUSE_SUBGRAPH and V in ops.SUBGRAPH_FUSED_V:
            # one launch for the whole polyline sub-graph; it leaves behind exactly what the per-layer backward reads
            o = ops.subgraph_fused_fwd(lane, [(lin.w, lin.b, ln.g, ln.b) for lin, ln in self.sub])
            for i, (lin, ln) in enumerate(self.sub):
                lin.x, lin.act, lin.y = (o["vec"], o["x1"], o["x2"])[i], 0, None
                ln.x, ln.mean, ln.rstd = o["y"][i], o["mean"][i], o["rstd"][i]
            self.args, self.argf, tok = o["arg"], o["argf"], o["tok"]
        else:
            x = ops.lane_to_vector(lane)
            self.args = []
            for lin, ln in self.sub:
                x, arg = ops.subgraph_pool_fwd(ln.fwd(lin.fwd(x)), G, V)
                self.args.append(arg)
            tok, self.argf = ops.segmax_fwd(x, G, V)
        self.qkv_out = self.qkv.fwd(tok).view(B, L, 384)
        self.prob, att0 = ops.l2l_row0_fwd(self.qkv_out, lane_num, 2, None)
        cat = torch.empty((B, 192), device=lane.device, dtype=torch.float32)
        self.to_out.fwd(att0, out=cat[:, :128])
        zeros = torch.zeros((B, 2), device=lane.device, dtype=torch.float32)
        self.pos3.fwd(self.pos_ln.fwd(self.pos0.fwd(zeros)), out=cat[:, 128:])
        f = self.af3.fwd(self.af_ln.fwd(self.af0.fwd(cat)))
        g = self.g3.fwd(self.g_ln.fwd(self.g0.fwd(f)))                 # (B, 64*64*64) channel-major
        return ops.transpose(g.view(B, 64, 4096)).view(B, 64, 64, 64)   # NHWC

    def bwd(self, dmap):
        B, L, G, V = self.B, self.L, self.G, self.V
        dg = ops.transpose(dmap.view(B, 4096, 64)).view(B, 64 * 4096)
        d = self.g0.bwd(self.g_ln.bwd(self.g3.bwd(dg)))
        dcat = self.af0.bwd(self.af_ln.bwd(self.af3.bwd(d)))
        self.pos0.bwd(self.pos_ln.bwd(self.pos3.bwd(dcat[:, 128:])), need_dx=False)
        datt0 = self.to_out.bwd(dcat[:, :128])
        dqkv = ops.l2l_row0_bwd(self.qkv_out, self.lane_num, self.prob, datt0, 2)
        dtok = self.qkv.bwd(dqkv.view(G, 384))
        d = ops.segmax_bwd(dtok, self.argf, G, V)
        for i in (2, 1, 0):
            lin, ln = self.sub[i]
            d = ops.subgraph_pool_bwd(d, self.args[i], G, V)
            d = lin.bwd(ln.bwd(d), need_dx=(i > 0))
        self.args = self.qkv_out = self.prob = None


class SpGAT:
    """Radar graph-attention encoder (model_rad.py:778-884)."""

    def __init__(self, st, prefix, cfg):
        self.nh, self.hid, self.alpha, self.p = cfg.nb_heads, cfg.hidden, cfg.alpha, cfg.attn_pdrop
        self.W = [(st.p(f"{prefix}.attention_{i}.W"), st.g(f"{prefix}.attention_{i}.W")) for i in range(self.nh)]
        self.a = [(st.p(f"{prefix}.attention_{i}.a"), st.g(f"{prefix}.attention_{i}.a")) for i in range(self.nh)]
        self.m1, self.m2 = Linear(st, prefix + ".mlp_1.0"), Linear(st, prefix + ".mlp_2.0")

    def fwd(self, radar, adj, seed, train):
        B, N, nh, hid = radar.shape[0], radar.shape[1], self.nh, self.hid
        dev = radar.device
        p = self.p if train else 0.0
        self.pp, self.seed, self.adj, self.B = p, seed, adj, B
        self.x = ops.dropout(radar, p, seed)
        E = 2 * hid
        self.Wh = torch.empty((nh, B, N, E), device=dev, dtype=torch.float32)
        self.z = torch.empty((nh, B, N, N), device=dev, dtype=torch.float32)
        hp = torch.empty((nh, B, N, E), device=dev, dtype=torch.float32)
        self.att, self.attd = [], []
        for i in range(nh):
            ops.gemm(self.x, self.W[i][0].t().unsqueeze(0), self.Wh[i])
            ops.gemm(self.Wh[i], self.a[i][0].t().unsqueeze(0), self.z[i])
            att, attd = ops.gat_softmax_fwd(self.z[i], adj, self.alpha, p, seed + 1 + i)
            self.att.append(att)
            self.attd.append(attd)
            ops.gemm(attd, self.Wh[i].transpose(-1, -2), hp[i])
        self.e1 = ops.elu_fwd(hp)
        e1d = ops.dropout(self.e1, p, seed + 5)
        self.e2 = ops.elu_fwd(e1d)
        o1 = torch.empty((B, nh * N, 256), device=dev, dtype=torch.float32)
        for i in range(nh):
            ops.gemm(self.e2[i], self.m1.w.unsqueeze(0), o1[:, i * N:(i + 1) * N, :], bias=self.m1.b)
        self.o1d = ops.dropout(o1, p, seed + 6)
        o2 = torch.empty((B, 256, 128), device=dev, dtype=torch.float32)
        ops.gemm(self.o1d.transpose(1, 2), self.m2.w.unsqueeze(0), o2, bias=self.m2.b)
        o2d = ops.dropout(o2, p, seed + 7)
        self.y = ops.radar_logsoftmax_fwd(o2d, B, 512)
        return self.y

    def bwd(self, dy):
        B, nh, N = self.B, self.nh, self.e2.shape[2]
        p, seed = self.pp, self.seed
        do2 = ops.dropout(ops.radar_logsoftmax_bwd(dy, self.y).view(B, 256, 128), p, seed + 7)
        # o2 = o1d^T W2^T + b2
        ops.colsum_(do2.view(B * 256, 128), self.m2.db)
        ops.gemm(do2.transpose(1, 2), self.o1d, self.m2.dw.unsqueeze(0).expand(B, -1, -1), accum=2)   # sum over b
        do1dT = torch.empty((B, 256, nh * N), device=dy.device, dtype=torch.float32)
        ops.gemm(do2, self.m2.w.t().unsqueeze(0), do1dT)
        do1 = ops.dropout(ops.transpose(do1dT), p, seed + 6)                  # (B, nh*N, 256)
        ops.colsum_(do1.view(B * nh * N, 256), self.m1.db)
        de2 = torch.empty_like(self.e2)
        for i in range(nh):
            sl = do1[:, i * N:(i + 1) * N, :]
            ops.gemm(sl.transpose(1, 2), self.e2[i].transpose(1, 2), self.m1.dw.unsqueeze(0).expand(B, -1, -1), accum=2)
            ops.gemm(sl, self.m1.w.t().unsqueeze(0), de2[i])
        de1 = ops.elu_bwd(ops.dropout(ops.elu_bwd(de2, self.e2), p, seed + 5), self.e1)   # (nh,B,N,E)
        for i in range(nh):
            Wh, attd = self.Wh[i], self.attd[i]
            dattd = torch.empty_like(attd)
            ops.gemm(de1[i], Wh, dattd)                                        # d(attd) = dhp Wh^T
            dWh = torch.empty_like(Wh)
            ops.gemm(attd.transpose(-1, -2), de1[i].transpose(-1, -2), dWh)    # attd^T dhp
            dz = ops.gat_softmax_bwd(self.z[i], self.adj, self.att[i], dattd, self.alpha, p, seed + 1 + i)
            E = Wh.shape[-1]
            ops.gemm(Wh.reshape(B * N, E).t(), dz.reshape(B * N, N).t(), self.a[i][1], accum=1)   # da = Wh^T dz
            ops.gemm(dz, self.a[i][0].unsqueeze(0), dWh, accum=1)              # dWh += dz a^T
            ops.gemm(self.x.reshape(B * N, -1).t(), dWh.reshape(B * N, E).t(), self.W[i][1], accum=1)   # dW = x^T dWh
        self.Wh = self.z = self.att = self.attd = self.e1 = self.e2 = self.o1d = self.y = None


class Head:
    """join MLP + GRU waypoint roll-out + L1 loss (model_rad.py:655-695, phase2_train_net.py:104)."""

    def __init__(self, st, pred_len):
        self.j = [Linear(st, f"join.{i}") for i in (0, 2, 4)]
        names = ("decoder.weight_ih", "decoder.weight_hh", "decoder.bias_ih", "decoder.bias_hh", "output.weight", "output.bias")
        self.P = [st.p(n) for n in names]
        self.G = [st.g(n) for n in names]
        self.steps = pred_len

    def fwd(self, fused, target_point):
        z = fused
        for lin in self.j:
            z = lin.fwd(z, act=1)
        pred, self.ctx = ops.gru_head_fwd(z, target_point, *self.P, self.steps)
        return pred

    def bwd(self, dpred):
        P, G = self.P, self.G
        d = ops.gru_head_bwd(dpred, self.ctx, P[0], P[1], P[4], G[0], G[1], G[2], G[3], G[4], G[5])
        for lin in reversed(self.j):
            d = lin.bwd(d)
        self.ctx = None
        return d


# --------------------------------------------------------------------------- the network
class _Net:
    """Kernel schedule of Encoder.forward (model_rad.py:492-611) + head, forward and backward.
    VARIANT "rad": camera + LiDAR + VectorNet map + radar GAT (model_rad.py);  "vec": the same without the radar
    branch (model_vec.py:488-600);  "img": the map is a rasterised image run through the map ResNet's own stem and
    layer1 instead of VectorNet, no radar (model_img.py:310-423)."""
    VARIANT = "rad"

    def __init__(self, st, cfg):
        e = "encoder."
        self.cfg = cfg
        var = self.VARIANT
        ip, mp, lp = e + "image_encoder.features", e + "img_map_encoder.features", e + "lidar_encoder._model"
        self.img_stem, self.lid_stem = Stem(st, ip), Stem(st, lp)
        self.img_layers = [ResLayer(st, f"{ip}.layer{i + 1}", RESNET34[i], 1 if i == 0 else 2) for i in range(4)]
        self.lid_layers = [ResLayer(st, f"{lp}.layer{i + 1}", RESNET18[i], 1 if i == 0 else 2) for i in range(4)]
        self.map_layers = [None] + [ResLayer(st, f"{mp}.layer{i + 1}", RESNET34[i], 2) for i in range(1, 4)]
        self.map_stem = None
        if var == "img":
            self.map_stem = Stem(st, mp)
            self.map_layers[0] = ResLayer(st, f"{mp}.layer1", RESNET34[0], 1)
        self.vectornet = VectorNet(st, e + "vectornet_encoder") if var != "img" else None
        self.gat = SpGAT(st, e + "radar_encoder", cfg) if var == "rad" else None
        self.gpts = [FusionGPT(st, f"{e}transformer{i + 1}", WIDTHS[i], 3 if (i < 3 or var != "rad") else 4, cfg, i) for i in range(4)]
        self.head = Head(st, cfg.pred_len)
        dev = st.device
        self.side = [torch.cuda.Stream(device=dev, priority=-1) for _ in range(3)]
        self.use_streams = True
        self.mean = torch.tensor(IMAGENET_MEAN, device=dev, dtype=torch.float32)
        self.std = torch.tensor(IMAGENET_STD, device=dev, dtype=torch.float32)

    # ---- stream-level parallelism -----------------------------------------------------------
    # Between two fusion points the image / LiDAR / map trunks (and VectorNet, the radar GAT) are
    # independent, and at B=16 most of their kernels fill only part of the 148 SMs.  They are issued on
    # side streams forked from / joined into the main stream with events, so the CUDA graph captured
    # by the engine contains parallel branches instead of one 1900-node chain.
    def _fork(self, n):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        for s in self.side[:n]:
            s.wait_event(ev)

    def _join(self, n):
        main = torch.cuda.current_stream()
        for s in self.side[:n]:
            ev = torch.cuda.Event()
            ev.record(s)
            main.wait_event(ev)

    def _parallel(self, *fns):
        """Run fns[0] on the main stream and fns[1:] on side streams; returns their results in order."""
        if not self.use_streams or len(fns) == 1:
            return [f() for f in fns]
        n = len(fns) - 1
        self._fork(n)
        out = [None] * len(fns)
        for i, f in enumerate(fns[1:]):
            with torch.cuda.stream(self.side[i]):
                out[i + 1] = f()
        out[0] = fns[0]()
        self._join(n)
        return out

    def forward(self, image, lidar, lane, lane_num, radar, radar_adj, target_point, velocity, seed, train):
        """`lane` is the padded lane tensor (rad / vec) or the rasterised map image (B,3,256,256) (img).  `lidar` and
        `radar_adj` may be zero-argument callables (the engine's BEV scatter / histogram unpacking / adjacency kernels):
        they are then evaluated INSIDE the LiDAR / radar branch, off the image trunk's stream."""
        val = lambda t: t() if callable(t) else t
        branches = [
            lambda: self.img_layers[0].fwd(self.img_stem.fwd(ops.nchw_to_nhwc(image, self.mean, self.std), train), train),
            lambda: self.lid_layers[0].fwd(self.lid_stem.fwd(ops.nchw_to_nhwc(val(lidar)), train), train)]
        if self.map_stem is not None:          # model_img.py:337-340, :348 -- the map image is NOT ImageNet-normalised
            branches.append(lambda: self.map_layers[0].fwd(self.map_stem.fwd(ops.nchw_to_nhwc(lane), train), train))
        else:
            branches.append(lambda: self.vectornet.fwd(lane, lane_num))
        if self.gat is not None:
            branches.append(lambda: self.gat.fwd(radar, val(radar_adj), seed + 900, train))
        out = self._parallel(*branches)
        img, lid, mp = out[:3]
        self._run_idle_hook()
        for s in range(3):
            tok = self.gpts[s].fwd([img, lid, mp], velocity, seed, train)
            img, lid, mp = self._parallel(
                lambda: self.img_layers[s + 1].fwd(ops.upsample_add_fwd(img, tok, 0), train),
                lambda: self.lid_layers[s + 1].fwd(ops.upsample_add_fwd(lid, tok, 1), train),
                lambda: self.map_layers[s + 1].fwd(ops.upsample_add_fwd(mp, tok, 2), train))
        feats = [img, lid, mp] + ([out[3]] if self.gat is not None else [])
        tok = self.gpts[3].fwd(feats, velocity, seed, train)
        fused = ops.pool_sum_fwd(feats, tok)
        return self.head.fwd(fused, target_point)

    mid_hook = None
    # Work the step needs before BACKWARD but not for the forward (the engine: zeroing the 420 MB gradient buffer).  It is
    # issued on its own stream once the stems and layer 1 are done -- the HBM-bound head of the step is over and the
    # latency-bound first fusion transformer leaves the memory system idle -- and joined at the start of backward().
    idle_hook = None
    _idle_done = None

    def _run_idle_hook(self):
        self._idle_done = None
        if self.idle_hook is None:
            return
        if not self.use_streams:
            self.idle_hook()
            return
        if getattr(self, "_idle_stream", None) is None:
            self._idle_stream = torch.cuda.Stream(device=self.mean.device, priority=0)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self._idle_stream.wait_event(ev)
        with torch.cuda.stream(self._idle_stream):
            self.idle_hook()
            self._idle_done = torch.cuda.Event()
            self._idle_done.record(self._idle_stream)

    def _bucket_done(self, which):
        """Backward has passed a fusion stage: every gradient of params.is_early_bucket() ("early": after the last
        stage) / params.is_mid_bucket() ("mid": after transformer2's backward) is final once the leaf streams are
        joined.  The engine hooks in here to split the captured step into three graphs and start that bucket's
        all-reduce + AdamW under the rest of backward."""
        if self.mid_hook is not None:
            _Aux.join_all()
            self.mid_hook(which)

    def _early_bucket_done(self):
        self._bucket_done("early")

    def backward(self, dpred):
        if self._idle_done is not None:
            torch.cuda.current_stream().wait_event(self._idle_done)
            self._idle_done = None
        dfused = self.head.bwd(dpred)
        nmod = 4 if self.gat is not None else 3
        dfe, dtok = ops.pool_sum_bwd(dfused, nmod)
        self.gpts[3].bwd(dtok, dfe)
        dimg, dlid, dmp = dfe[0], dfe[1], dfe[2]
        for s in (2, 1, 0):
            B, C = dimg.shape[0], dimg.shape[3] // 2
            dtok = torch.empty((B, 192, C), device=dimg.device, dtype=torch.float32)

            def trunk(layers, d, m):
                def run():
                    g = layers[s + 1].bwd(d)
                    ops.upsample_add_bwd_(g, dtok, m)          # disjoint 64-token slices of dtok
                    return g
                return run
            branches = [trunk(self.img_layers, dimg, 0), trunk(self.lid_layers, dlid, 1), trunk(self.map_layers, dmp, 2)]
            if s == 2 and self.gat is not None:
                branches.append(lambda: self.gat.bwd(dfe[3]))  # same side stream as its forward
            if s < 2:
                _Aux.defer_begin()                             # layer3 / layer2 weight gradients: under the next GPT backward
            dimg, dlid, dmp = self._parallel(*branches)[:3]
            if s < 2:
                _Aux.defer_flush()
            if s == 2:
                self._early_bucket_done()
            self.gpts[s].bwd(dtok, [dimg, dlid, dmp])
            if s == 1:
                self._bucket_done("mid")                       # layer3 x3 (deferred leaves ran under transformer2), transformer3 / 2
        self._parallel(lambda: self.img_stem.bwd(self.img_layers[0].bwd(dimg)),
                       lambda: self.lid_stem.bwd(self.lid_layers[0].bwd(dlid)),
                       (lambda: self.map_stem.bwd(self.map_layers[0].bwd(dmp))) if self.map_stem is not None
                       else (lambda: self.vectornet.bwd(dmp)))
        _Aux.join_all()


class _NetVec(_Net):
    VARIANT = "vec"


class _NetImg(_Net):
    VARIANT = "img"


class _WholeNet(torch.autograd.Function):
    """Lets `loss.backward()` (Engine.train, phase2_train_net.py:108) drive the hand-written backward."""

    @staticmethod
    def forward(ctx, model, inputs, *params):
        ctx.model = model
        with torch.no_grad():
            return model._forward_impl(*inputs)

    @staticmethod
    def backward(ctx, dpred):
        m = ctx.model
        m.store.flat_grad.zero_()
        m.net.backward(dpred.contiguous())
        grads = []
        for k, p in m._param_items:
            grads.append(None if is_unused(k, m.VARIANT) else m.store.torch_view(k, grad=True).clone())
        return (None, None, *grads)


class PIDController:
    def __init__(self, K_P=1.0, K_I=0.0, K_D=0.0, n=20):
        self._K_P, self._K_I, self._K_D = K_P, K_I, K_D
        self._window = deque([0 for _ in range(n)], maxlen=n)

    def step(self, error):
        self._window.append(error)
        if len(self._window) >= 2:
            integral, derivative = np.mean(self._window), self._window[-1] - self._window[-2]
        else:
            integral, derivative = 0.0, 0.0
        return self._K_P * error + self._K_I * integral + self._K_D * derivative


class MMFN(nn.Module):
    VARIANT = "rad"          # parameter inventory (params.param_spec) and kernel schedule of this model variant
    NET = _Net

    def __init__(self, config, device):
        super().__init__()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ops.MmfnError("mmfn_b200.MMFN runs on CUDA (sm_100a) only; there is no CPU path")
        self.config = config
        self.pred_len = config.pred_len
        self.turn_controller = PIDController(config.turn_KP, config.turn_KI, config.turn_KD, config.turn_n)
        self.speed_controller = PIDController(config.speed_KP, config.speed_KI, config.speed_KD, config.speed_n)
        self.store = ParamStore(config, self.device, self.VARIANT)
        self.store.register(self)
        self._param_items = list(self.named_parameters())
        self.net = self.NET(self.store, config)
        self.seed = 0
        self.reset_parameters()
        self._maybe_load_pretrained()

    # ---- initialisation following the reference constructors ---------------------------------
    @torch.no_grad()
    def reset_parameters(self, seed=42):
        g = torch.Generator().manual_seed(seed)
        for k, p in self._param_items:
            shape, leaf = tuple(p.shape), k.split(".")[-1]
            if p.dim() == 4:                                          # torchvision: kaiming_normal_(fan_out, relu)
                t = torch.randn(shape, generator=g) * math.sqrt(2.0 / (shape[0] * shape[2] * shape[3]))
            elif leaf == "pos_emb":
                t = torch.zeros(shape)
            elif leaf in ("W", "a"):                                  # xavier_normal_(gain=1.414), model_rad.py:790-794
                t = torch.randn(shape, generator=g) * 1.414 * math.sqrt(2.0 / (shape[0] + shape[1]))
            elif ".transformer" in k:                                 # GPT._init_weights, model_rad.py:170-177
                t = torch.ones(shape) if (".ln" in k and leaf == "weight") else (
                    torch.randn(shape, generator=g) * 0.02 if p.dim() == 2 else torch.zeros(shape))
            elif p.dim() == 2:                                        # nn.Linear / GRUCell default
                fan_in = 64 if k.startswith("decoder") else shape[1]
                t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
            else:
                is_norm = leaf == "weight"
                if is_norm:
                    t = torch.ones(shape)
                elif ".bn" in k or "downsample.1" in k or _is_ln_bias(k):
                    t = torch.zeros(shape)
                else:
                    fan_in = 64 if k.startswith("decoder") else self._fan_in_of_bias(k)
                    t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(fan_in)
            p.copy_(t.to(self.device))
        for k, b in self.named_buffers():
            if k.endswith("running_mean"):
                b.zero_()
            elif k.endswith("running_var"):
                b.fill_(1.0)
            else:
                b.zero_()

    # ---- ImageNet initialisation of the two ResNet-34 trunks --------------------------------
    PRETRAINED_FILE = "resnet34-b627a593.pth"           # what models.resnet34(pretrained=True) downloads

    @torch.no_grad()
    def load_torchvision_resnet34(self, weights, trunks=("image_encoder", "img_map_encoder")):
        """The reference builds both ResNet-34 trunks as `models.resnet34(pretrained=True)` with `fc` removed
        (ImageCNN, model_rad.py:22-23; image_encoder / img_map_encoder :427-428).  `weights`: a torchvision resnet34
        state_dict or the path of its checkpoint file.  Copies every tensor but `fc.*` into
        encoder.{image_encoder,img_map_encoder}.features.* (the RGB+LiDAR variant has only image_encoder)."""
        if isinstance(weights, (str, bytes)) or hasattr(weights, "__fspath__"):
            weights = torch.load(weights, map_location="cpu")
        own = self.state_dict()
        part = {}
        for trunk in trunks:
            prefix = f"encoder.{trunk}.features."
            if not any(k.startswith(prefix) for k in own):
                continue
            for k, v in weights.items():
                if k.startswith("fc."):
                    continue
                if prefix + k not in own or tuple(own[prefix + k].shape) != tuple(v.shape):
                    raise ops.MmfnError(f"load_torchvision_resnet34: {k} does not fit {prefix}{k}")
                part[prefix + k] = v
        if not part:
            raise ops.MmfnError("load_torchvision_resnet34: no ResNet-34 trunk in this model")
        self.load_state_dict(part, strict=False)
        return sorted(part)

    def _maybe_load_pretrained(self):
        """Reference behaviour when the ImageNet checkpoint is available locally (torch hub cache or
        $MMFN_RESNET34_WEIGHTS); otherwise the trunks keep the seeded kaiming init -- there is no network here."""
        import os
        cands = [os.environ.get("MMFN_RESNET34_WEIGHTS"),
                 os.path.join(torch.hub.get_dir(), "checkpoints", self.PRETRAINED_FILE)]
        for c in cands:
            if c and os.path.isfile(c):
                self.load_torchvision_resnet34(c)
                self.pretrained_from = c
                return
        self.pretrained_from = None

    def load_state_dict(self, state_dict, strict=True, **kw):
        """nn.Module.load_state_dict + refresh of the bf16 weight shadow the bf16 tensor-core kernels read."""
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self.store.sync_shadow()
        return out

    def _fan_in_of_bias(self, k):
        w = k[: -len("bias")] + "weight"
        return self.store.shapes[w][1] if w in self.store.shapes else 64

    # ---- reference surface -----------------------------------------------------------------------
    def forward(self, image_list, lidar_list, map_list, vectormaps_list, radar_list, radar_adj, target_point, velocity):
        """Reference signature (model_rad.py:666, identical in model_vec.py:653 and model_img.py).  Inputs a variant
        does not consume (maps_list for rad / vec, vectormaps for img, radar for vec / img) are ignored and may be None."""
        if self.VARIANT == "img":
            lane, lane_num = map_list[0], None
        else:
            lane, lane_num = vectormaps_list[0][0], vectormaps_list[1][0]
        radar, adj = (radar_list[0], radar_adj[0]) if self.VARIANT == "rad" else (None, None)
        inputs = (image_list[0], lidar_list[0], lane, lane_num, radar, adj, target_point, velocity)
        if torch.is_grad_enabled() and self.training:
            return _WholeNet.apply(self, inputs, *[p for _, p in self._param_items])
        with torch.no_grad():
            return self._forward_impl(*inputs)

    def _forward_impl(self, image, lidar, lane, lane_num, radar, radar_adj, target_point, velocity):
        f32 = lambda t: t.to(device=self.device, dtype=torch.float32).contiguous()
        if image.dtype == torch.uint8:                     # camera bytes are converted inside the layout kernel
            image = image.to(self.device).contiguous()
        else:
            image = f32(image)
        lidar, lane, target_point, velocity = map(f32, (lidar, lane, target_point, velocity))
        if radar is not None:
            radar, radar_adj = f32(radar), f32(radar_adj)
        if lane_num is not None:
            lane_num = lane_num.to(device=self.device, dtype=torch.int32).contiguous()
        if self.training:
            self.seed += 1000
            self.store.flat_nbt.add_(self._nbt_step())     # BatchNorm.num_batches_tracked
        if ops.BF16:                                       # a caller-side optimizer (torch.optim) moves only the fp32 masters
            self.store.sync_shadow()
        return self.net.forward(image, lidar, lane, lane_num, radar, radar_adj, target_point, velocity,
                                self.seed, self.training)

    def _nbt_step(self):
        """+1 for every BatchNorm that runs; the map ResNet's stem/layer1 BNs never do (stay 0)."""
        if not hasattr(self, "_nbt_inc"):
            inc = [0 if is_unused(k, self.VARIANT) else 1 for k in self.store.nbt_index]
            self._nbt_inc = torch.tensor(inc, device=self.device, dtype=torch.int64)
        return self._nbt_inc

    def control_pid(self, waypoints, velocity):
        """PID steering/throttle from predicted waypoints (model_rad.py:697-739); CPU, inference only."""
        assert waypoints.size(0) == 1
        wp = waypoints[0].data.cpu().numpy()
        wp[:, 1] *= -1                                     # forward is negative y in the waypoint frame
        speed = velocity[0].data.cpu().numpy()
        cfg = self.config
        desired_speed = np.linalg.norm(wp[0] - wp[1]) * 2.0
        brake = desired_speed < cfg.brake_speed or (speed / desired_speed) > cfg.brake_ratio
        aim = (wp[1] + wp[0]) / 2.0
        angle = np.degrees(np.pi / 2 - np.arctan2(aim[1], aim[0])) / 90
        if speed < 0.01:
            angle = np.array(0.0)
        steer = np.clip(self.turn_controller.step(angle), -1.0, 1.0)
        delta = np.clip(desired_speed - speed, 0.0, cfg.clip_delta)
        throttle = np.clip(self.speed_controller.step(delta), 0.0, cfg.max_throttle)
        throttle = throttle if not brake else 0.0
        metadata = {
            "speed": float(speed.astype(np.float64)), "steer": float(steer), "throttle": float(throttle),
            "brake": float(brake), "wp_2": tuple(wp[1].astype(np.float64)), "wp_1": tuple(wp[0].astype(np.float64)),
            "desired_speed": float(desired_speed.astype(np.float64)), "angle": float(angle.astype(np.float64)),
            "aim": tuple(aim.astype(np.float64)), "delta": float(delta.astype(np.float64)),
        }
        return steer, throttle, brake, metadata


class MMFNVec(MMFN):
    """Drop-in for team_code/mmfn_utils/models/model_vec.py:MMFN (camera + LiDAR + VectorNet map, no radar)."""
    VARIANT = "vec"
    NET = _NetVec


class MMFNImg(MMFN):
    """Drop-in for team_code/mmfn_utils/models/model_img.py:MMFN -- the DEFAULT train.yaml entry point
    (run_steps/config/train.yaml:13): camera + LiDAR + rasterised map image."""
    VARIANT = "img"
    NET = _NetImg


def _is_ln_bias(k):
    return any(s in k for s in ("mlp.1.bias", "pos_emb.1.bias", "agent_fusion.1.bias", "generator.1.bias"))
