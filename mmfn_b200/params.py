"""Parameter inventory and flat HBM layout of MMFN (model_rad variant).

`param_spec()` reproduces the reference state_dict (key order, shapes, dtypes) of
team_code/mmfn_utils/models/model_rad.py:MMFN without importing it; `ParamStore` places every
tensor in ONE flat fp32 buffer (plus a mirror for gradients) so the optimiser and the gradient
all-reduce are single launches over contiguous memory:

  [ trained parameters ... | parameters that never receive a gradient ]   (each 16-byte aligned)

Convolution filters are stored KRSC (what the NHWC implicit-GEMM kernels read) and exposed to
PyTorch as the channels_last-strided (K,C,R,S) view, so state_dicts stay interchangeable with the
reference.  key/query/value weights (and biases) of a block are adjacent so QKV is one GEMM.
"""
import torch
from torch import nn

RESNET34 = (3, 4, 6, 3)
RESNET18 = (2, 2, 2, 2)
WIDTHS = (64, 128, 256, 512)


def _bn(prefix, c):
    return [(prefix + ".weight", (c,), "f"), (prefix + ".bias", (c,), "f"),
            (prefix + ".running_mean", (c,), "buf"), (prefix + ".running_var", (c,), "buf"),
            (prefix + ".num_batches_tracked", (), "nbt")]


def _resnet(prefix, blocks, in_ch):
    spec = [(prefix + ".conv1.weight", (64, in_ch, 7, 7), "conv")] + _bn(prefix + ".bn1", 64)
    cin = 64
    for li, (n, c) in enumerate(zip(blocks, WIDTHS), start=1):
        for b in range(n):
            p = f"{prefix}.layer{li}.{b}"
            spec.append((p + ".conv1.weight", (c, cin if b == 0 else c, 3, 3), "conv"))
            spec += _bn(p + ".bn1", c)
            spec.append((p + ".conv2.weight", (c, c, 3, 3), "conv"))
            spec += _bn(p + ".bn2", c)
            if b == 0 and li > 1:
                spec.append((p + ".downsample.0.weight", (c, cin, 1, 1), "conv"))
                spec += _bn(p + ".downsample.1", c)
        cin = c
    return spec


def _linear(prefix, out_f, in_f, bias=True):
    s = [(prefix + ".weight", (out_f, in_f), "f")]
    if bias:
        s.append((prefix + ".bias", (out_f,), "f"))
    return s


def _ln(prefix, c):
    return [(prefix + ".weight", (c,), "f"), (prefix + ".bias", (c,), "f")]


def _gpt(prefix, c, n_tok, n_layer, block_exp):
    spec = [(prefix + ".pos_emb", (1, n_tok, c), "f")] + _linear(prefix + ".vel_emb", c, 1)
    for i in range(n_layer):
        b = f"{prefix}.blocks.{i}"
        spec += _ln(b + ".ln1", c) + _ln(b + ".ln2", c)
        for nm in ("key", "query", "value", "proj"):
            spec += _linear(f"{b}.attn.{nm}", c, c)
        spec += _linear(b + ".mlp.0", block_exp * c, c) + _linear(b + ".mlp.2", c, block_exp * c)
    return spec + _ln(prefix + ".ln_f", c)


def _head_spec():
    spec = _linear("join.0", 256, 512) + _linear("join.2", 128, 256) + _linear("join.4", 64, 128)
    spec += [("decoder.weight_ih", (192, 2), "f"), ("decoder.weight_hh", (192, 64), "f"),
             ("decoder.bias_ih", (192,), "f"), ("decoder.bias_hh", (192,), "f")]
    return spec + _linear("output", 2, 64)


def transfuser_param_spec(cfg):
    """state_dict of benchmarks/transfuser/model.py:TransFuser (:403-427): RGB + LiDAR trunks, four GPTs over
    (n_views + 1) * seq_len * 64 tokens (:147), join MLP, GRU cell, output layer."""
    e = "encoder."
    spec = _resnet(e + "image_encoder.features", RESNET34, 3)
    spec += _resnet(e + "lidar_encoder._model", RESNET18, 2)
    ntok = (cfg.n_views + 1) * cfg.seq_len * cfg.vert_anchors * cfg.horz_anchors
    for i, c in enumerate(WIDTHS, start=1):
        spec += _gpt(f"{e}transformer{i}", c, ntok, cfg.n_layer, cfg.block_exp)
    return spec + _head_spec()


def param_spec(cfg, variant="rad"):
    """[(key, shape, kind)] in reference state_dict order. kind: conv | f | buf | nbt.
    variant: "rad" (model_rad.py: camera + LiDAR + VectorNet map + radar GAT), "vec" (model_vec.py: no radar),
    "img" (model_img.py: rasterised map image through the map ResNet, no VectorNet, no radar),
    "transfuser" (benchmarks/transfuser/model.py: camera + LiDAR only)."""
    if variant == "transfuser":
        return transfuser_param_spec(cfg)
    assert variant in ("rad", "vec", "img"), variant
    e = "encoder."
    spec = _resnet(e + "image_encoder.features", RESNET34, 3)
    spec += _resnet(e + "img_map_encoder.features", RESNET34, 3)
    spec += _resnet(e + "lidar_encoder._model", RESNET18, 2)
    if variant != "img":
        v = e + "vectornet_encoder"
        cin = 7
        for i in range(3):
            spec += _linear(f"{v}.lane_subgraph.layers.mlp_{i}.mlp.0", 64, cin) + _ln(f"{v}.lane_subgraph.layers.mlp_{i}.mlp.1", 64)
            cin = 128
        spec += _linear(v + ".pos_emb.0", 64, 2) + _ln(v + ".pos_emb.1", 64) + _linear(v + ".pos_emb.3", 64, 64)
        spec += _linear(v + ".L2L.to_qkv", 384, 128, bias=False) + _linear(v + ".L2L.to_out.0", 128, 128)
        spec += _linear(v + ".agent_fusion.0", 128, 192) + _ln(v + ".agent_fusion.1", 128) + _linear(v + ".agent_fusion.3", 128, 128)
        spec += _linear(v + ".generator.0", 64, 128) + _ln(v + ".generator.1", 64) + _linear(v + ".generator.3", 64 * 64 * 64, 64)
    if variant == "rad":
        r = e + "radar_encoder"
        nh, hid = cfg.nb_heads, cfg.hidden
        for i in range(nh):
            spec += [(f"{r}.attention_{i}.W", (5, 2 * hid), "f"), (f"{r}.attention_{i}.a", (2 * hid, hid), "f")]
        spec += _linear(r + ".mlp_1.0", 256, nh * hid) + _linear(r + ".mlp_2.0", 128, nh * hid)
    ntok = (cfg.n_views + 2) * cfg.seq_len * cfg.vert_anchors * cfg.horz_anchors
    for i, c in enumerate((64, 128, 256), start=1):
        spec += _gpt(f"{e}transformer{i}", c, ntok, cfg.n_layer, cfg.block_exp)
    ntok4 = ntok + (cfg.seq_len * cfg.vert_anchors * cfg.horz_anchors if variant == "rad" else 0)
    spec += _gpt(e + "transformer4", 512, ntok4, cfg.n_layer, cfg.block_exp)
    return spec + _head_spec()


def is_unused(key, variant="rad"):
    """Parameters of the map ResNet that model_rad / model_vec never touch (stem + layer1; the VectorNet output
    enters at layer2): they get no gradient in the reference, so torch.optim skips them entirely (no weight decay
    either).  model_img runs the whole map ResNet; the TransFuser variant has none."""
    if variant in ("img", "transfuser"):
        return False
    p = "encoder.img_map_encoder.features."
    return key.startswith(p) and (key[len(p):].startswith(("conv1.", "bn1.", "layer1.")))


EARLY_BUCKET_PREFIXES = ("encoder.image_encoder.features.layer4.", "encoder.img_map_encoder.features.layer4.",
                         "encoder.lidar_encoder._model.layer4.", "encoder.radar_encoder.", "encoder.transformer4.",
                         "join.", "decoder.", "output.")


def is_early_bucket(key):
    """Parameters whose gradients are complete once backward has passed the last fusion stage (model_rad.py:577-611
    in reverse): the waypoint head, transformer4, the radar GAT and layer4 of the three ResNets."""
    return key.startswith(EARLY_BUCKET_PREFIXES)


MID_BUCKET_PREFIXES = ("encoder.image_encoder.features.layer3.", "encoder.img_map_encoder.features.layer3.",
                       "encoder.lidar_encoder._model.layer3.", "encoder.transformer3.", "encoder.transformer2.")


def is_mid_bucket(key):
    """Parameters whose gradients are complete once backward has passed the SECOND fusion stage (layer3 of every
    trunk, transformer3, transformer2): the second overlap-able range of the flat buffer."""
    return key.startswith(MID_BUCKET_PREFIXES)


def _numel(shape):
    n = 1
    for s in shape:
        n *= s
    return n


class ParamStore:
    def __init__(self, cfg, device, variant="rad"):
        self.spec = param_spec(cfg, variant)
        self.variant = variant
        self.device = torch.device(device)
        fkeys = [(k, s) for k, s, kind in self.spec if kind in ("conv", "f")]
        order = self._flat_order([k for k, _ in fkeys])
        shapes = dict(fkeys)
        self.offsets = {}
        off = 0
        # flat layout: [late bucket | mid bucket | early bucket | never-used].  "early" = parameters whose gradients are
        # final first in backward (head, transformer4, radar encoder, layer4 of every trunk), "mid" = final after the
        # second fusion stage (layer3 of every trunk, transformer3, transformer2): each is ONE contiguous range --
        # early [n_mid, n_active), mid [n_late, n_mid) -- so its all-reduce + AdamW can start while the rest of
        # backward still runs; only the late range [0, n_late) (layers 1-2, stems, transformer1, VectorNet: ~20 % of
        # the gradient bytes) is exchanged after backward.
        for group in ("late", "mid", "early", "unused"):
            if group == "mid":
                self.n_late = off
            if group == "early":
                self.n_mid = off
            if group == "unused":
                self.n_active = off
            for k in order:
                g = "unused" if is_unused(k, variant) else ("early" if is_early_bucket(k) else ("mid" if is_mid_bucket(k) else "late"))
                if g != group:
                    continue
                self.offsets[k] = off
                off += (_numel(shapes[k]) + 3) // 4 * 4
        self.n_total = off
        self.shapes = shapes
        self.kinds = {k: kind for k, _, kind in self.spec}
        self.flat = torch.zeros(self.n_total, device=self.device, dtype=torch.float32)
        self.flat_grad = torch.zeros(self.n_total, device=self.device, dtype=torch.float32)
        # bf16 shadow of the master weights (same offsets): the B operand of every bf16 tensor-core kernel
        # (BASELINE configs[2]).  Refreshed by the AdamW kernel (engine) or sync_shadow() (after a state_dict load).
        self.flat16 = torch.zeros(self.n_total, device=self.device, dtype=torch.bfloat16) if self.device.type == "cuda" else None
        bufs = [(k, s) for k, s, kind in self.spec if kind == "buf"]
        self.buf_offsets, boff = {}, 0
        for k, s in bufs:
            self.buf_offsets[k] = boff
            boff += _numel(s)
        self.flat_buf = torch.zeros(boff, device=self.device, dtype=torch.float32)
        nbt = [k for k, _, kind in self.spec if kind == "nbt"]
        self.nbt_index = {k: i for i, k in enumerate(nbt)}
        self.flat_nbt = torch.zeros(len(nbt), device=self.device, dtype=torch.int64)

    @staticmethod
    def _flat_order(keys):
        """reference order, except that each block's key/query/value weights, then biases, are adjacent."""
        out, done = [], set()
        for k in keys:
            if k in done:
                continue
            if k.endswith(".attn.key.weight"):
                base = k[: -len("key.weight")]
                grp = [base + f"{n}.{t}" for t in ("weight", "bias") for n in ("key", "query", "value")]
                out += grp
                done.update(grp)
            else:
                out.append(k)
                done.add(k)
        return out

    # native (kernel-facing) views ---------------------------------------------------------
    def _native(self, flat, key):
        shape, off = self.shapes[key], self.offsets[key]
        t = flat[off: off + _numel(shape)]
        if self.kinds[key] == "conv":
            co, c, r, s = shape
            return t.view(co, r, s, c)
        return t.view(shape)

    def p(self, key):
        return self._native(self.flat, key)

    def g(self, key):
        return self._native(self.flat_grad, key)

    def p16(self, key):
        """bf16 shadow of p(key) (None on CPU stores)"""
        return None if self.flat16 is None else self._native(self.flat16, key)

    def fused16(self, keys):
        if self.flat16 is None:
            return None
        off = self.offsets[keys[0]]
        n = sum(_numel(self.shapes[k]) for k in keys)
        sh = self.shapes[keys[0]]
        return self.flat16[off: off + n].view(len(keys) * sh[0], *sh[1:])

    def sync_shadow(self):
        """flat16 <- bf16(flat): after load_state_dict / reset_parameters, or before an eager bf16 forward."""
        from ._lib import lib
        lib().f32_to_bf16(self.flat.data_ptr(), self.flat16.data_ptr(), self.n_total, torch.cuda.current_stream().cuda_stream)

    def fused(self, keys, grad=False):
        """One 2-D/1-D view over parameters that are adjacent in the flat buffer (e.g. key|query|value)."""
        flat = self.flat_grad if grad else self.flat
        off = self.offsets[keys[0]]
        n = 0
        for k in keys:
            assert self.offsets[k] == off + n, "parameters are not adjacent"
            assert _numel(self.shapes[k]) % 4 == 0
            n += _numel(self.shapes[k])
        sh = self.shapes[keys[0]]
        t = flat[off: off + n]
        return t.view(len(keys) * sh[0], *sh[1:])

    def buf(self, key):
        shape, off = dict((k, s) for k, s, kd in self.spec if kd == "buf")[key], self.buf_offsets[key]
        return self.flat_buf[off: off + _numel(shape)].view(shape)

    # torch-facing views (reference shapes) ------------------------------------------------
    def torch_view(self, key, grad=False):
        t = self._native(self.flat_grad if grad else self.flat, key)
        if self.kinds[key] == "conv":
            return t.permute(0, 3, 1, 2)
        return t

    def register(self, root: nn.Module):
        """Create the reference module tree (plain containers) under `root` and register every
        parameter / buffer under its reference name, in reference order."""
        bufshape = {k: s for k, s, kd in self.spec if kd == "buf"}
        for key, shape, kind in self.spec:
            parts = key.split(".")
            mod = root
            for name in parts[:-1]:
                if name not in mod._modules:
                    mod.add_module(name, nn.Module())
                mod = mod._modules[name]
            leaf = parts[-1]
            if kind in ("conv", "f"):
                prm = nn.Parameter(self.torch_view(key), requires_grad=True)
                mod.register_parameter(leaf, prm)
            elif kind == "buf":
                off = self.buf_offsets[key]
                mod.register_buffer(leaf, self.flat_buf[off: off + _numel(bufshape[key])].view(shape))
            else:
                mod.register_buffer(leaf, self.flat_nbt[self.nbt_index[key]])
