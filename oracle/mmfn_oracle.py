"""CPU fp32 oracle for the MMFN training step (TEST INFRASTRUCTURE, not product code).

A functional torch restatement of the reference hot path, driven by a flat ``state_dict``
with the reference's key names, so the same weights can be fed to the reference module
(in the build container), to this oracle and to the CUDA implementation:

  MMFN.forward           team_code/mmfn_utils/models/model_rad.py:666-695
  Encoder.forward        model_rad.py:492-611
  GPT / RadarGPT.forward model_rad.py:211-247, :962-1000  (Block :127-133, SelfAttention :92-109)
  VectornetEncoder       model_rad.py:369-417 (Subgraph :270-283, MaskSelfAttention :302-325)
  SpGAT                  model_rad.py:800-847, :877-884
  torchvision BasicBlock / ResNet stem (third-party; SURVEY.md section 8c cheat-sheet)
  Engine.train step      run_steps/phase2_train_net.py:60-110 (+ AdamW defaults, :256)

Dropout is the identity here (parity runs use embd/attn/resid pdrop = 0; the reference's RNG
stream cannot be reproduced).  BatchNorm runs in train mode (batch statistics) and updates
the running buffers in place like the reference, or in eval mode with running statistics.
Pinned against the real reference module by tools/make_goldens.py -> tests/golden/*.npz.
"""
import math

import torch
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


class Params:
    """state_dict accessor with a key prefix."""

    def __init__(self, sd, prefix=""):
        self.sd, self.prefix = sd, prefix

    def __call__(self, name):
        return self.sd[self.prefix + name]

    def sub(self, name):
        return Params(self.sd, self.prefix + name + ".")


def _bn(x, p, train, momentum=0.1, eps=1e-5):
    if train:
        p.sd[p.prefix + "num_batches_tracked"] += 1
    return F.batch_norm(x, p("running_mean"), p("running_var"), p("weight"), p("bias"),
                        training=train, momentum=momentum, eps=eps)


def _basic_block(x, p, stride, train):
    out = F.conv2d(x, p("conv1.weight"), stride=stride, padding=1)
    out = F.relu(_bn(out, p.sub("bn1"), train))
    out = F.conv2d(out, p("conv2.weight"), stride=1, padding=1)
    out = _bn(out, p.sub("bn2"), train)
    if (p.prefix + "downsample.0.weight") in p.sd:
        x = _bn(F.conv2d(x, p("downsample.0.weight"), stride=stride), p.sub("downsample.1"), train)
    return F.relu(out + x)


def _res_layer(x, p, nblocks, stride, train):
    for i in range(nblocks):
        x = _basic_block(x, p.sub(str(i)), stride if i == 0 else 1, train)
    return x


def _stem(x, p, train):
    x = F.conv2d(x, p("conv1.weight"), stride=2, padding=3)
    x = F.relu(_bn(x, p.sub("bn1"), train))
    return F.max_pool2d(x, 3, 2, 1)


def _attention(x, p, n_head):
    B, T, C = x.shape
    hs = C // n_head
    k = F.linear(x, p("key.weight"), p("key.bias")).view(B, T, n_head, hs).transpose(1, 2)
    q = F.linear(x, p("query.weight"), p("query.bias")).view(B, T, n_head, hs).transpose(1, 2)
    v = F.linear(x, p("value.weight"), p("value.bias")).view(B, T, n_head, hs).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(hs)), dim=-1)
    y = (att @ v).transpose(1, 2).contiguous().view(B, T, C)
    return F.linear(y, p("proj.weight"), p("proj.bias"))


def _block(x, p, n_head):
    C = x.shape[-1]
    x = x + _attention(F.layer_norm(x, (C,), p("ln1.weight"), p("ln1.bias")), p.sub("attn"), n_head)
    h = F.layer_norm(x, (C,), p("ln2.weight"), p("ln2.bias"))
    h = F.relu(F.linear(h, p("mlp.0.weight"), p("mlp.0.bias")))
    return x + F.linear(h, p("mlp.2.weight"), p("mlp.2.bias"))


def _gpt(feats, velocity, p, cfg):
    """feats: list of (B, C, 8, 8) maps, modality-major token order.  Returns the same list shape."""
    B, C = feats[0].shape[:2]
    tok = torch.cat([f.view(B, 1, C, 8, 8) for f in feats], dim=1).permute(0, 1, 3, 4, 2).reshape(B, -1, C)
    vel = F.linear(velocity.unsqueeze(1), p("vel_emb.weight"), p("vel_emb.bias"))
    x = p("pos_emb") + tok + vel.unsqueeze(1)
    for i in range(cfg.n_layer):
        x = _block(x, p.sub(f"blocks.{i}"), cfg.n_head)
    x = F.layer_norm(x, (C,), p("ln_f.weight"), p("ln_f.bias"))
    x = x.view(B, len(feats), 8, 8, C).permute(0, 1, 4, 2, 3)
    return [x[:, m].contiguous() for m in range(len(feats))]


def _ln_act_mlp(x, p, i_lin, i_ln, act):
    x = F.linear(x, p(f"{i_lin}.weight"), p(f"{i_lin}.bias"))
    x = F.layer_norm(x, (x.shape[-1],), p(f"{i_ln}.weight"), p(f"{i_ln}.bias"))
    return act(x)


def vectornet(lane, lane_num, p):
    """lane (B, L, P, 5), lane_num (B,) -> (B, 64, 64, 64)."""
    B, L = lane.shape[:2]
    x = torch.cat([lane[:, :, :-1, 0:2], lane[:, :, 1:, 0:2], lane[:, :, 1:, 2:]], dim=-1).float()
    for i in range(3):
        x = _ln_act_mlp(x, p.sub(f"lane_subgraph.layers.mlp_{i}.mlp"), 0, 1, F.relu)
        mx = x.max(dim=-2, keepdim=True)[0].expand_as(x)
        x = torch.cat([x, mx], dim=-1)
    tok = x.max(dim=-2)[0]                                          # (B, L, 128)
    qkv = F.linear(tok, p("L2L.to_qkv.weight")).view(B, L, 3, 2, 64)
    q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))      # (B, 2, L, 64)
    dots = (q @ k.transpose(-1, -2)) * (64 ** -0.5)
    valid = torch.arange(L)[None, :] < lane_num.to(torch.int64)[:, None]
    dots = dots.masked_fill(~valid[:, None, None, :], -1e9)
    att = (torch.softmax(dots, dim=-1) @ v).transpose(1, 2).reshape(B, L, 128)
    tok = F.linear(att, p("L2L.to_out.0.weight"), p("L2L.to_out.0.bias"))
    pos = _ln_act_mlp(torch.zeros(B, L, 2), p.sub("pos_emb"), 0, 1, F.gelu)
    pos = F.linear(pos, p("pos_emb.3.weight"), p("pos_emb.3.bias"))
    fuse = _ln_act_mlp(torch.cat([tok, pos], dim=-1), p.sub("agent_fusion"), 0, 1, F.gelu)
    fuse = F.linear(fuse, p("agent_fusion.3.weight"), p("agent_fusion.3.bias"))
    g = _ln_act_mlp(fuse[:, 0], p.sub("generator"), 0, 1, F.gelu)
    g = F.linear(g, p("generator.3.weight"), p("generator.3.bias"))
    return g.view(B, 64, 64, 64)


def spgat(radar, adj, p, alpha, nheads):
    """radar (B, 81, 5), adj (B, 81, 81) -> (B, 512, 8, 8)."""
    heads = []
    for i in range(nheads):
        Wh = radar @ p(f"attention_{i}.W")
        e = F.leaky_relu(Wh @ p(f"attention_{i}.a"), alpha)
        att = torch.softmax(torch.where(adj > 0, e, torch.full_like(e, -9e15)), dim=-1)
        heads.append(F.elu(att @ Wh))
    x = F.elu(torch.cat(heads, dim=1))
    x = F.linear(x, p("mlp_1.0.weight"), p("mlp_1.0.bias"))
    x = F.linear(x.transpose(1, 2), p("mlp_2.0.weight"), p("mlp_2.0.bias"))
    x = x.reshape(x.shape[0], 8, 8, 512).transpose(1, 3)
    return F.log_softmax(x, dim=1)


def encoder(sd, cfg, image, lidar, lane, lane_num, radar, radar_adj, velocity, train=True, taps=None, variant="rad"):
    """variant "rad": model_rad.py:492-611; "vec": model_vec.py:488-600 (no radar branch, transformer4 is a plain
    3-modality GPT); "img": model_img.py:310-423 (`lane` is the rasterised map image (B,3,256,256), fed
    un-normalised through the map ResNet's stem and layer1; no VectorNet, no radar)."""
    p = Params(sd, "encoder.")
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    img_p, map_p, lid_p = p.sub("image_encoder.features"), p.sub("img_map_encoder.features"), p.sub("lidar_encoder._model")
    img = _stem((image - mean) / std, img_p, train)
    lid = _stem(lidar, lid_p, train)
    img = _res_layer(img, img_p.sub("layer1"), 3, 1, train)
    lid = _res_layer(lid, lid_p.sub("layer1"), 2, 1, train)
    if variant == "img":
        mp = _res_layer(_stem(lane, map_p, train), map_p.sub("layer1"), 3, 1, train)
    else:
        mp = vectornet(lane, lane_num, p.sub("vectornet_encoder"))
    if taps is not None:
        taps.update(img_l1=img, lid_l1=lid, map_gen=mp)
    pool = lambda t: F.adaptive_avg_pool2d(t, (8, 8))
    blocks = {"img": (4, 6, 3), "lid": (2, 2, 2)}
    for s, scale in ((1, 8), (2, 4), (3, 2)):
        outs = _gpt([pool(img), pool(lid), pool(mp)], velocity, p.sub(f"transformer{s}"), cfg)
        up = lambda t: F.interpolate(t, scale_factor=scale, mode="bilinear", align_corners=True)
        img, lid, mp = img + up(outs[0]), lid + up(outs[1]), mp + up(outs[2])
        if taps is not None:
            taps[f"img_f{s}"] = img
        img = _res_layer(img, img_p.sub(f"layer{s + 1}"), blocks["img"][s - 1], 2, train)
        mp = _res_layer(mp, map_p.sub(f"layer{s + 1}"), blocks["img"][s - 1], 2, train)
        lid = _res_layer(lid, lid_p.sub(f"layer{s + 1}"), blocks["lid"][s - 1], 2, train)
    if variant != "rad":
        outs = _gpt([pool(img), pool(lid), pool(mp)], velocity, p.sub("transformer4"), cfg)
        return sum((f + o).mean(dim=(2, 3)) for f, o in zip((img, lid, mp), outs))
    rad = spgat(radar, radar_adj, p.sub("radar_encoder"), cfg.alpha, cfg.nb_heads)
    outs = _gpt([pool(img), pool(lid), pool(mp), rad], velocity, p.sub("transformer4"), cfg)
    feats = [img + outs[0], lid + outs[1], mp + outs[2], rad + outs[3]]
    if taps is not None:
        taps.update(rad=rad, img_l4=feats[0])
    return sum(f.mean(dim=(2, 3)) for f in feats)


def forward(sd, cfg, image, lidar, lane, lane_num, radar, radar_adj, target_point, velocity, train=True, taps=None,
            variant="rad"):
    """-> pred_wp (B, pred_len, 2).  image (B,3,256,256) 0..255, lidar (B,2,256,256)."""
    fused = encoder(sd, cfg, image, lidar, lane, lane_num, radar, radar_adj, velocity, train, taps, variant)
    return head(sd, cfg, fused, target_point)


def head(sd, cfg, fused, target_point):
    """join MLP + GRUCell waypoint roll-out (model_rad.py:676-695; identical in benchmarks/transfuser/model.py:440-458)."""
    z = fused
    for i in (0, 2, 4):
        z = F.relu(F.linear(z, sd[f"join.{i}.weight"], sd[f"join.{i}.bias"]))
    x = torch.zeros(z.shape[0], 2)
    wps = []
    for _ in range(cfg.pred_len):
        xin = x + target_point
        gi = F.linear(xin, sd["decoder.weight_ih"], sd["decoder.bias_ih"])
        gh = F.linear(z, sd["decoder.weight_hh"], sd["decoder.bias_hh"])
        i_r, i_z, i_n = gi.chunk(3, 1)
        h_r, h_z, h_n = gh.chunk(3, 1)
        r, zg = torch.sigmoid(i_r + h_r), torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        z = (1 - zg) * n + zg * z
        x = x + F.linear(z, sd["output.weight"], sd["output.bias"])
        wps.append(x)
    return torch.stack(wps, dim=1)


def l1_loss(pred, gt):
    return (pred - gt).abs().mean()


def is_float_param(k, v):
    return v.dtype.is_floating_point and not (k.endswith("running_mean") or k.endswith("running_var"))


def train_step(sd, cfg, batch, opt_state=None, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, wd=0.01, forward_fn=None):
    """One Engine.train iteration on CPU.  Mutates sd (params + BN buffers) in place.
    Returns (loss, pred_wp, grads dict).  opt_state: {'t': int, 'm': {}, 'v': {}} or None (no update)."""
    names = [k for k, v in sd.items() if is_float_param(k, v)]
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
    work = dict(sd)
    work.update(leaves)
    pred = (forward_fn or forward)(work, cfg, *batch["inputs"], train=True)
    loss = l1_loss(pred, batch["gt_waypoints"])
    loss.backward()
    for k in sd:                      # BN buffers were updated inside `work`
        if k not in leaves:
            sd[k] = work[k]
    grads = {k: leaves[k].grad for k in names}
    if opt_state is not None:
        opt_state["t"] += 1
        t = opt_state["t"]
        for k in names:
            g = grads[k]
            if g is None:             # torch.optim skips parameters without a gradient entirely
                continue
            m = opt_state["m"].setdefault(k, torch.zeros_like(g))
            v = opt_state["v"].setdefault(k, torch.zeros_like(g))
            with torch.no_grad():
                pk = sd[k]
                pk.mul_(1 - lr * wd)
                m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
                v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
                denom = (v.sqrt() / math.sqrt(1 - betas[1] ** t)).add_(eps)
                pk.addcdiv_(m, denom, value=-lr / (1 - betas[0] ** t))
    return loss.detach(), pred.detach(), grads
