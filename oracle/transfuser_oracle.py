"""CPU fp32 oracle for the RGB+LiDAR-only variant (TEST INFRASTRUCTURE, not product code).

Functional restatement of team_code/benchmarks/transfuser/model.py, reusing the building blocks of
mmfn_oracle (same torchvision ResNets, same Block / SelfAttention):

  Encoder.forward     model.py:305-387  two modalities, 2 x 64 tokens per GPT (:147, :216-249),
                      F.interpolate(scale_factor, mode='bilinear') with the DEFAULT align_corners=False (:338-339,
                      :351-352, :364-365), no up-sampling at the last stage (:374-376), global pools summed (:378-387)
  TransFuser.forward  model.py:429-458  (join MLP, GRUCell roll-out -- identical to MMFN's head)

Pinned against the real reference module by tools/make_goldens.py -> tests/golden/transfuser_golden_b2.npz.
"""
import torch
import torch.nn.functional as F

from . import mmfn_oracle as mo


def encoder(sd, cfg, image, lidar, velocity, train=True):
    p = mo.Params(sd, "encoder.")
    mean = torch.tensor(mo.IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(mo.IMAGENET_STD).view(1, 3, 1, 1)
    img_p, lid_p = p.sub("image_encoder.features"), p.sub("lidar_encoder._model")
    img = mo._stem((image - mean) / std, img_p, train)
    lid = mo._stem(lidar, lid_p, train)
    img = mo._res_layer(img, img_p.sub("layer1"), 3, 1, train)
    lid = mo._res_layer(lid, lid_p.sub("layer1"), 2, 1, train)
    pool = lambda t: F.adaptive_avg_pool2d(t, (8, 8))
    blocks = {"img": (4, 6, 3), "lid": (2, 2, 2)}
    for s, scale in ((1, 8), (2, 4), (3, 2)):
        outs = mo._gpt([pool(img), pool(lid)], velocity, p.sub(f"transformer{s}"), cfg)
        up = lambda t: F.interpolate(t, scale_factor=scale, mode="bilinear")
        img, lid = img + up(outs[0]), lid + up(outs[1])
        img = mo._res_layer(img, img_p.sub(f"layer{s + 1}"), blocks["img"][s - 1], 2, train)
        lid = mo._res_layer(lid, lid_p.sub(f"layer{s + 1}"), blocks["lid"][s - 1], 2, train)
    outs = mo._gpt([pool(img), pool(lid)], velocity, p.sub("transformer4"), cfg)
    return (img + outs[0]).mean(dim=(2, 3)) + (lid + outs[1]).mean(dim=(2, 3))


def forward(sd, cfg, image, lidar, target_point, velocity, train=True):
    """-> pred_wp (B, pred_len, 2).  image (B,3,256,256) 0..255, lidar (B,2,256,256)."""
    return mo.head(sd, cfg, encoder(sd, cfg, image, lidar, velocity, train), target_point)


def train_step(sd, cfg, batch, opt_state=None, **kw):
    """One optimisation step (benchmarks/transfuser/train.py uses the same L1 + AdamW recipe)."""
    return mo.train_step(sd, cfg, batch, opt_state, forward_fn=forward, **kw)
