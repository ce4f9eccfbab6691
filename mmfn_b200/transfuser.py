"""B200-native TransFuser (RGB + LiDAR-BEV only): the topology of BASELINE.json configs[3] ("RGB+LiDAR
only, map/radar branches off").  Drop-in for team_code/benchmarks/transfuser/model.py:TransFuser (:393-458):

    TransFuser(config, device)
    forward(image_list, lidar_list, target_point, velocity) -> (B, pred_len, 2)
    control_pid(waypoints, velocity)

Same state_dict keys as the reference.  It is the MMFN schedule (model_rad._Net) with two modalities:
every fusion transformer sees 2 x 64 tokens (model.py:147), the token maps are up-sampled with
F.interpolate's DEFAULT align_corners=False (model.py:338-339 -- MMFN uses align_corners=True), and the
final feature is the sum of the two globally pooled trunks (model.py:375-385).
"""
import torch

from . import ops
from .model_rad import MMFN, FusionGPT, Head, ResLayer, Stem, _Aux, _Net, _WholeNet, IMAGENET_MEAN, IMAGENET_STD
from .params import RESNET18, RESNET34, WIDTHS


class _NetTF(_Net):
    """Kernel schedule of transfuser Encoder.forward (model.py:305-387) + the GRU head."""

    def __init__(self, st, cfg):          # noqa: super().__init__ builds the 4-modality network; not called
        e = "encoder."
        self.cfg = cfg
        ip, lp = e + "image_encoder.features", e + "lidar_encoder._model"
        self.img_stem, self.lid_stem = Stem(st, ip), Stem(st, lp)
        self.img_layers = [ResLayer(st, f"{ip}.layer{i + 1}", RESNET34[i], 1 if i == 0 else 2) for i in range(4)]
        self.lid_layers = [ResLayer(st, f"{lp}.layer{i + 1}", RESNET18[i], 1 if i == 0 else 2) for i in range(4)]
        self.gpts = [FusionGPT(st, f"{e}transformer{i + 1}", WIDTHS[i], 2, cfg, i) for i in range(4)]
        self.head = Head(st, cfg.pred_len)
        dev = st.device
        self.side = [torch.cuda.Stream(device=dev, priority=-1) for _ in range(3)]
        self.use_streams = True
        self.mean = torch.tensor(IMAGENET_MEAN, device=dev, dtype=torch.float32)
        self.std = torch.tensor(IMAGENET_STD, device=dev, dtype=torch.float32)

    def forward(self, image, lidar, lane, lane_num, radar, radar_adj, target_point, velocity, seed, train):
        img, lid = self._parallel(
            lambda: self.img_layers[0].fwd(self.img_stem.fwd(ops.nchw_to_nhwc(image, self.mean, self.std), train), train),
            lambda: self.lid_layers[0].fwd(self.lid_stem.fwd(ops.nchw_to_nhwc(lidar() if callable(lidar) else lidar), train), train))
        self._run_idle_hook()
        for s in range(3):
            tok = self.gpts[s].fwd([img, lid], velocity, seed, train)
            img, lid = self._parallel(
                lambda: self.img_layers[s + 1].fwd(ops.upsample_add_fwd(img, tok, 0, align_corners=False), train),
                lambda: self.lid_layers[s + 1].fwd(ops.upsample_add_fwd(lid, tok, 1, align_corners=False), train))
        tok = self.gpts[3].fwd([img, lid], velocity, seed, train)
        fused = ops.pool_sum_fwd([img, lid], tok)
        return self.head.fwd(fused, target_point)

    def backward(self, dpred):
        if self._idle_done is not None:
            torch.cuda.current_stream().wait_event(self._idle_done)
            self._idle_done = None
        dfused = self.head.bwd(dpred)
        dfe, dtok = ops.pool_sum_bwd(dfused, 2)
        self.gpts[3].bwd(dtok, dfe)
        dimg, dlid = dfe
        for s in (2, 1, 0):
            B, C = dimg.shape[0], dimg.shape[3] // 2
            dtok = torch.empty((B, 128, C), device=dimg.device, dtype=torch.float32)

            def trunk(layers, d, m):
                def run():
                    g = layers[s + 1].bwd(d)
                    ops.upsample_add_bwd_(g, dtok, m, align_corners=False)
                    return g
                return run
            dimg, dlid = self._parallel(trunk(self.img_layers, dimg, 0), trunk(self.lid_layers, dlid, 1))
            if s == 2:
                self._early_bucket_done()
            self.gpts[s].bwd(dtok, [dimg, dlid])
            if s == 1:
                self._bucket_done("mid")
        self._parallel(lambda: self.img_stem.bwd(self.img_layers[0].bwd(dimg)),
                       lambda: self.lid_stem.bwd(self.lid_layers[0].bwd(dlid)))
        _Aux.join_all()


class TransFuser(MMFN):
    VARIANT = "transfuser"
    NET = _NetTF

    def forward(self, image_list, lidar_list, target_point, velocity):
        inputs = (image_list[0], lidar_list[0], None, None, None, None, target_point, velocity)
        if torch.is_grad_enabled() and self.training:
            return _WholeNet.apply(self, inputs, *[p for _, p in self._param_items])
        with torch.no_grad():
            return self._forward_impl(*inputs)

    def _forward_impl(self, image, lidar, lane, lane_num, radar, radar_adj, target_point, velocity):
        f32 = lambda t: t.to(device=self.device, dtype=torch.float32).contiguous()
        image = image.to(self.device).contiguous() if image.dtype == torch.uint8 else f32(image)
        lidar, target_point, velocity = f32(lidar), f32(target_point), f32(velocity)
        if self.training:
            self.seed += 1000
            self.store.flat_nbt.add_(self._nbt_step())
        if ops.BF16:
            self.store.sync_shadow()
        return self.net.forward(image, lidar, None, None, None, None, target_point, velocity, self.seed, self.training)
