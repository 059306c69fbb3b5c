"""Seeded synthetic checkpoints (TEST INFRASTRUCTURE).  No real weights exist offline: every parity and
bench run uses these.  Formats mirror what the reference loads:
  * TrackNet: {'param_dict': {'seq_len': 8, 'bg_mode': 'concat'}, 'model': state_dict}   (ball_tracker.py:253-265)
  * YOLO: {'scale','nc','kpt_shape','model': state_dict with ultralytics key names}  (our stand-in for the pickled
    ultralytics module that `YOLO(model_path)` loads at players_tracker.py:303)
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .tracknet import TrackNetOracle
from .yolov8 import YoloV8

SEEDS = {"detect": 1, "pose13": 2, "court12": 3, "tracknet": 4}


def _init_convs(model: nn.Module, g: torch.Generator, gain: float):
    for m in model.modules():
        if isinstance(m, nn.Conv2d) and m.weight.requires_grad:
            fan_in = m.weight.shape[1] * m.weight.shape[2] * m.weight.shape[3]
            m.weight.data = torch.randn(m.weight.shape, generator=g) * (gain / math.sqrt(fan_in))
            if m.bias is not None:
                m.bias.data = torch.randn(m.bias.shape, generator=g) * 0.05
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.data = 0.9 + 0.2 * torch.rand(m.weight.shape, generator=g)
            m.bias.data = 0.05 * torch.randn(m.bias.shape, generator=g)
            m.running_mean.data = 0.05 * torch.randn(m.running_mean.shape, generator=g)
            m.running_var.data = 0.9 + 0.2 * torch.rand(m.running_var.shape, generator=g)


def _calib_yolo_input() -> torch.Tensor:
    """One synthetic scene (padel_analytics_b200.synth) at 640x360, letterboxed to 384x640 like the detect path."""
    from padel_analytics_b200 import synth

    f = synth.make_frames(1, 360, 640, start=3)[0].flip(-1).float() / 255.0  # RGB
    x = torch.full((1, 3, 384, 640), 114.0 / 255.0)
    x[0, :, 12:372] = f.permute(2, 0, 1)
    return x


def _calib_tracknet_input() -> torch.Tensor:
    from padel_analytics_b200 import synth

    fr = synth.make_frames(8, 288, 512, start=3).flip(-1)  # RGB
    med = synth.make_median(288, 512)
    chans = [med] + [fr[i] for i in range(8)]
    return torch.cat([c.permute(2, 0, 1) for c in chans]).unsqueeze(0).float() / 255.0


@torch.no_grad()
def _standardise(conv: nn.Conv2d, x: torch.Tensor, mean: float, std: float, rows=None):
    """Rescale a final conv so its outputs on `x` have the given mean / std (rows: subset of output channels)."""
    sel = slice(None) if rows is None else rows
    conv.bias.data[sel] = 0.0
    z = conv(x)
    zs = z[:, sel]
    scale = std / float(zs.std().clamp_min(1e-6))
    conv.weight.data[sel] *= scale
    conv.bias.data[sel] = mean - float(zs.mean()) * scale


def make_tracknet(seed: int = SEEDS["tracknet"], frac_above: float = 1e-2) -> dict:
    """Random TrackNet whose heat-maps cross 0.5 on roughly `frac_above` of the pixels (a few blobs per frame)."""
    g = torch.Generator().manual_seed(seed)
    net = TrackNetOracle(27, 8).eval()
    _init_convs(net, g, gain=math.sqrt(2.0))
    x = _calib_tracknet_input()
    with torch.no_grad():
        pool = lambda t: torch.nn.functional.max_pool2d(t, 2, 2)
        up = lambda t: torch.nn.functional.interpolate(t, scale_factor=2, mode="nearest")
        x1 = net.down_block_1(x)
        x2 = net.down_block_2(pool(x1))
        x3 = net.down_block_3(pool(x2))
        y = net.bottleneck(pool(x3))
        y = net.up_block_1(torch.cat([up(y), x3], 1))
        y = net.up_block_2(torch.cat([up(y), x2], 1))
        y = net.up_block_3(torch.cat([up(y), x1], 1))
    # the 8 per-frame heat-map heads share one direction (+10% individual part) so that the temporal ensemble of
    # different windows' channels agrees on where the blobs are, as a trained TrackNet's heads do
    w = net.predictor.weight.data
    w[:] = w[:1] + 0.1 * w
    z = -math.sqrt(2.0) * torch.erfinv(torch.tensor(2 * frac_above - 1.0)).item()  # upper quantile of N(0,1)
    _standardise(net.predictor, y, mean=-z * 1.5, std=1.5)
    return {"param_dict": {"seq_len": 8, "bg_mode": "concat"}, "model": net.state_dict()}


def calib_from_frame(frame_bgr, imgsz: int = 640) -> torch.Tensor:
    """A natural frame (HWC uint8 BGR) as a calibration input: RGB, letterboxed like the detect path."""
    from .yolov8 import letterbox

    lb = letterbox(frame_bgr, imgsz, auto=True)[..., ::-1]
    return torch.from_numpy(lb.transpose(2, 0, 1).copy()).float().unsqueeze(0) / 255.0


def make_yolo(kind: str, scale: str = "n", seed: int | None = None, cls_mean: float | None = None,
              calib: torch.Tensor | None = None) -> dict:
    """kind: 'detect' (nc=80), 'pose13' (nc=1, 13x3 kpts), 'court12' (nc=1, 12x3 kpts).  Last layers are
    standardised on a calibration image so that O(1%) of the anchors exceed the trackers' confidence thresholds
     with varied box sizes; for 'detect' class 0 (person) dominates the other 79."""
    nc, kpt = {"detect": (80, None), "pose13": (1, (13, 3)), "court12": (1, (12, 3))}[kind]
    g = torch.Generator().manual_seed(SEEDS[kind] if seed is None else seed)
    net = YoloV8(scale, nc, kpt).eval()
    _init_convs(net, g, gain=1.6)
    head = net.model[22]
    with torch.no_grad():  # calib: (1,3,H,W) input to standardise on (default: the synthetic scene)
        feats = net.features(_calib_yolo_input() if calib is None else calib)
    # default: dense detections (parity tests want many candidates); bench.py passes a lower cls_mean so that a
    # frame yields a realistic handful of players
    if cls_mean is None:
        cls_mean = {"detect": -2.8, "pose13": -4.2, "court12": -3.4}[kind]
    for l in range(3):
        with torch.no_grad():
            hb = head.cv2[l][1](head.cv2[l][0](feats[l]))
            hc = head.cv3[l][1](head.cv3[l][0](feats[l]))
        _standardise(head.cv2[l][2], hb, mean=0.0, std=2.5)  # peaky DFL distributions -> varied box sizes
        if nc > 1:
            _standardise(head.cv3[l][2], hc, mean=-7.0, std=1.0, rows=slice(1, nc))
            _standardise(head.cv3[l][2], hc, mean=cls_mean, std=1.3, rows=slice(0, 1))
        else:
            _standardise(head.cv3[l][2], hc, mean=cls_mean, std=1.3)
        if kpt:
            with torch.no_grad():
                hk = head.cv4[l][1](head.cv4[l][0](feats[l]))
            _standardise(head.cv4[l][2], hk, mean=0.0, std=1.5)
            head.cv4[l][2].bias.data[2::3] += 1.0  # most keypoints visible (conf > 0.5), some not
    return {"scale": scale, "nc": nc, "kpt_shape": kpt, "model": net.state_dict()}


def load_tracknet(ckpt: dict) -> TrackNetOracle:
    net = TrackNetOracle(27, 8)
    net.load_state_dict(ckpt["model"])
    return net.eval()


def load_yolo(ckpt: dict) -> YoloV8:
    net = YoloV8(ckpt["scale"], ckpt["nc"], tuple(ckpt["kpt_shape"]) if ckpt["kpt_shape"] else None)
    net.load_state_dict(ckpt["model"])
    return net.eval()
