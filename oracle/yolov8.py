"""Plain-PyTorch restatement of ultralytics YOLOv8 detect / pose and its predict() pipeline (CPU oracle).

ultralytics is a third-party dependency of the reference (requirements.txt:9, unpinned; 8.3.x era) that is absent
from /root/reference and from this image.  This file restates its published YOLOv8 architecture and inference
pipeline and plugs in where the reference calls it:
    /root/reference/trackers/players_tracker/players_tracker.py:303,351-359
    /root/reference/trackers/players_keypoints_tracker/players_keypoints_tracker.py:238,285-292
    /root/reference/trackers/keypoints_tracker/keypoints_tracker.py:169,238-245
State-dict key names follow ultralytics (`model.{i}.conv.weight`, `model.22.cv2.{l}.2.bias`, ...) so real
checkpoints' state dicts load.  PARITY UNPINNED for the network arithmetic (no runnable ultralytics here); the
pre/post-processing uses the very same third-party calls (cv2.resize, copyMakeBorder, torchvision.ops.nms).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import cv2
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision

SCALES = {  # depth, width, max_channels
    "n": (0.33, 0.25, 1024), "s": (0.33, 0.50, 1024), "m": (0.67, 0.75, 768),
    "l": (1.00, 1.00, 512), "x": (1.00, 1.25, 512),
}


def _ch(c, width, max_ch):
    return int(math.ceil(min(c, max_ch) * width / 8) * 8)


def _rep(n, depth):
    return max(round(n * depth), 1) if n > 1 else n


class ConvBnAct(nn.Module):
    """Conv2d(bias=False, pad=k//2) + BatchNorm2d(eps=1e-3) + SiLU."""

    def __init__(self, c1, c2, k=1, s=1):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, k // 2, bias=False)
        self.bn = nn.BatchNorm2d(c2, eps=1e-3, momentum=0.03)

    def forward(self, x):
        return F.silu(self.bn(self.conv(x)))


class Bottleneck(nn.Module):
    def __init__(self, c, shortcut):
        super().__init__()
        self.cv1 = ConvBnAct(c, c, 3)
        self.cv2 = ConvBnAct(c, c, 3)
        self.add = shortcut

    def forward(self, x):
        y = self.cv2(self.cv1(x))
        return x + y if self.add else y


class C2f(nn.Module):
    def __init__(self, c1, c2, n, shortcut):
        super().__init__()
        self.c = c2 // 2
        self.cv1 = ConvBnAct(c1, 2 * self.c, 1)
        self.cv2 = ConvBnAct((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(Bottleneck(self.c, shortcut) for _ in range(n))

    def forward(self, x):
        y = list(self.cv1(x).chunk(2, 1))
        for m in self.m:
            y.append(m(y[-1]))
        return self.cv2(torch.cat(y, 1))


class SPPF(nn.Module):
    def __init__(self, c1, c2, k=5):
        super().__init__()
        self.cv1 = ConvBnAct(c1, c1 // 2, 1)
        self.cv2 = ConvBnAct(c1 // 2 * 4, c2, 1)
        self.k = k

    def forward(self, x):
        y = [self.cv1(x)]
        for _ in range(3):
            y.append(F.max_pool2d(y[-1], self.k, 1, self.k // 2))
        return self.cv2(torch.cat(y, 1))


class _Dfl(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(16, 1, 1, bias=False).requires_grad_(False)
        self.conv.weight.data[:] = torch.arange(16, dtype=torch.float).view(1, 16, 1, 1)


def _branch(cin, cmid, cout):
    return nn.Sequential(ConvBnAct(cin, cmid, 3), ConvBnAct(cmid, cmid, 3), nn.Conv2d(cmid, cout, 1))


class DetectHead(nn.Module):
    reg_max = 16

    def __init__(self, nc, ch):
        super().__init__()
        self.nc = nc
        cb = max(16, ch[0] // 4, 64)
        cc = max(ch[0], min(nc, 100))
        self.cv2 = nn.ModuleList(_branch(c, cb, 64) for c in ch)
        self.cv3 = nn.ModuleList(_branch(c, cc, nc) for c in ch)
        self.dfl = _Dfl()
        self.strides = (8, 16, 32)

    def raw(self, feats):
        """Per level (B, 64+nc[+nk], h, w) raw head maps."""
        return [torch.cat([self.cv2[i](f), self.cv3[i](f)], 1) for i, f in enumerate(feats)]

    @staticmethod
    def anchors(shapes, strides):
        pts, st = [], []
        for (h, w), s in zip(shapes, strides):
            sx = torch.arange(w, dtype=torch.float32) + 0.5
            sy = torch.arange(h, dtype=torch.float32) + 0.5
            gy, gx = torch.meshgrid(sy, sx, indexing="ij")
            pts.append(torch.stack((gx, gy), -1).view(-1, 2))
            st.append(torch.full((h * w, 1), float(s)))
        return torch.cat(pts).T, torch.cat(st).T  # (2,A), (1,A)

    def decode_boxes(self, raws):
        B = raws[0].shape[0]
        x = torch.cat([r[:, : 64 + self.nc].reshape(B, 64 + self.nc, -1) for r in raws], 2)
        box, cls = x.split((64, self.nc), 1)
        anc, st = self.anchors([r.shape[2:] for r in raws], self.strides)
        anc, st = anc.to(box.device), st.to(box.device)
        b, _, a = box.shape
        dist = box.view(b, 4, 16, a).transpose(2, 1).softmax(1)
        dist = (dist * torch.arange(16, dtype=torch.float32, device=box.device).view(1, 16, 1, 1)).sum(1)  # DFL expectation
        lt, rb = dist.chunk(2, 1)
        x1y1, x2y2 = anc.unsqueeze(0) - lt, anc.unsqueeze(0) + rb
        dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * st
        return torch.cat((dbox, cls.sigmoid()), 1), anc, st

    def forward(self, feats):
        y, _, _ = self.decode_boxes(self.raw(feats))
        return y


class PoseHead(DetectHead):
    def __init__(self, nc, kpt_shape, ch):
        super().__init__(nc, ch)
        self.kpt_shape = kpt_shape
        self.nk = kpt_shape[0] * kpt_shape[1]
        ck = max(ch[0] // 4, self.nk)
        self.cv4 = nn.ModuleList(_branch(c, ck, self.nk) for c in ch)

    def raw(self, feats):
        return [torch.cat([self.cv2[i](f), self.cv3[i](f), self.cv4[i](f)], 1) for i, f in enumerate(feats)]

    def forward(self, feats):
        raws = self.raw(feats)
        y, anc, st = self.decode_boxes(raws)
        B = raws[0].shape[0]
        kpt = torch.cat([r[:, 64 + self.nc:].reshape(B, self.nk, -1) for r in raws], 2)
        K, D = self.kpt_shape
        k = kpt.view(B, K, D, -1).clone()
        k[:, :, 0] = (k[:, :, 0] * 2.0 + (anc[0] - 0.5)) * st
        k[:, :, 1] = (k[:, :, 1] * 2.0 + (anc[1] - 0.5)) * st
        if D == 3:
            k[:, :, 2] = k[:, :, 2].sigmoid()
        return torch.cat([y, k.view(B, self.nk, -1)], 1)


class YoloV8(nn.Module):
    """Layers 0..22 of yolov8{,-pose}.yaml. `model` is a ModuleList so keys read `model.{i}....`."""

    def __init__(self, scale="n", nc=80, kpt_shape=None):
        super().__init__()
        d, w, mc = SCALES[scale]
        c = [_ch(v, w, mc) for v in (64, 128, 256, 512, 1024)]
        r3, r6 = _rep(3, d), _rep(6, d)
        up = nn.Upsample(scale_factor=2, mode="nearest")
        layers = [
            ConvBnAct(3, c[0], 3, 2), ConvBnAct(c[0], c[1], 3, 2), C2f(c[1], c[1], r3, True),
            ConvBnAct(c[1], c[2], 3, 2), C2f(c[2], c[2], r6, True),
            ConvBnAct(c[2], c[3], 3, 2), C2f(c[3], c[3], r6, True),
            ConvBnAct(c[3], c[4], 3, 2), C2f(c[4], c[4], r3, True), SPPF(c[4], c[4]),
            up, nn.Identity(), C2f(c[4] + c[3], c[3], r3, False),
            up, nn.Identity(), C2f(c[3] + c[2], c[2], r3, False),
            ConvBnAct(c[2], c[2], 3, 2), nn.Identity(), C2f(c[2] + c[3], c[3], r3, False),
            ConvBnAct(c[3], c[3], 3, 2), nn.Identity(), C2f(c[3] + c[4], c[4], r3, False),
        ]
        head = PoseHead(nc, kpt_shape, c[2:]) if kpt_shape else DetectHead(nc, c[2:])
        self.model = nn.ModuleList(layers + [head])
        self.scale, self.nc, self.kpt_shape = scale, nc, kpt_shape
        self.names = {i: f"class{i}" for i in range(nc)}
        if nc == 80:
            self.names[0] = "person"

    def features(self, x):
        m = self.model
        x = m[1](m[0](x))
        x = m[2](x)
        p3 = m[4](m[3](x))
        p4 = m[6](m[5](p3))
        p5 = m[9](m[8](m[7](p4)))
        h4 = m[12](torch.cat([m[10](p5), p4], 1))
        o3 = m[15](torch.cat([m[13](h4), p3], 1))
        o4 = m[18](torch.cat([m[16](o3), h4], 1))
        o5 = m[21](torch.cat([m[19](o4), p5], 1))
        return [o3, o4, o5]

    def forward(self, x):
        return self.model[22](self.features(x))

    def raw_heads(self, x):
        return self.model[22].raw(self.features(x))


# ------------------------------------------------------------------------------------------------------------
# predict() pipeline
# ------------------------------------------------------------------------------------------------------------
def letterbox(im: np.ndarray, imgsz: int, auto: bool, stride: int = 32):
    h, w = im.shape[:2]
    r = min(imgsz / h, imgsz / w)
    new_unpad = int(round(w * r)), int(round(h * r))
    dw, dh = imgsz - new_unpad[0], imgsz - new_unpad[1]
    if auto:
        dw, dh = np.mod(dw, stride), np.mod(dh, stride)
    dw /= 2
    dh /= 2
    if (w, h) != new_unpad:
        im = cv2.resize(im, new_unpad, interpolation=cv2.INTER_LINEAR)
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return cv2.copyMakeBorder(im, top, bottom, left, right, cv2.BORDER_CONSTANT, value=(114, 114, 114))


def xywh2xyxy(x):
    y = torch.empty_like(x)
    xy, wh = x[..., :2], x[..., 2:] / 2
    y[..., :2] = xy - wh
    y[..., 2:] = xy + wh
    return y


def non_max_suppression(pred, conf_thres, iou_thres, classes, max_det, nc, max_nms=30000, max_wh=7680):
    bs = pred.shape[0]
    mi = 4 + nc
    xc = pred[:, 4:mi].amax(1) > conf_thres
    pred = pred.transpose(-1, -2).clone()
    pred[..., :4] = xywh2xyxy(pred[..., :4])
    out = [torch.zeros((0, 6 + pred.shape[-1] - mi), device=pred.device)] * bs
    for xi, x in enumerate(pred):
        x = x[xc[xi]]
        if not x.shape[0]:
            continue
        box, cls, mask = x.split((4, nc, x.shape[1] - mi), 1)
        conf, j = cls.max(1, keepdim=True)
        x = torch.cat((box, conf, j.float(), mask), 1)[conf.view(-1) > conf_thres]
        if classes is not None:
            x = x[(x[:, 5:6] == torch.tensor(classes, dtype=x.dtype, device=x.device)).any(1)]
        n = x.shape[0]
        if not n:
            continue
        if n > max_nms:
            x = x[x[:, 4].argsort(descending=True)[:max_nms]]
        c = x[:, 5:6] * max_wh
        i = torchvision.ops.nms(x[:, :4] + c, x[:, 4], iou_thres)[:max_det]
        out[xi] = x[i]
    return out


def scale_boxes(img1_shape, boxes, img0_shape):
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad = (round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1),
           round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1))
    boxes = boxes.clone()
    boxes[..., [0, 2]] -= pad[0]
    boxes[..., [1, 3]] -= pad[1]
    boxes[..., :4] /= gain
    boxes[..., [0, 2]] = boxes[..., [0, 2]].clamp(0, img0_shape[1])
    boxes[..., [1, 3]] = boxes[..., [1, 3]].clamp(0, img0_shape[0])
    return boxes


def scale_coords(img1_shape, coords, img0_shape):
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad = (img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2
    coords = coords.clone()
    coords[..., 0] -= pad[0]
    coords[..., 1] -= pad[1]
    coords[..., 0] /= gain
    coords[..., 1] /= gain
    coords[..., 0] = coords[..., 0].clamp(0, img0_shape[1])
    coords[..., 1] = coords[..., 1].clamp(0, img0_shape[0])
    return coords


@dataclass
class Boxes:
    data: torch.Tensor  # (N,6) xyxy, conf, cls

    @property
    def xyxy(self):
        return self.data[:, :4]

    @property
    def conf(self):
        return self.data[:, 4]

    @property
    def cls(self):
        return self.data[:, 5]

    @property
    def id(self):
        return None

    def __len__(self):
        return self.data.shape[0]


@dataclass
class Keypoints:
    data: torch.Tensor  # (N,K,D)

    @property
    def xy(self):
        return self.data[..., :2]

    @property
    def conf(self):
        return self.data[..., 2] if self.data.shape[-1] == 3 else None


@dataclass
class Result:
    boxes: Boxes
    keypoints: Keypoints | None
    names: dict
    orig_shape: tuple


class YOLO:
    """Stand-in for `ultralytics.YOLO` with the surface the reference trackers use (predict / to / names)."""

    def __init__(self, model: YoloV8 | str):
        if isinstance(model, (str, bytes)) or hasattr(model, "__fspath__"):
            ck = torch.load(model, map_location="cpu", weights_only=False)
            net = YoloV8(ck["scale"], ck["nc"], tuple(ck["kpt_shape"]) if ck.get("kpt_shape") else None)
            net.load_state_dict(ck["model"])
            model = net
        self.net = model.eval()
        self.names = model.names
        self.last_preprocessed = None

    def to(self, device):
        return self

    @torch.no_grad()
    def predict(self, source, conf=0.25, iou=0.7, imgsz=640, device=None, classes=None, max_det=300, **kw):
        ims = []
        for s in source:
            if isinstance(s, np.ndarray):
                ims.append(s)  # assumed BGR
            else:
                ims.append(np.asarray(s)[:, :, ::-1])  # PIL RGB -> BGR
        same = len({im.shape for im in ims}) == 1
        lb = [letterbox(np.ascontiguousarray(im), imgsz, auto=same) for im in ims]
        x = np.stack(lb)[..., ::-1].transpose(0, 3, 1, 2)
        x = torch.from_numpy(np.ascontiguousarray(x)).float() / 255
        self.last_preprocessed = x
        pred = self.net(x)
        nc = self.net.nc
        dets = non_max_suppression(pred, conf, iou, classes, max_det, nc)
        results = []
        for det, im in zip(dets, ims):
            det = det.clone()
            det[:, :4] = scale_boxes(x.shape[2:], det[:, :4], im.shape)
            kp = None
            if self.net.kpt_shape:
                K, D = self.net.kpt_shape
                k = det[:, 6:].view(-1, K, D) if len(det) else det[:, 6:].view(0, K, D)
                k = scale_coords(x.shape[2:], k, im.shape)
                if D == 3:  # ultralytics Keypoints: points with conf < 0.5 are zeroed
                    mask = k[..., 2] < 0.5
                    k[..., :2][mask] = 0
                kp = Keypoints(k)
            results.append(Result(Boxes(det[:, :6]), kp, self.names, im.shape[:2]))
        return results
