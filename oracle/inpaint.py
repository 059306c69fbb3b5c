"""CPU oracle for the InpaintNet stage of BallTracker.predict_frames (TEST INFRASTRUCTURE ONLY).

Follows:
  * InpaintNetOracle        <- /root/reference/trackers/ball_tracker/models.py:77-130 (Conv1DBlock / Double1DConv / InpaintNet)
  * generate_inpaint_mask   <- /root/reference/trackers/ball_tracker/ball_tracker.py:100-136
  * make_sequences          <- /root/reference/trackers/ball_tracker/dataset.py:387-429 (_gen_input_from_pred_dict,
                               sliding_step=1, no padding) + :493-503 (__getitem__ coordinate branch: /w, /h)
  * inpaint_stage           <- /root/reference/trackers/ball_tracker/ball_tracker.py:525-673 (blend, COOR_TH threshold,
                               temporal ensemble on coordinates, second threshold) + predict.py:91-147 (predict)
Pinned by tests/golden/inpaint_ref.npz, produced by the reference's own code (tests/golden/make_golden.py; the
reference's hard-coded .cuda() calls are redirected to the CPU there).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from .tracknet import ensemble_weight


class _C1(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, 3, padding="same", bias=True)

    def forward(self, x):
        return nn.functional.leaky_relu(self.conv(x))


class _D1(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv_1 = _C1(cin, cout)
        self.conv_2 = _C1(cout, cout)

    def forward(self, x):
        return self.conv_2(self.conv_1(x))


class InpaintNetOracle(nn.Module):
    def __init__(self):
        super().__init__()
        self.down_1, self.down_2, self.down_3 = _C1(3, 32), _C1(32, 64), _C1(64, 128)
        self.buttleneck = _D1(128, 256)  # (sic) attribute name as in the reference checkpoint
        self.up_1, self.up_2, self.up_3 = _C1(384, 128), _C1(192, 64), _C1(96, 32)
        self.predictor = nn.Conv1d(32, 2, 3, padding="same")

    def forward(self, x, m):
        x = torch.cat([x, m], 2).permute(0, 2, 1)
        x1 = self.down_1(x)
        x2 = self.down_2(x1)
        x3 = self.down_3(x2)
        x = self.buttleneck(x3)
        x = self.up_1(torch.cat([x, x3], 1))
        x = self.up_2(torch.cat([x, x2], 1))
        x = self.up_3(torch.cat([x, x1], 1))
        return torch.sigmoid(self.predictor(x)).permute(0, 2, 1)


def make_inpaintnet(seed: int = 5, seq_len: int = 16) -> dict:
    """Seeded checkpoint in the reference's format {'param_dict': {'seq_len'}, 'model': state_dict}
    (ball_tracker.py:268-272)."""
    g = torch.Generator().manual_seed(seed)
    net = InpaintNetOracle()
    for m in net.modules():
        if isinstance(m, nn.Conv1d):
            fan = m.weight.shape[1] * 3
            m.weight.data = torch.randn(m.weight.shape, generator=g) * (1.4 / math.sqrt(fan))
            m.bias.data = torch.randn(m.bias.shape, generator=g) * 0.05
    return {"param_dict": {"seq_len": seq_len}, "model": net.state_dict()}


def load_inpaintnet(ckpt: dict) -> InpaintNetOracle:
    net = InpaintNetOracle()
    net.load_state_dict(ckpt["model"])
    return net.eval()


def generate_inpaint_mask(y, vis, th_h: float = 30):
    y = np.array(y)
    vis = np.array(vis)
    mask = np.zeros_like(y)
    i = j = 0
    while j < len(vis):
        while i < len(vis) - 1 and vis[i] == 1:
            i += 1
        j = i
        while j < len(vis) - 1 and vis[j] == 0:
            j += 1
        if j == i:
            break
        elif i == 0 and y[j] > th_h:
            mask[:j] = 1
        elif (i > 1 and y[i - 1] > th_h) and (j < len(vis) and y[j] > th_h):
            mask[i:j] = 1
        i = j
    return mask.tolist()


def make_sequences(x, y, mask, seq_len: int, img_wh):
    """(S, L, 2) normalised float32 coordinates, (S, L, 1) masks, (S, L) frame indices; S = T - L + 1."""
    T = len(x)
    S = max(T - seq_len + 1, 0)
    coor = np.zeros((S, seq_len, 2), np.float32)
    m = np.zeros((S, seq_len, 1), np.float32)
    idx = np.zeros((S, seq_len), np.int64)
    for s in range(S):
        for f in range(seq_len):
            coor[s, f] = (x[s + f], y[s + f])  # float32 storage of the integer pixel coordinates (dataset.py:390)
            m[s, f, 0] = mask[s + f]
            idx[s, f] = s + f
    w, h = img_wh
    coor[:, :, 0] = coor[:, :, 0] / w  # dataset.py:499-500 (float32 / python int)
    coor[:, :, 1] = coor[:, :, 1] / h
    return coor, m, idx


def coordinate_ensemble(seq_pred: torch.Tensor, total: int) -> torch.Tensor:
    """Closed form of ball_tracker.py:584-652 on (S, L, 2) per-window coordinates -> (total, 2): head = plain mean of the
    available windows, middle = weighted sum, tail = mean with divisor L - frame_i (same scheme as the heat-maps)."""
    S, L, _ = seq_pred.shape
    w = ensemble_weight(L)
    out = []
    for n in range(total if S >= 1 else 0):
        terms = []
        for k in range(L):
            s = n - (L - 1) + k
            terms.append(seq_pred[s, L - 1 - k] if 0 <= s < S else torch.zeros(2))
        t = torch.stack(terms)
        if n < S and n >= L - 1:
            e = (t * w[:, None]).sum(0)
        else:
            e = t.sum(0)
            e = e / ((n + 1) if n < S else (L - (n - (S - 1))))
        out.append(e)
    return torch.stack(out) if out else torch.zeros((0, 2))


@torch.no_grad()
def inpaint_stage(net, x, y, vis, video_wh, seq_len: int, net_hw=(288, 512), batch_size: int = 8):
    """TrackNet (x, y, vis) pixel lists -> inpainted {'Frame','X','Y','Visibility'} as ball_tracker.py:525-673."""
    W_img, H_img = video_wh
    HEIGHT, WIDTH = net_hw
    coor_th = 50.0 / math.sqrt(HEIGHT ** 2 + WIDTH ** 2)  # COOR_TH (ball_tracker.py:250-251)
    img_scaler = (W_img / WIDTH, H_img / HEIGHT)
    mask = generate_inpaint_mask(y, vis, th_h=H_img * 0.05)
    coor, m, idx = make_sequences(x, y, mask, seq_len, (W_img, H_img))
    T = len(x)
    if len(coor) == 0:
        return {"Frame": [], "X": [], "Y": [], "Visibility": []}, mask
    coor_t, m_t = torch.from_numpy(coor), torch.from_numpy(m)
    outs = []
    for i in range(0, len(coor_t), batch_size):
        c, mm = coor_t[i:i + batch_size], m_t[i:i + batch_size]
        o = net(c, mm)
        o = o * mm + c * (1 - mm)
        th = (o[:, :, 0] < coor_th) & (o[:, :, 1] < coor_th)
        o[th] = 0.0
        outs.append(o)
    seq_pred = torch.cat(outs)
    ens = coordinate_ensemble(seq_pred, T)
    th = (ens[:, 0] < coor_th) & (ens[:, 1] < coor_th)
    ens[th] = 0.0
    X, Y, V = [], [], []
    ens_np = ens.numpy()  # float32; predict.py:125-127 multiplies numpy float32 scalars (float32 arithmetic)
    for n in range(T):
        c_p = ens_np[n]
        cx = int(c_p[0] * WIDTH * img_scaler[0])
        cy = int(c_p[1] * HEIGHT * img_scaler[1])
        X.append(cx), Y.append(cy), V.append(0 if (cx == 0 and cy == 0) else 1)
    return {"Frame": list(range(T)), "X": X, "Y": Y, "Visibility": V}, mask
