"""Epilogue cost by activation: same conv shapes with SiLU / ReLU / none (is the SiLU math the limiter?)."""
import sys
import torch
sys.path.insert(0, ".")
from padel_analytics_b200 import _lib as L
from padel_analytics_b200.engine import ops

def t(N, H, W, cin, cout, k, s, act):
    x = torch.randn(N, H, W, cin, device="cuda").half()
    w = torch.randn(cout, cin, k, k) * 0.05
    wp, bp = ops.pack_conv_weight(w, torch.zeros(cout), cin, ops.pad16(cout), "cuda")
    out = torch.zeros(N, H // s, W // s, ops.pad16(cout), device="cuda", dtype=torch.float16)
    d = ops.make_conv_desc(x, 0, cin, wp, bp, k, s, act, out, 0, L.OUT_F16_NHWC)
    for _ in range(3):
        ops.conv2d(d)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.conv2d(d)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 * 1e3

for shape in [(32, 160, 160, 64, 64, 1, 1), (32, 320, 320, 32, 32, 1, 1), (32, 320, 320, 16, 16, 3, 1),
              (32, 320, 320, 64, 32, 1, 1), (32, 160, 160, 64, 192, 3, 1), (32, 640, 640, 16, 32, 3, 2),
              (32, 80, 80, 128, 128, 3, 1), (32, 160, 160, 32, 32, 3, 1)]:
    r = {n: t(*shape, a) for n, a in (("silu", L.ACT_SILU), ("relu", L.ACT_RELU), ("none", L.ACT_NONE))}
    N, H, W, cin, cout, k, s = shape
    gb = N * H * W * cin * 2 / 1e9 + N * (H // s) * (W // s) * cout * 2 / 1e9
    print(shape, {n: round(v, 1) for n, v in r.items()}, f"bytes {gb:.3f} GB -> {gb / r['silu'] * 1e3:.2f} / {gb / r['none'] * 1e3:.2f} TB/s")
