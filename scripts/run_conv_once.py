"""Run one conv shape a few times -- target for `ncu --set full --import-source on` captures.
usage: run_conv_once.py N H W cin cout k s [act]"""
import sys
import torch
sys.path.insert(0, ".")
from padel_analytics_b200 import _lib as L
from padel_analytics_b200.engine import ops

N, H, W, cin, cout, k, s = (int(a) for a in sys.argv[1:8])
act = int(sys.argv[8]) if len(sys.argv) > 8 else L.ACT_SILU
x = torch.randn(N, H, W, cin, device="cuda").half()
w = torch.randn(cout, cin, k, k) * 0.05
wp, bp = ops.pack_conv_weight(w, torch.zeros(cout), cin, ops.pad16(cout), "cuda")
out = torch.zeros(N, H // s, W // s, ops.pad16(cout), device="cuda", dtype=torch.float16)
d = ops.make_conv_desc(x, 0, cin, wp, bp, k, s, act, out, 0, L.OUT_F16_NHWC)
for _ in range(4):
    ops.conv2d(d)
torch.cuda.synchronize()
print("done")
