"""Per-role clock64 timeline of CTA 0's first tiles for a given conv shape (bring-up diagnostics)."""
import ctypes as C, os, sys
os.environ.setdefault("PADEL_B200_LIB", "padel_analytics_b200/libpadel_b200_debug.so")  # python -m padel_analytics_b200.build --debug
import torch
sys.path.insert(0, ".")
from padel_analytics_b200 import _lib as L
from padel_analytics_b200.engine import ops
lib = L.lib()
lib.pb_debug_conv_timeline.restype = None
lib.pb_debug_conv_timeline.argtypes = [C.c_void_p]
def run(N, H, W, cin, cout, k, s=1, mode=L.OUT_F16_NHWC, act=L.ACT_SILU):
    x = torch.randn(N, H, W, cin, device="cuda").half()
    w = torch.randn(cout, cin, k, k) * 0.05
    wp, bp = ops.pack_conv_weight(w, torch.zeros(cout), cin, ops.pad16(cout), "cuda")
    out = torch.zeros(N, H // s, W // s, ops.pad16(cout), device="cuda", dtype=torch.float16)
    buf = torch.zeros(4 * 64 * 4, dtype=torch.int64, device="cuda")
    d = ops.make_conv_desc(x, 0, cin, wp, bp, k, s, act, out, 0, mode)
    for _ in range(2):
        ops.conv2d(d)
    lib.pb_debug_conv_timeline(buf.data_ptr())
    ops.conv2d(d)
    torch.cuda.synchronize()
    lib.pb_debug_conv_timeline(None)
    t = buf.cpu().view(4, 64, 4)
    t0 = int(t[0, 0, 0])
    print(f"--- N{N} {H}x{W} cin{cin} cout{cout} k{k} s{s}: cycles relative to first producer stamp")
    print("tile | prod start,end | mma start, got-acc, got-full, committed | epi start, got-tmem_full, done")
    for i in list(range(0, 14)) + [30, 31, 32, 33]:
        r = lambda a: int(a) - t0
        print(f"{i:3d} | {r(t[0,i,0]):7d} {r(t[0,i,1]):7d} | {r(t[1,i,0]):7d} {r(t[1,i,1]):7d} {int(t[1,i,2]):7d} {r(t[1,i,3]):7d} | {r(t[2,i,0]):7d} {r(t[2,i,1]):7d} {r(t[2,i,2]):7d}")
import os
os.environ["PADEL_B200_CONV_HALO"] = "1"
print("halo: prod = [tile start, got a_empty] ; mma = [start, got acc, OPERAND-WAIT CYCLES (not a timestamp), committed] ; epi = [start, got tmem_full, done]")
run(32, 288, 512, 64, 64, 3, act=L.ACT_RELU)
run(32, 288, 512, 192, 64, 3, act=L.ACT_RELU)
run(32, 160, 160, 32, 32, 3)
