"""clock64 timeline (CTA 0, first tiles) of the large-spatial / few-channel YOLO-pose layers (HBM-bound ones)."""
import ctypes as C, os, sys
os.environ.setdefault("PADEL_B200_LIB", "padel_analytics_b200/libpadel_b200_debug.so")  # python -m padel_analytics_b200.build --debug
import torch
sys.path.insert(0, ".")
from padel_analytics_b200 import _lib as L
from padel_analytics_b200.engine import ops
lib = L.lib()
lib.pb_debug_conv_timeline.restype = None
lib.pb_debug_conv_timeline.argtypes = [C.c_void_p]

def show(tag, d, keep):
    buf = torch.zeros(4 * 64 * 4, dtype=torch.int64, device="cuda")
    for _ in range(2):
        ops.conv2d(d)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.conv2d(d)
    e1.record(); torch.cuda.synchronize()
    lib.pb_debug_conv_timeline(buf.data_ptr())
    ops.conv2d(d)
    torch.cuda.synchronize()
    lib.pb_debug_conv_timeline(None)
    t = buf.cpu().view(4, 64, 4)
    t0 = int(t[0, 0, 0])
    print(f"--- {tag}: {e0.elapsed_time(e1) / 5 * 1e3:.1f} us/launch")
    print("tile | prod a,b | mma start, got-acc, c, committed | epi start, got-tmem_full, done")
    for i in list(range(0, 10)) + [30, 31, 32, 33, 34, 35]:
        r = lambda a: int(a) - t0
        print(f"{i:3d} | {r(t[0,i,0]):7d} {r(t[0,i,1]):7d} | {r(t[1,i,0]):7d} {r(t[1,i,1]):7d} {int(t[1,i,2]) if abs(int(t[1,i,2])) < 10**7 else r(t[1,i,2]):7d} {r(t[1,i,3]):7d} | {r(t[2,i,0]):7d} {r(t[2,i,1]):7d} {r(t[2,i,2]):7d}")

def run(N, H, W, cin, cout, k, s=1, act=L.ACT_SILU):
    x = torch.randn(N, H, W, cin, device="cuda").half()
    w = torch.randn(cout, cin, k, k) * 0.05
    wp, bp = ops.pack_conv_weight(w, torch.zeros(cout), cin, ops.pad16(cout), "cuda")
    out = torch.zeros(N, H // s, W // s, ops.pad16(cout), device="cuda", dtype=torch.float16)
    d = ops.make_conv_desc(x, 0, cin, wp, bp, k, s, act, out, 0, L.OUT_F16_NHWC)
    show(f"N{N} {H}x{W} cin{cin} cout{cout} k{k} s{s}", d, (x, wp, bp, out))

def stem(N, H, W, cout):
    xp = torch.randn(N, H + 2, W + 2, 4, device="cuda").half()
    w = torch.randn(cout, 3, 3, 3) * 0.05
    wp, bp = ops.pack_stem_weight(w, torch.zeros(cout), ops.pad16(cout), "cuda")
    out = torch.zeros(N, H // 2, W // 2, ops.pad16(cout), device="cuda", dtype=torch.float16)
    d = ops.make_stem_desc(xp, wp, bp, L.ACT_SILU, out)
    show(f"stem N{N} {H}x{W} cout{cout}", d, (xp, wp, bp, out))

stem(32, 1280, 1280, 16)
run(32, 640, 640, 16, 32, 3, 2)
run(32, 320, 320, 32, 32, 1)
run(32, 320, 320, 16, 16, 3)
run(32, 320, 320, 64, 32, 1)
run(32, 160, 160, 64, 64, 1)
run(32, 160, 160, 64, 192, 3)
