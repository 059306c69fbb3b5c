"""How do consecutive small layers overlap?  GPU-wide nanosecond stamps (%globaltimer) of the first and last CTA of
each launch of a chain of identical convs: entry, after griddepcontrol.wait, exit.  Needs the debug library
(python -m padel_analytics_b200.build --debug).  Usage: python scripts/exp_chain_timeline.py [PDL 0|1]"""
import ctypes as C, os, sys
os.environ.setdefault("PADEL_B200_LIB", "padel_analytics_b200/libpadel_b200_debug.so")
import torch
sys.path.insert(0, ".")
from padel_analytics_b200 import _lib as L
from padel_analytics_b200.engine import ops
lib = L.lib()
lib.pb_debug_conv_timeline.restype = None
lib.pb_debug_conv_timeline.argtypes = [C.c_void_p]

def chain(tag, N, H, W, c, k, n=8):
    a = torch.randn(N, H, W, c, device="cuda").half()
    b = torch.zeros_like(a)
    w = torch.randn(c, c, k, k) * 0.05
    wp, bp = ops.pack_conv_weight(w, torch.zeros(c), c, c, "cuda")
    bufs = [torch.zeros(4 * 64 * 4, dtype=torch.int64, device="cuda") for _ in range(n)]
    P = ops.Program()
    for i in range(n):
        lib.pb_debug_conv_timeline(bufs[i].data_ptr())
        x, y = (a, b) if i % 2 == 0 else (b, a)
        P.conv(ops.make_conv_desc(x, 0, c, wp, bp, k, 1, L.ACT_SILU, y, 0, L.OUT_F16_NHWC))
    lib.pb_debug_conv_timeline(None)
    for _ in range(3):
        P.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); P.run(); e1.record(); torch.cuda.synchronize()
    t = [bf.cpu().view(4, 64, 4)[3] for bf in bufs]
    t0 = int(t[0][0, 0])
    print(f"--- {tag}: N{N} {H}x{W} c{c} k{k}, {n} launches, {e0.elapsed_time(e1) * 1e3 / n:.1f} us per launch (PDL {os.environ.get('PADEL_B200_PDL', '1')})")
    print("launch | first CTA: entry  after-wait  exit | last CTA: entry  after-wait  exit   (us since launch 0 entry)")
    for i in range(n):
        r = lambda v: (int(v) - t0) / 1e3
        print(f"{i:3d} | {r(t[i][0,0]):8.2f} {r(t[i][0,1]):8.2f} {r(t[i][0,2]):8.2f} | {r(t[i][1,0]):8.2f} {r(t[i][1,1]):8.2f} {r(t[i][1,2]):8.2f}")

chain("P4 bottleneck conv", 32, 24, 40, 64, 3)
chain("P5 bottleneck conv", 32, 12, 20, 128, 3)
chain("P5 1x1", 32, 12, 20, 256, 1)
chain("P3 bottleneck conv", 32, 48, 80, 32, 3)
