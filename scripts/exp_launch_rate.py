"""GPU front-end cost per kernel launch vs launch shape (debug library): entry-to-entry period of chains of empty
kernels.  See csrc/debug/debug_launch.cu.  Usage: python scripts/exp_launch_rate.py"""
import ctypes as C, os, sys
os.environ.setdefault("PADEL_B200_LIB", "padel_analytics_b200/libpadel_b200_debug.so")
import torch
sys.path.insert(0, ".")
from padel_analytics_b200 import _lib as L
lib = L.lib()
fn = lib.pb_debug_launch_chain
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
names = ["148x480, no smem, int param", "+ 120 KB dynamic smem", "+ 640-byte struct param",
         "+ two CUtensorMap __grid_constant__ params", "same, programmatic stream serialization",
         "120 KB smem + tcgen05.alloc/dealloc 512 columns", "+ mbarrier init", "+ one 4 KB TMA load", "+ one UMMA + commit",
         "+ 32-byte global store per thread", "+ 32 more TMA boxes in flight at exit (waited)", "variant 10 on 37 CTAs", "variant 10 on 74 CTAs"]
n = 40
for v, name in enumerate(names):
    out = torch.zeros(n, dtype=torch.int64, device="cuda")
    for _ in range(3):
        L.check(fn(out.data_ptr(), v, n, L.stream_ptr()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(fn(out.data_ptr(), v, n, L.stream_ptr()))
    e1.record()
    torch.cuda.synchronize()
    t = out.cpu()
    d = (t[1:] - t[:-1]).float() / 1e3
    print(f"variant {v} ({name}): entry-to-entry period median {d.median():.2f} us, min {d.min():.2f}, max {d.max():.2f}; "
          f"events {e0.elapsed_time(e1) * 1e3 / n:.2f} us per launch")
