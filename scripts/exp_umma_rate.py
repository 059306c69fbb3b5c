"""Micro-benchmark: cycles per tcgen05.mma (M=128, K=16) with A from shared memory vs A staged through TMEM."""
import ctypes as C, os, sys
os.environ.setdefault("PADEL_B200_LIB", "padel_analytics_b200/libpadel_b200_debug.so")  # python -m padel_analytics_b200.build --debug
import torch
sys.path.insert(0, ".")
from padel_analytics_b200 import _lib as L
lib = L.lib()
fn = lib.pb_debug_umma_rate
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
out = torch.zeros(148, dtype=torch.int64, device="cuda")
iters = 4096
for N in (16, 32, 64, 128, 256):
    for mode in (0, 1):
        for _ in range(2):
            L.check(fn(out.data_ptr(), N, mode, iters, L.stream_ptr()))
            torch.cuda.synchronize()
        cyc = out.float().mean().item() / iters
        print(f"N={N:3d} mode={'SS (A in smem)' if mode == 0 else 'TS (tcgen05.cp A -> TMEM)'}: {cyc:7.1f} cycles per UMMA "
              f"(math floor {max(N, 32) * 128 / 256 if False else 128 * N / 256:.0f})")
