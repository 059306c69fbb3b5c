"""Experiment: tcgen05 SWIZZLE_128B K-major A operand with a 128B-row-shifted start and non-1024B group stride."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from padel_analytics_b200 import _lib as L

lib = L.lib()
fn = lib.pb_debug_umma_shift
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
r = torch.arange(192).view(192, 1)
c = torch.arange(64).view(1, 64)
A = ((r % 64) * 32 + (c % 32)).half().cuda().contiguous()      # value identifies (row % 64, col % 32) exactly
Bm = torch.eye(64).half().cuda().contiguous()                    # D[m][n] = A[row(m)][n]
for sbo in (1024, 1280, 2048):
    for shift in (0, 1, 2, 3, 5, 8, 9):
        for mode in (0, 1):
            D = torch.zeros(128, 64, device="cuda")
            L.check(fn(A.data_ptr(), Bm.data_ptr(), D.data_ptr(), shift, sbo, mode, L.stream_ptr()))
            torch.cuda.synchronize()
            m = torch.arange(128)
            rows = shift + (m // 8) * (sbo // 128) + (m % 8)
            ok_rows = rows < 192
            exp = A.float().cpu()[rows.clamp(max=191)]
            got = D.cpu()
            match = (got == exp)[ok_rows]
            rowhit = ((got[:, 0] // 32) == (exp[:, 0] // 32))[ok_rows]
            print(f"sbo={sbo:5d} shift={shift} base_off_mode={mode}: exact={bool(match.all())} "
                  f"frac={match.float().mean():.3f} row_ok={rowhit.float().mean():.3f}")
