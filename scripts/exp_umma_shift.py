"""Experiment: tcgen05 K-major swizzled A operand (SW128/64/32) with a row-shifted start and non-canonical group stride.
Result (B200): the swizzle XOR is applied on absolute shared-memory address bits, so base_offset = 0 is exact for
any row shift and any SBO -> conv taps can be served from one halo tile."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from padel_analytics_b200 import _lib as L

lib = L.lib()
fn = lib.pb_debug_umma_shift
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
for rb in (128, 64, 32):
    K = rb // 2
    r = torch.arange(192).view(192, 1)
    c = torch.arange(K).view(1, K)
    A = ((r % 64) * 32 + (c % 32)).half().cuda().contiguous()
    Bm = torch.zeros(64, K)
    Bm[:K, :K] = torch.eye(K)
    Bm = Bm.half().cuda().contiguous()
    for sbo_rows in (8, 10):
        for shift in (0, 1, 3, 9, 11):
            for mode in (0, 1):
                D = torch.zeros(128, 64, device="cuda")
                L.check(fn(A.data_ptr(), Bm.data_ptr(), D.data_ptr(), shift, sbo_rows * rb, mode, rb, L.stream_ptr()))
                torch.cuda.synchronize()
                m = torch.arange(128)
                rows = shift + (m // 8) * sbo_rows + (m % 8)
                ok = rows < 192
                exp = A.float().cpu()[rows.clamp(max=191)]
                got = D.cpu()[:, :K]
                match = (got == exp)[ok]
                print(f"row_bytes={rb:3d} sbo_rows={sbo_rows:2d} shift={shift:2d} base_off_mode={mode}: "
                      f"exact={bool(match.all())} frac={match.float().mean():.3f}")
