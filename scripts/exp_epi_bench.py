"""Cost of one 16-channel epilogue chunk per warp vs number of epilogue warps and enabled parts (csrc/debug_epi.cu)."""
import ctypes as C, os, sys
os.environ.setdefault("PADEL_B200_LIB", "padel_analytics_b200/libpadel_b200_debug.so")  # python -m padel_analytics_b200.build --debug
import torch
sys.path.insert(0, ".")
from padel_analytics_b200 import _lib as L
lib = L.lib()
lib.pb_debug_epi_bench.restype = C.c_int
lib.pb_debug_epi_bench.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
iters = 2048
names = {1: "ld", 2: "bias+pack+store", 3: "ld+store", 7: "ld+store+silu", 15: "ld+store+silu2", 6: "store+silu",
         19: "ld(pipelined)+store", 23: "ld(pipelined)+store+silu", 31: "ld(pipelined)+store+silu2", 0: "bias+pack only"}
print("cycles per chunk per warp (mean over warps) | per-SM chunk interval = that / nwarps")
for parts in (0, 1, 2, 3, 19, 6, 7, 23, 15, 31):
    row = []
    for nw in (4, 8, 12, 16):
        cyc = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
        out = torch.zeros(148 * nw * 32 * 64, dtype=torch.float16, device="cuda")
        for _ in range(2):
            L.check(lib.pb_debug_epi_bench(cyc.data_ptr(), out.data_ptr(), nw, parts, iters, None))
        torch.cuda.synchronize()
        c = cyc.view(148, 16)[:, :nw].float().mean().item() / iters
        row.append(f"{nw:2d}w {c:7.1f} ({c / nw:6.1f})")
    print(f"{names[parts]:28s} " + "  ".join(row))
