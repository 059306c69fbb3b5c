"""Per-layer lower bounds vs measured times (input: the per-op table written by scripts/layer_times.py).
Bounds (B200, this pool's measured peaks): tensor = FLOPs / min(1.44 PFLOP/s sustained, issue cap at small N: an
M128xNxK16 UMMA takes max(60, N/2) cycles, CTA pairs double M); hbm = (input + output bytes) / 6.57 TB/s;
store = fp16/fp32 output bytes / (148 SMs x 16 B/clk x 1.965 GHz) (scripts/exp_epi_bench.py)."""
import re, sys

PEAK, HBM, CLK, SMS = 1439.9e12, 6573.8e9, 1.965e9, 148
pat = re.compile(r"\s*(\d+)\s+conv\s+([\d.]+) us\s+([\d.]+) TF/s\s+([\d.]+) GB/s\s+N(\d+) (\d+)x(\d+) cin(\d+) cout(\d+) k(\d) s(\d) m(\d)")
model = None
tot = {}
rows = []
for line in open(sys.argv[1]):
    if line.startswith("=="):
        model = line.split()[1].rstrip(":")
        continue
    m = pat.match(line)
    if not m:
        continue
    i, us, tf, gbs, N, H, W, cin, cout, k, s, mode = m.groups()
    us, tf, gbs = float(us), float(tf), float(gbs)
    N, H, W, cin, cout, k, s, mode = map(int, (N, H, W, cin, cout, k, s, mode))
    flops = tf * 1e12 * us * 1e-6
    byts = gbs * 1e9 * us * 1e-6
    Ho, Wo = H // s, W // s
    esz = 4 if mode in (2, 3) else 2
    out_b = N * Ho * Wo * cout * esz * (4 if mode == 1 else 1)
    npad = max(16, (cout + 15) // 16 * 16)
    n_tile = min(npad, 256)
    cap = PEAK if n_tile >= 128 else PEAK * 0 + SMS * CLK * (2 * 128 * n_tile * 16) / max(60.0, n_tile / 2)
    t_tensor = flops / min(PEAK, cap) * 1e6
    t_hbm = byts / HBM * 1e6
    t_store = out_b / (SMS * 16 * CLK) * 1e6
    bound = max(t_tensor, t_hbm, t_store)
    which = "tensor" if bound == t_tensor else ("hbm" if bound == t_hbm else "store")
    rows.append((model, int(i), f"N{N} {H}x{W} {cin}->{cout} k{k}s{s} m{mode}", us, t_tensor, t_hbm, t_store, which, bound / us))
    a = tot.setdefault(model, [0.0, 0.0])
    a[0] += us
    a[1] += bound
print("| model | op | layer | measured us | tensor us | hbm us | store us | binding | bound/measured |")
print("|---|---|---|---|---|---|---|---|---|")
for r in rows:
    if r[3] >= 40:
        print(f"| {r[0]} | {r[1]} | {r[2]} | {r[3]:.0f} | {r[4]:.0f} | {r[5]:.0f} | {r[6]:.0f} | {r[7]} | {r[8]:.2f} |")
print()
for k_, (a, b) in tot.items():
    print(f"{k_}: convs measured {a/1e3:.2f} ms, sum of per-layer bounds {b/1e3:.2f} ms ({b/a:.2f})")
