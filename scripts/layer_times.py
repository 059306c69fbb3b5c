"""Per-op device times of the four programs (CUDA events between ops): which layers bind, and on what."""
import sys
import torch
sys.path.insert(0, ".")
import bench
from oracle import weights as OW
from padel_analytics_b200 import synth
from padel_analytics_b200.engine import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["ball", "players", "pose", "court"]
hw = (1080, 1920)
ckpts = {k: OW.make_yolo(k) for k in ("detect", "pose13", "court12")}
ckpts["tracknet"] = OW.make_tracknet()
tr, med = bench.build_trackers(B, hw, ckpts, "cuda")
fr = synth.make_frames(B, 1080, 1920, device="cuda")
for k in ("players", "pose", "court"):
    tr[k].detect_sample(fr)  # builds the program for the right input size
progs = {"ball": tr["ball"].tracknet.prog}
for k in ("players", "pose", "court"):
    progs[k] = list(tr[k].model._progs.values())[0]["prog"]
for name in which:
    p = progs[name]
    t = ops.time_program_ops(p, repeats=5)
    print(f"== {name}: {sum(t):.3f} ms total, conv {sum(x for x, kd in zip(t, p.kinds) if kd == 'conv'):.3f} ms")
    for i, (ms, kd, fl, by) in enumerate(zip(t, p.kinds, p.flops, p.bytes)):
        d = p.descs[i] if hasattr(p, "descs") else None
        shape = ""
        if d is not None and kd == "conv":
            shape = f"N{d.N} {d.H}x{d.W} cin{d.cin} cout{d.cout_pad} k{d.ksize} s{d.stride} m{d.out_mode}"
        print(f"{i:3d} {kd:5s} {ms*1e3:9.1f} us  {fl/ms/1e9 if ms else 0:8.1f} TF/s  {by/ms/1e6 if ms else 0:8.1f} GB/s  {shape}")
