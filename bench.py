#!/usr/bin/env python
"""Headline benchmark: frames/sec through the four trackers (BASELINE.json metric) on synthetic 1080p frames.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch 32] [--res 1080p|4k]

One *step* = one batch of `--batch` frames through all four trackers' hot path (PlayerTracker YOLOv8n-detect,
PlayerKeypointsTracker YOLOv8n-pose 13x3 @1280, KeypointsTracker YOLOv8n-pose 12x3 @640, BallTracker TrackNet 27->8).
N > 1 (torchrun, one rank per GPU): every rank runs its own shard of frames (weak scaling, no data-path collective;
NCCL only broadcasts the weights at init and gathers detection counts at the end).

Printed JSON (rank 0, one line): see the repository prompt's contract — value (device-resident frames), e2e (pinned
host frames through the tracker API, H2D and result D2H inside the timed region), roofline (dominant kernel =
conv_tc_kernel, event-timed live), cpu_baseline (the CPU oracle on this box's host cores, bounded sample), clocks.
`--impl reference` times that CPU oracle as the main arm (the reference's own Python path cannot travel to the GPU
box: ultralytics/supervision are not installed anywhere; oracle/ restates it — DESIGN.md §oracle).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

RES = {"1080p": (1080, 1920), "4k": (2160, 3840), "720p": (720, 1280)}


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(tflops_burst=d.get("bf16_tflops"), tflops_sustained=d.get("bf16_tflops_sustained"),
                    hbm_gbs=d.get("hbm_gbs"), source="measured")
    return dict(tflops_burst=1590.0, tflops_sustained=1400.0, hbm_gbs=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])), mx.append(float(r[2]))
            except Exception:  # noqa: BLE001
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (port of the reference's CPU path) on the host cores, bounded sample
# ------------------------------------------------------------------------------------------------------------
def cpu_threads() -> int:
    """Threads given to the CPU oracle: every host core up to 64 (beyond that PyTorch's CPU convolutions on these
    small networks get slower, not faster)."""
    return max(1, min(os.cpu_count() or 1, 64))


def cpu_oracle_fps(hw, n_yolo=2, n_ball=10, seed=1234):
    """All-four-trackers frames/s of the CPU oracle: N / sum_t time_t(N) measured per tracker on small samples
    (YOLO trackers: n_yolo frames; ball: n_ball frames -> n_ball-7 windows) and normalised per frame."""
    import cv2
    from PIL import Image

    from oracle import tracknet as OT
    from oracle import weights as OW
    from oracle import yolov8 as OY
    from padel_analytics_b200 import synth

    torch.set_num_threads(cpu_threads())
    H, W = hw
    frames = [f.numpy() for f in synth.make_frames(max(n_yolo, n_ball), H, W, seed=seed)]
    per_frame = {}
    with torch.no_grad():
        for kind, imgsz, conf, classes, max_det in (("detect", 640, 0.5, [0], 300), ("pose13", 1280, 0.25, [0], 300),
                                                    ("court12", 640, 0.5, None, 12)):
            yolo = OY.YOLO(OW.load_yolo(OW.make_yolo(kind)))
            t0 = time.perf_counter()
            if kind == "detect":  # players_tracker.py:346-359
                sample = [cv2.cvtColor(f, cv2.COLOR_BGR2RGB) for f in frames[:n_yolo]]
            else:  # players_keypoints_tracker.py:260-292 / keypoints_tracker.py:190-245
                sample = [Image.fromarray(cv2.cvtColor(f, cv2.COLOR_BGR2RGB)).resize((imgsz, imgsz))
                          for f in frames[:n_yolo]]
            yolo.predict(sample, conf=conf, iou=0.7, imgsz=imgsz, classes=classes, max_det=max_det)
            per_frame[kind] = (time.perf_counter() - t0) / n_yolo
        net = OW.load_tracknet(OW.make_tracknet())
        med = synth.make_median(H, W, seed=seed).numpy()
        t0 = time.perf_counter()
        OT.run_ball_oracle(net, frames[:n_ball], med, (W, H), batch_size=8)
        per_frame["ball"] = (time.perf_counter() - t0) / (n_ball - 7)  # one window per frame in steady state
    return 1.0 / sum(per_frame.values()), per_frame


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    hw = RES[args.res]
    vals = []
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_oracle_fps(hw, n_yolo=1, n_ball=8)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fps, per = cpu_oracle_fps(hw, n_yolo=1, n_ball=8)
        vals.append(fps)
    dt = time.perf_counter() - t0
    v = statistics.median(vals)
    cores = cpu_threads()
    sample = "per step: 1 frame per YOLO tracker + 8 frames (1 window) ball, per-frame times summed"
    print(json.dumps({
        "impl": "reference", "metric": "frames/sec through trackers.runner (all 4 trackers)", "value": v,
        "unit": "frames/s", "n_gpus": 0, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"all-four trackers (YOLOv8n detect/pose13@1280/court12@640 + TrackNet), {args.res}, "
                               f"CPU oracle", "per_frame_s": per},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), file=JSON_OUT, flush=True)


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
def build_trackers(B, hw, ckpts, dev):
    from padel_analytics_b200 import synth
    from padel_analytics_b200.trackers import BallTracker, KeypointsTracker, PlayerKeypointsTracker, PlayerTracker
    from padel_analytics_b200.trackers import sv_compat as sv

    H, W = hw
    vi = sv.VideoInfo(width=W, height=H, fps=30.0, total_frames=None)
    poly = sv.PolygonZone(np.array([[W // 10, H // 10], [9 * W // 10, H // 10], [9 * W // 10, 9 * H // 10],
                                    [W // 10, 9 * H // 10]]), frame_resolution_wh=(W, H))
    med = synth.make_median(H, W).numpy()
    tr = {
        "players": PlayerTracker(ckpts["detect"], poly, batch_size=B),
        "pose": PlayerKeypointsTracker(ckpts["pose13"], 1280, batch_size=B, load_path=None, save_path=None),
        "court": KeypointsTracker(ckpts["court12"], batch_size=B, model_type="yolo"),
        "ball": BallTracker(ckpts["tracknet"], None, batch_size=B, median=med),
    }
    for t in tr.values():
        t.video_info_post_init(vi)
    return tr, med


# The contract is ONE JSON line on stdout: everything else this process prints (tracker banners, library chatter)
# is sent to stderr, the JSON line goes to the real stdout.
JSON_OUT = sys.stdout


def main():
    sys.stdout = sys.stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--res", default="1080p", choices=list(RES))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-host", action="store_true", help="cProfile the timed region's host side (stderr)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch.distributed as dist

    # The only oracle import of the product arm: it CONSTRUCTS the seeded synthetic checkpoints (a stand-in for
    # torch.load of real .pt files, none of which exist offline) before anything is timed.  No oracle code computes
    # anything inside the warm-up or timed regions; the trackers below run on libpadel_b200.so only.
    from oracle import weights as OW
    from padel_analytics_b200 import _lib as L
    from padel_analytics_b200 import synth
    from padel_analytics_b200.engine import ops
    from padel_analytics_b200.engine.tracknet_engine import bbox_to_xyv

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # weights: generated on rank 0, broadcast over NCCL (the only init-time collective)
    if rank == 0:
        # sparse heads: a handful of players per frame like a real padel rally (the dense defaults are for parity tests)
        ckpts = {"detect": OW.make_yolo("detect", cls_mean=-5.0), "pose13": OW.make_yolo("pose13", cls_mean=-5.7),
                 "court12": OW.make_yolo("court12")}
        ckpts["tracknet"] = OW.make_tracknet()
    else:
        ckpts = None
    if world > 1:
        box = [ckpts]
        dist.broadcast_object_list(box, src=0, device=dev)
        ckpts = box[0]

    B = args.batch
    hw = RES[args.res]
    H, W = hw
    trackers, med = build_trackers(B, hw, ckpts, dev)
    ball = trackers["ball"]

    # frames: NBUF distinct batches resident in HBM (+ pinned host copies for the e2e leg); each batch (B*H*W*3 bytes
    # = 199 MB at 1080p/32) alone exceeds the 126 MB L2 and activations are GBs, so no L2 flush is needed.
    NBUF = 3
    dev_batches = [synth.make_frames(B, H, W, start=rank * 100000 + i * B, device=dev) for i in range(NBUF)]
    host_batches = [b.cpu().pin_memory() for b in dev_batches]
    from padel_analytics_b200.trackers.runner import FusedPass

    # The measured path is the fused single pass (trackers/runner.py::FusedPass): one upload per batch shared by the
    # four trackers, device work of all four enqueued back to back, host post-processing overlapped.
    named = {"players": trackers["players"], "pose": trackers["pose"], "court": trackers["court"], "ball": ball}
    fused = FusedPass(named, hw, B, total_frames=10 ** 9)  # steady state: the tail flush is never reached
    ball._pipe.push_frames(dev_batches[0][:7])  # prime the 8-frame window so every step yields B windows

    def run_steps(batches, steps):
        nd = 0
        for out in fused.run(batches[i % NBUF] for i in range(steps)):
            nd += sum(len(p) for p in out["players"]) + sum(len(p) for p in out["pose"])
        return nd

    import gc

    def timed(batches, steps):
        gc.collect()
        gc.freeze()  # keep the (large, static) engine object graph out of the cyclic collector's way
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = L.lib().pb_launch_count()
        t0 = time.perf_counter()
        e0.record()
        nd = run_steps(batches, steps)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, wall * 1e3], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.barrier()
            ms, wall = t[0].item(), t[1].item() / 1e3
        return ms, wall, L.lib().pb_launch_count() - l0, nd

    run_steps(dev_batches, args.warmup)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if args.profile_host:
        import cProfile
        import pstats

        pr = cProfile.Profile()
        pr.enable()
    ms_dev, wall_dev, launches, ndet = timed(dev_batches, args.steps)
    if args.profile_host:
        pr.disable()
        pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(35)
    run_steps(host_batches, 2)
    ms_e2e, wall_e2e, _, _ = timed(host_batches, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    frames_total = B * args.steps * world
    value = frames_total / (max(ms_dev, wall_dev * 1e3) / 1e3)
    e2e = frames_total / (max(ms_e2e, wall_e2e * 1e3) / 1e3)

    # roofline of the dominant kernel (conv_tc_kernel): algorithmic FLOPs of every conv launch of one step divided
    # by the event-timed duration of those launches (per-op CUDA events on the launch stream, rank 0 only)
    roof = None
    if rank == 0:
        progs = [("tracknet", ball.tracknet.prog)]
        for k in ("players", "pose", "court"):
            for st in trackers[k].model._progs.values():
                progs.append((k, st["prog"]))
        per_kernel = {}
        per_model = {}
        all_ms = 0.0
        for name, p in progs:
            t = ops.time_program_ops(p, repeats=3)
            kn = p.op_kernels()
            for ti, k, f, by in zip(t, kn, p.flops, p.bytes):
                e = per_kernel.setdefault(k, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
                e["ms"] += ti
                e["flops"] += f
                e["bytes"] += by
                e["launches"] += 1
            cm = sum(ti for ti, kd in zip(t, p.kinds) if kd == "conv")
            cf = sum(f for f, kd in zip(p.flops, p.kinds) if kd == "conv")
            all_ms += sum(t)
            per_model[name] = {"conv_ms": round(cm, 3), "all_ops_ms": round(sum(t), 3),
                               "gflop_per_frame": round(cf / B / 1e9, 3), "tflops": round(cf / cm / 1e9, 1)}
        pk = _peaks()
        dom = max((k for k in per_kernel if k.startswith("conv")), key=lambda k: per_kernel[k]["ms"])
        d = per_kernel[dom]
        achieved = d["flops"] / (d["ms"] / 1e3) / 1e12
        conv_ms = sum(v["ms"] for k, v in per_kernel.items() if k.startswith("conv"))
        conv_fl = sum(v["flops"] for k, v in per_kernel.items() if k.startswith("conv"))
        traffic = None
        tf = ROOT / "profiles" / "r01_tracknet_dram_bytes.json"
        if tf.exists() and dom == "conv_halo_kernel":
            traffic = json.loads(tf.read_text())
        roof = {"bound": "tensor", "kernel": dom, "achieved": round(achieved, 1), "peak": pk["tflops_sustained"],
                "peak_kind": f"{pk['source']} cuBLAS bf16 sustained", "unit": "TFLOP/s",
                "frac": round(achieved / pk["tflops_sustained"], 4),
                "traffic": (traffic or {}).get("dram_gb_per_step_tracknet_halo_launches"),
                "traffic_note": (traffic or {}).get("note"),
                "launches_per_step": d["launches"], "kernel_ms_per_step": round(d["ms"], 3),
                "algorithmic_gflop_per_step": round(d["flops"] / 1e9, 1),
                "algorithmic_act_gb_per_step": round(d["bytes"] / 1e9, 2),
                "all_conv_kernels": {"ms_per_step": round(conv_ms, 3), "tflops": round(conv_fl / conv_ms / 1e9, 1),
                                     "frac": round(conv_fl / conv_ms / 1e9 / pk["tflops_sustained"], 4)},
                "program_ms_per_step": round(all_ms, 3),
                "per_kernel": {k: {"ms": round(v["ms"], 3), "launches": v["launches"],
                                   "tflops": round(v["flops"] / v["ms"] / 1e9, 1) if v["flops"] else 0.0,
                                   "gbs": round(v["bytes"] / v["ms"] / 1e6, 1)} for k, v in per_kernel.items()},
                "per_model": per_model}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        fps, per = cpu_oracle_fps(hw)
        cpu = {"value": round(fps, 4), "unit": "frames/s", "cores": cpu_threads(), "kind": "port",
               "sample": "2 frames per YOLO tracker + 10 frames (3 windows) ball on the host cores, "
                         "per-frame times summed over the four trackers",
               "per_frame_s": {k: round(v, 4) for k, v in per.items()}}

    if world > 1:
        cnt = torch.tensor([ndet], device=dev)
        gathered = [torch.zeros_like(cnt) for _ in range(world)]
        dist.all_gather(gathered, cnt)  # detection counts gathered to every rank (tiny)
        ndet = int(sum(int(g.item()) for g in gathered))

    if rank == 0:
        h2d = B * H * W * 3  # one pinned-host -> device upload per batch, shared by the four trackers
        d2h = sum(int(np.prod(st[k]["host"][0][0].shape)) * 4 for t in ("players", "pose", "court")
                  for st in trackers[t].model._progs.values() for k in st if isinstance(k, tuple)) + (B + 7) * 16
        print(json.dumps({
            "metric": "frames/sec through trackers.runner (all 4 trackers)", "value": round(value, 2),
            "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(max(ms_dev, wall_dev * 1e3) / args.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16 storage, f32 accumulate", "data": "synthetic",
            "config": {"workload": f"all four trackers, synthetic {args.res} frames, batch_size={B} per GPU "
                                   f"(BASELINE.json configs[1]); YOLOv8n detect@384x640 + pose13x3@1280 + "
                                   f"court12x3@640 + TrackNet 27->8@288x512, seeded random weights",
                       "global_batch": B * world, "l2": "inputs (199 MB/batch) and activations exceed L2; no flush",
                       "parallelism": f"frames sharded over {world} GPU(s), no data-path collective",
                       "detections_in_timed_region": ndet,
                       "pass": "fused single pass: one upload per batch shared by the four trackers, one batch of "
                               "look-ahead, YOLO chains on their own streams"},
            "e2e": {"value": round(e2e, 2), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(max(ms_e2e, wall_e2e * 1e3) / args.steps, 3)},
            "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
            "timing": {"device_ms_per_step": round(ms_dev / args.steps, 3), "wall_ms_per_step": round(wall_dev * 1e3 / args.steps, 3),
                       "e2e_device_ms_per_step": round(ms_e2e / args.steps, 3),
                       "e2e_wall_ms_per_step": round(wall_e2e * 1e3 / args.steps, 3)},
        }), file=JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
