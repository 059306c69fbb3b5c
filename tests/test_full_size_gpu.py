"""BASELINE.json's full-size configurations, checked through size-independent properties: a frame's result must not
depend on the batch it travels in (batch invariance against the small batches the oracle-parity tests cover), and
sharded == unsharded.  configs[2] pose batch 128 @1080p, configs[3] ball batch 256, configs[4] all four on 4K
frames batch 64 (one shard of the 8-GPU split)."""
import numpy as np
import pytest
import torch

from oracle import weights as OW
from padel_analytics_b200 import synth
from padel_analytics_b200.engine.yolo_engine import YoloEngine
from padel_analytics_b200.trackers import BallTracker, KeypointsTracker, PlayerKeypointsTracker, PlayerTracker
from padel_analytics_b200.trackers import sv_compat as sv
from padel_analytics_b200.trackers.runner import FusedPass, ball_shard_frames, shard_range

pytestmark = pytest.mark.gpu


def _rows(results):
    return [(r.boxes.data.numpy().copy(), None if r.keypoints is None else r.keypoints.data.numpy().copy())
            for r in results]


def _same(a, b):
    return all(np.array_equal(x[0], y[0]) and (x[1] is None or np.array_equal(x[1], y[1])) for x, y in zip(a, b))


def test_pose_batch128_is_batch_invariant():
    """configs[2]: PlayerKeypointsTracker (YOLOv8-pose 13 kpts @1280), 1080p, batch_size=128."""
    ck = OW.make_yolo("pose13")  # dense head: ~130 detections per frame stress decode + NMS
    frames = synth.make_frames(128, 1080, 1920, start=200, device="cuda")
    big = YoloEngine(ck, max_batch=128)
    got = _rows(big.predict_frames(frames, "pil_square", conf=0.25, iou=0.7, imgsz=1280, classes=[0]))
    assert len(got) == 128 and sum(len(g[0]) for g in got) > 128
    small = YoloEngine(ck, max_batch=4)
    for lo in (0, 62, 124):
        ref = _rows(small.predict_frames(frames[lo:lo + 4], "pil_square", conf=0.25, iou=0.7, imgsz=1280, classes=[0]))
        assert _same(ref, got[lo:lo + 4]), f"frames {lo}..{lo+3} differ between batch 4 and batch 128"


def test_ball_batch256_is_batch_invariant_and_shards():
    """configs[3]: BallTracker only, 1080p, batch_size=256, frames sharded across 2 GPUs (shards run back to back)."""
    T = 300
    ck = OW.make_tracknet()
    frames = synth.make_frames(T, 1080, 1920, start=50, device="cuda")
    med = synth.make_median(1080, 1920).numpy()
    vi = sv.VideoInfo(width=1920, height=1080, fps=30.0, total_frames=T)
    batches = lambda lo, hi, B: (frames[i:min(i + B, hi)] for i in range(lo, hi, B))
    big = BallTracker(ck, None, batch_size=256, median=med)
    big.video_info_post_init(vi)
    full = big.track_xyv(batches(0, T, 256), T)
    assert sorted(full) == list(range(T))
    small = BallTracker(ck, None, batch_size=8, median=med)
    small.video_info_post_init(vi)
    ref = small.track_xyv(batches(0, T, 8), T)
    assert ref == full, "ball results depend on the batch size"
    merged = {}
    for r in range(2):
        lo, hi = shard_range(T, r, 2)
        flo, fhi = ball_shard_frames(T, lo, hi)
        merged.update(big.track_xyv(batches(flo, fhi, 256), T, first_frame=flo, emit_range=(lo, hi)))
    assert merged == full, "2-way sharded ball tracking differs from the unsharded pass"
    assert sum(v[2] for v in full.values()) > 0, "vacuous: the ball was never visible"


def test_all_four_4k_batch64_fused_pass():
    """configs[4]: all four trackers on 4K frames, batch_size=64 (the per-GPU shard of the 8-GPU split)."""
    H, W, B = 2160, 3840, 64
    frames = synth.make_frames(B, H, W, start=7, device="cuda")
    med = synth.make_median(H, W).numpy()
    vi = sv.VideoInfo(width=W, height=H, fps=30.0, total_frames=B)
    poly = sv.PolygonZone(np.array([[0, 0], [W - 1, 0], [W - 1, H - 1], [0, H - 1]]), frame_resolution_wh=(W, H))
    cks = {"detect": OW.make_yolo("detect", cls_mean=-5.0), "pose13": OW.make_yolo("pose13", cls_mean=-5.7),
           "court12": OW.make_yolo("court12"), "tracknet": OW.make_tracknet()}

    def build(bs):
        tr = {"players": PlayerTracker(cks["detect"], poly, batch_size=bs),
              "pose": PlayerKeypointsTracker(cks["pose13"], 1280, batch_size=bs, load_path=None, save_path=None),
              "court": KeypointsTracker(cks["court12"], batch_size=bs, model_type="yolo"),
              "ball": BallTracker(cks["tracknet"], None, batch_size=bs, median=med)}
        for t in tr.values():
            t.video_info_post_init(vi)
        return tr

    def run(tr, bs):
        out = {"players": [], "pose": [], "court": []}
        ball = {}
        for res in FusedPass(tr, (H, W), bs, total_frames=B).run(frames[i:i + bs] for i in range(0, B, bs)):
            for k in out:
                out[k] += [o.serialize() for o in res[k]]
            ball.update(res["ball"])
        return out, ball

    big_out, big_ball = run(build(64), 64)
    small_out, small_ball = run(build(8), 8)
    assert big_ball == small_ball and sorted(big_ball) == list(range(B))
    for k in big_out:
        assert len(big_out[k]) == B
        assert big_out[k] == small_out[k], f"{k}: batch 64 and batch 8 disagree on 4K frames"
