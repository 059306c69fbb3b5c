"""TrackingRunner: the per-tracker pass over a video (API of /root/reference/trackers/runner.py:37-236), plus the
multi-GPU sharded variant: one process per GPU, contiguous frame ranges, no per-batch collectives —
detections are gathered once per tracker and the sequential host stages (ByteTrack ids, JSON) run on rank 0.
The drawing / data-collection pass (runner.py:91-173) is outside the hot path and is not reproduced.
"""
from __future__ import annotations

import timeit
from typing import Callable, Iterable, Optional

import numpy as np
import torch

from . import sv_compat as sv
from .ball_tracker import Ball, BallTracker
from .keypoints_tracker import KeypointsTracker
from .players_keypoints_tracker import PlayerKeypointsTracker
from .players_tracker import PlayerTracker
from .tracker import Tracker, sampler


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous frame range of `rank`: [rank*N/R, (rank+1)*N/R)."""
    return rank * total // world, (rank + 1) * total // world


def ball_shard_frames(total: int, start: int, end: int) -> tuple[int, int]:
    """Frames a ball shard must read to emit frames [start,end): 7 windows of history are recomputed and a window
    spans 8 frames => [start-7, end+7) clipped to the video."""
    return max(0, start - 7), min(total, end + 7)


class TrackingRunner:
    def __init__(self, trackers: dict[str, Tracker] | list[Tracker], video_path: Optional[str] = None,
                 inference_path: Optional[str] = None, start: int = 0, end: Optional[int] = None,
                 collect_data: bool = False, video_info=None):
        if isinstance(trackers, dict):
            trackers = list(trackers.values())
        self.trackers = {str(t): t for t in trackers}
        self.video_path = video_path
        self.inference_path = inference_path
        self.start, self.end = start, end
        if video_info is None and video_path is not None:
            video_info = sv.VideoInfo.from_video_path(video_path)
        self.video_info = video_info
        if video_info is not None:
            total = video_info.total_frames
            self.total_frames = (total if end is None else min(end, total)) - start if total is not None else None
            for t in self.trackers.values():
                t.video_info_post_init(video_info)  # runner.py:61-62
        self.timings: dict[str, float] = {}

    def restart(self) -> None:
        for t in self.trackers.values():
            t.restart()

    def _frames(self, lo: int, hi: int) -> Iterable[np.ndarray]:
        return sv.get_video_frames_generator(self.video_path, start=self.start + lo, end=self.start + hi)

    def run(self, frame_source: Optional[Callable[[int, int], Iterable[np.ndarray]]] = None,
            total_frames: Optional[int] = None) -> dict[str, float]:
        """Sequential per-tracker pass (runner.py:185-234).  `frame_source(lo, hi)` yields frames lo..hi-1 (defaults to
        decoding `video_path`).  Under torch.distributed (world_size > 1) each rank processes its contiguous shard and
        rank 0 assembles the results; with a single process this is the reference's plain loop."""
        import torch.distributed as dist

        src = frame_source or self._frames
        total = total_frames if total_frames is not None else self.total_frames
        dist_on = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist_on else (0, 1)
        lo, hi = shard_range(total, rank, world)
        for name, tracker in self.trackers.items():
            if len(tracker) != 0:  # cached predictions were loaded (runner.py:187-191)
                continue
            tracker.to(tracker.DEVICE)
            t0 = timeit.default_timer()
            if isinstance(tracker, BallTracker):
                flo, fhi = ball_shard_frames(total, lo, hi)
                part = tracker.track_xyv(src(flo, fhi), total, first_frame=flo, emit_range=(lo, hi))
            elif isinstance(tracker, (PlayerTracker, PlayerKeypointsTracker, KeypointsTracker)) and \
                    getattr(tracker, "fixed_keypoints_detection", None) is None:
                part, hw = [], None
                for sample in sampler(src(lo, hi), tracker.batch_size):
                    hw = sample[0].shape[:2]
                    part += tracker.detect_sample(sample)
                part = (part, hw)
            else:
                part = list(tracker.predict_and_update(src(lo, hi), total_frames=hi - lo).predictions)
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            if dist_on:
                gathered = [None] * world if rank == 0 else None
                dist.gather_object(part, gathered, dst=0)
            else:
                gathered = [part]
            if rank == 0:
                self._assemble(tracker, gathered, total)
            self.timings[name] = timeit.default_timer() - t0
            tracker.to("cpu")
            if rank == 0:
                tracker.save_predictions()
        return self.timings

    @staticmethod
    def _assemble(tracker: Tracker, parts: list, total: int) -> None:
        if isinstance(tracker, BallTracker):
            xyv = {}
            for p in parts:
                xyv.update(p)
            tracker.results.predictions = [
                Ball(frame=n, xy=(xyv[n][0], xyv[n][1]), visibility=xyv[n][2]) if n in xyv
                else Ball(frame=n, xy=(0.0, 0.0), visibility=0) for n in range(total)]
        elif parts and isinstance(parts[0], tuple):
            results, hw = [], None
            for res, h in parts:
                results += res
                hw = hw or h
            if isinstance(tracker, PlayerTracker):
                tracker.results.predictions = tracker.postprocess(results)  # ordered => ByteTrack ids are consistent
            else:
                tracker.results.predictions = tracker.postprocess(results, hw)
        else:
            tracker.results.predictions = [o for p in parts for o in p]
