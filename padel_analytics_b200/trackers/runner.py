"""TrackingRunner: the per-tracker pass over a video (API of /root/reference/trackers/runner.py:37-236), plus the
multi-GPU sharded variant: one process per GPU, contiguous frame ranges, no per-batch collectives —
detections are gathered once per tracker and the sequential host stages (ByteTrack ids, JSON) run on rank 0.
The drawing / data-collection pass (runner.py:91-173) is outside the hot path and is not reproduced.
"""
from __future__ import annotations

import timeit
import os
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


class FusedPass:
    """One pass over the video feeding ALL trackers from a single upload per batch (the reference decodes and uploads
    the video once per tracker, runner.py:185-234).  Per batch: the next batch's host->device copy
    runs on a copy stream while this batch computes; the four trackers' device work is enqueued back to back without
    host synchronisation, and each tracker's host post-processing (ByteTrack, result objects) overlaps with the device
    work of the trackers behind it.

    `streams` (default env PADEL_B200_STREAMS, else 1): 0 = every tracker on the caller's stream; 1 = the YOLO trackers
    each on their own stream (their layers are small and latency-bound at batch 32 — many launch fewer CTAs than there
    are SMs — so three independent chains fill the machine), the ball tracker after them on the caller's stream;
    2 = all trackers concurrent.  The kernels and their inputs are the same in every mode, so are the results."""

    def __init__(self, trackers: dict[str, Tracker], frame_hw: tuple[int, int], batch_size: int, total_frames: int,
                 first_frame: int = 0, emit_range: Optional[tuple[int, int]] = None, streams: Optional[int] = None):
        self.trackers = trackers
        self.mode = int(os.environ.get("PADEL_B200_STREAMS", "1")) if streams is None else streams
        self.side = {name: torch.cuda.Stream() for name in trackers}
        self.hw = tuple(frame_hw)
        self.B = batch_size
        self.dev = torch.device("cuda")
        self.copy_stream = torch.cuda.Stream()
        self.staging = [torch.empty((batch_size,) + self.hw + (3,), dtype=torch.uint8, device=self.dev)
                        for _ in range(2)]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [None, None]  # per staging slot: event after the device work that read it
        for t in trackers.values():
            if isinstance(t, BallTracker):
                t.stream_begin(self.hw, total_frames, first_frame, emit_range)

    def _upload(self, frames, slot: int) -> torch.Tensor:
        if not isinstance(frames, torch.Tensor):
            frames = torch.from_numpy(np.stack(frames))
        n = frames.shape[0]
        if frames.device.type == "cuda":
            return frames
        with torch.cuda.stream(self.copy_stream):
            if self.consumed[slot] is not None:  # the batch that last used this slot must have been read
                self.copy_stream.wait_event(self.consumed[slot])
            self.staging[slot][:n].copy_(frames, non_blocking=True)
            self.ready[slot].record(self.copy_stream)
        return self.staging[slot][:n]

    def _process(self, fr: torch.Tensor) -> dict:
        return self._finish(self._launch(fr))

    def _launch(self, fr: torch.Tensor):
        """Enqueue the device work of every tracker for this batch (no host synchronisation)."""
        pending = []
        main = torch.cuda.current_stream()
        forked = []
        order = list(self.trackers.items())
        if self.mode == 1:  # YOLO chains first (concurrent), the ball tracker joins behind them
            order.sort(key=lambda kv: isinstance(kv[1], BallTracker))
        for name, t in order:  # enqueue everything first ...
            if getattr(t, "fixed_keypoints_detection", None) is not None:
                pending.append((name, t, None))
                continue
            is_ball = isinstance(t, BallTracker)
            own = self.mode == 2 or (self.mode == 1 and not is_ball)
            if own:
                s = self.side[name]
                s.wait_stream(main)
                forked.append(s)
                ctx = torch.cuda.stream(s)
            else:
                if self.mode == 1:
                    for s in forked:
                        main.wait_stream(s)
                ctx = torch.cuda.stream(main)
            with ctx:
                pending.append((name, t, t.stream_push_async(fr) if is_ball else t.detect_sample_async(fr)))
        for s in forked:  # the caller's stream (and the next upload into this staging slot) follows all of them
            main.wait_stream(s)
        pending.sort(key=lambda p: list(self.trackers).index(p[0]))
        return pending, fr.shape[0]

    def _finish(self, launched) -> dict:
        """Wait for each tracker's results in turn and run its host post-processing."""
        pending, nfr = launched
        out = {}
        for name, t, fin in pending:  # ... then finish in the same order
            if isinstance(t, BallTracker):
                out[name] = fin()
            elif fin is None:
                out[name] = [t.fixed_keypoints_detection] * nfr
            elif isinstance(t, PlayerTracker):
                out[name] = t.postprocess(fin())
            else:
                out[name] = t.postprocess(fin(), self.hw)
        return out

    def run(self, batches: Iterable):
        """batches: iterable of uint8 (n,H,W,3) BGR batches (pinned host tensors, device tensors or lists of frames),
        n <= batch_size.  Yields one {tracker name: results} dict per batch.

        One batch of look-ahead: batch i+1 is pulled from `batches` and enqueued before batch i's results are yielded.
        A batch (pinned host tensor: copied asynchronously into a staging slot; device tensor: read in place) must stay
        untouched until ITS OWN results have been yielded, i.e. a producer that reuses buffers needs at least two."""
        it = iter(batches)
        main = torch.cuda.current_stream()

        def start(frames, i):
            """upload (copy stream) + enqueue all device work of batch i; returns the launch record"""
            dev = self._upload(frames, i % 2)
            staged = dev.data_ptr() == self.staging[i % 2].data_ptr()
            if staged:
                main.wait_event(self.ready[i % 2])
            rec = self._launch(dev)
            if staged:
                ev = torch.cuda.Event()
                ev.record(main)
                self.consumed[i % 2] = ev
            return rec

        cur = next(it, None)
        if cur is None:
            return
        rec, i = start(cur, 0), 0
        while rec is not None:
            # one batch of look-ahead: batch i+1 is uploaded and fully enqueued before batch i's results are collected,
            # so the device never waits for the host post-processing (ByteTrack, result objects) of the batch before
            nxt = next(it, None)
            nrec = start(nxt, i + 1) if nxt is not None else None
            yield self._finish(rec)
            rec, i = nrec, i + 1


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
            xyv = tracker.inpaint_xyv(xyv, total)  # whole-trajectory stage: after the shards are merged
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
