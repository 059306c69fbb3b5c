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
from ..engine.yolo_engine import ResultBlock
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
                 first_frame: int = 0, emit_range: Optional[tuple[int, int]] = None, streams: Optional[int] = None,
                 raw: bool = False, median=None):
        """raw=True: the YOLO trackers' entries are the engines' raw per-frame Results (no polygon filter / ByteTrack /
        result objects) -- what a shard hands to rank 0, where the order-dependent host stages run once over the
        ordered gather.  median: background for the ball tracker (defaults to BallTracker.median)."""
        self.trackers = trackers
        self.raw = raw
        self.mode = int(os.environ.get("PADEL_B200_STREAMS", "1")) if streams is None else streams
        # side streams: the YOLO chains at high priority (their CTAs are placed first whenever SMs free up), the ball
        # tracker (mode 2 only) at normal priority
        self.side = {name: torch.cuda.Stream(priority=0 if isinstance(t, BallTracker) else -1)
                     for name, t in trackers.items()}
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
                t.stream_begin(self.hw, total_frames, first_frame, emit_range, median=median)

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
            elif self.raw:
                out[name] = fin()
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
            total_frames: Optional[int] = None, fused: Optional[bool] = None) -> dict[str, float]:
        """The runner pass (runner.py:185-234).  `frame_source(lo, hi)` yields frames lo..hi-1 (defaults to decoding
        `video_path`).

        fused (default: on when two or more trackers still need inference): ONE pass over the frames feeds every
        such tracker from a single decode + upload per batch (`FusedPass`; the reference decodes and uploads once per
        tracker, runner.py:215-220).  fused=False is the reference's plain loop, one full pass per tracker.
        Trackers with cached predictions (runner.py:187-191) or a fixed keypoints detection never enter the fused set.

        Under torch.distributed (world_size > 1) each rank processes its contiguous shard, fixed-capacity detection
        records are all-gathered, and rank 0 runs the order-dependent host stages (polygon filter, ByteTrack ids,
        InpaintNet, result objects) over the ordered frames."""
        import gc

        # The pass creates hundreds of small result objects per frame and keeps them all (the reference's results API);
        # none of them form cycles, so the cyclic collector only costs time (measured: +30 % on a 4096-frame job as
        # its generations fill up).  Collection is suspended for the duration of the pass.
        gc_was_on = gc.isenabled()
        gc.disable()
        try:
            return self._run(frame_source, total_frames, fused)
        finally:
            if gc_was_on:
                gc.enable()

    def _run(self, frame_source, total_frames, fused) -> dict[str, float]:
        import torch.distributed as dist

        src = frame_source or self._frames
        total = total_frames if total_frames is not None else self.total_frames
        dist_on = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist_on else (0, 1)
        lo, hi = shard_range(total, rank, world)
        todo = {n: t for n, t in self.trackers.items() if len(t) == 0}  # cached predictions: skipped (runner.py:187-191)
        model = {n: t for n, t in todo.items() if isinstance(t, (PlayerTracker, PlayerKeypointsTracker, BallTracker))
                 or (isinstance(t, KeypointsTracker) and t.fixed_keypoints_detection is None and t.model is not None)}
        if fused is None:
            fused = len(model) >= 2 and torch.cuda.is_available()
        if fused and model:
            self._run_fused(model, src, total, lo, hi, rank, world, dist_on)
            todo = {n: t for n, t in todo.items() if n not in model}
        for name, tracker in todo.items():
            tracker.to(tracker.DEVICE)
            t0 = timeit.default_timer()
            if isinstance(tracker, BallTracker):
                flo, fhi = ball_shard_frames(total, lo, hi)
                median = self._ball_median(tracker, src, total, rank, dist_on)
                part = tracker.track_xyv(src(flo, fhi), total, first_frame=flo, emit_range=(lo, hi), median=median)
                part = _ball_records(part, lo, hi)
            elif isinstance(tracker, (PlayerTracker, PlayerKeypointsTracker, KeypointsTracker)) and \
                    getattr(tracker, "fixed_keypoints_detection", None) is None:
                part, hw = [], None
                for sample in sampler(src(lo, hi), tracker.batch_size):
                    hw = sample[0].shape[:2]
                    part += tracker.detect_sample(sample)
                part = (_yolo_records(part, tracker), hw)
            else:
                part = list(tracker.predict_and_update(src(lo, hi), total_frames=hi - lo).predictions)
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            self._gather_and_assemble(tracker, part, total, rank, world, dist_on)
            self.timings[name] = timeit.default_timer() - t0
            tracker.to("cpu")
            if rank == 0:
                tracker.save_predictions()
        return self.timings

    # ---- fused single pass -------------------------------------------------------------------------------------
    def _ball_median(self, tracker: BallTracker, src, total: int, rank: int, dist_on: bool):
        """Background median of the ball tracker.  The reference takes it from the first `median_max_sample_num`
        frames of the VIDEO (iterable.py:58-73): under sharding rank 0 computes it from those frames (device
        selection kernel) and broadcasts it, so every shard feeds TrackNet the same background."""
        import torch.distributed as dist

        if tracker.median is not None:
            return tracker.median
        from .ball_tracker import median_background

        med = None
        if rank == 0:
            m = min(total, tracker.median_max_sample_num)
            frames = list(src(0, m))
            if frames and isinstance(frames[0], torch.Tensor) and frames[0].dim() == 4:  # batched source
                frames = torch.cat([f.to("cuda") for f in frames])[:m]
            med = median_background(frames)
        if dist_on:
            dev = _comm_device()
            shape = torch.zeros(3, dtype=torch.int64, device=dev)
            if rank == 0:
                shape = torch.tensor(med.shape, dtype=torch.int64, device=dev)
            dist.broadcast(shape, src=0)
            buf = torch.from_numpy(med).to(dev) if rank == 0 else \
                torch.empty(tuple(int(v) for v in shape.tolist()), dtype=torch.uint8, device=dev)
            dist.broadcast(buf, src=0)
            med = buf.cpu().numpy()
        return med

    def _run_fused(self, model: dict, src, total: int, lo: int, hi: int, rank: int, world: int, dist_on: bool):
        """One pass over this rank's frames for every tracker in `model`.  A ball shard needs 7 frames of history and
        7 of look-ahead (`ball_shard_frames`); the YOLO results of those halo frames are simply dropped."""
        ball = next((t for t in model.values() if isinstance(t, BallTracker)), None)
        flo, fhi = ball_shard_frames(total, lo, hi) if ball is not None else (lo, hi)
        B = min(t.batch_size for t in model.values())  # every engine is sized for its own tracker's batch_size
        for t in model.values():
            t.to(t.DEVICE)
        t0 = timeit.default_timer()
        median = self._ball_median(ball, src, total, rank, dist_on) if ball is not None else None
        t_med = timeit.default_timer() - t0
        it = iter(src(flo, fhi))
        first = next(it, None)
        parts = {n: [] for n in model}
        hw = None
        if first is not None:
            # a frame source may yield single HWC frames (video decode) or ready uint8 (n,H,W,3) batches, n <= B
            # (pinned host or device tensors: no per-frame host copy)
            batched = isinstance(first, torch.Tensor) and first.dim() == 4
            hw = tuple(first.shape[1:3]) if batched else tuple(first.shape[:2])
            single = not dist_on
            fp = FusedPass(model, hw, B, total_frames=total, first_frame=flo, emit_range=(lo, hi), raw=not single,
                           median=median)
            pinned = None if batched else [torch.empty((B,) + hw + (3,), dtype=torch.uint8).pin_memory() for _ in range(3)]

            def batches():
                import itertools

                if batched:
                    yield from itertools.chain([first], it)
                    return
                chunk_it = sampler(itertools.chain([first], it), B)
                for i, chunk in enumerate(chunk_it):
                    buf = pinned[i % 3]
                    for j, f in enumerate(chunk):
                        buf[j].copy_(torch.from_numpy(np.ascontiguousarray(f)))
                    yield buf[:len(chunk)]

            pos = flo
            for out in fp.run(batches()):
                n = None
                for name, t in model.items():
                    if isinstance(t, BallTracker):
                        parts[name].append(out[name])
                    else:
                        res = out[name]
                        n = len(res)
                        a, b = max(lo - pos, 0), min(hi - pos, n)  # halo frames of a ball shard are dropped
                        if b > a:
                            if isinstance(res, ResultBlock):
                                parts[name].append(res[a:b])
                            else:
                                parts[name] += res[a:b]
                pos += n if n is not None else B
        torch.cuda.synchronize()
        t_pass = timeit.default_timer() - t0
        t1 = timeit.default_timer()
        for name, t in model.items():
            if isinstance(t, BallTracker):
                xyv = {}
                for d in parts[name]:
                    xyv.update(d)
                part = _ball_records(xyv, lo, hi)
            elif not dist_on or not all(isinstance(p, ResultBlock) for p in parts[name]):
                part = parts[name]  # already post-processed (or fixed) objects, in frame order
            else:
                part = (_yolo_records(parts[name], t), hw)
            self._gather_and_assemble(t, part, total, rank, world, dist_on)
            self.timings[name] = t_pass  # one shared pass: per-tracker times are not separable
            t.to("cpu")
            if rank == 0:
                t.save_predictions()
        self.timings["_fused_pass"] = t_pass
        self.timings["_median"] = t_med
        self.timings["_gather_assemble"] = timeit.default_timer() - t1

    def _gather_and_assemble(self, tracker: Tracker, part, total: int, rank: int, world: int, dist_on: bool) -> None:
        if not dist_on:
            if rank == 0:
                self._assemble(tracker, [part], total)
            return
        import torch.distributed as dist

        if isinstance(part, tuple) and isinstance(part[0], YoloRecords):
            gathered = [(r, part[1]) for r in _all_gather_yolo(part[0], total, rank, world)]
        elif isinstance(part, BallRecords):
            gathered = _all_gather_ball(part, total, rank, world)
        else:  # cached / fixed objects: ragged Python lists, gathered as objects
            gathered = [None] * world if rank == 0 else None
            dist.gather_object(part, gathered, dst=0)
        if rank == 0:
            self._assemble(tracker, gathered, total)

    @staticmethod
    def _assemble(tracker: Tracker, parts: list, total: int) -> None:
        if isinstance(tracker, BallTracker):
            xyv = {}
            for p in parts:
                xyv.update(p.to_dict() if isinstance(p, BallRecords) else p)
            xyv = tracker.inpaint_xyv(xyv, total)  # whole-trajectory stage: after the shards are merged
            tracker.results.predictions = [
                Ball(frame=n, xy=(xyv[n][0], xyv[n][1]), visibility=xyv[n][2]) if n in xyv
                else Ball(frame=n, xy=(0.0, 0.0), visibility=0) for n in range(total)]
        elif parts and isinstance(parts[0], tuple):
            hw = next((h for _, h in parts if h), None)
            if all(isinstance(res, YoloRecords) for res, _ in parts):  # dense records: stay dense
                results = ResultBlock.concat([res.to_results(tracker) for res, _ in parts])
            else:
                results = []
                for res, _ in parts:
                    results += list(res.to_results(tracker)) if isinstance(res, YoloRecords) else list(res)
            if isinstance(tracker, PlayerTracker):
                tracker.results.predictions = tracker.postprocess(results)  # ordered => ByteTrack ids are consistent
            else:
                tracker.results.predictions = tracker.postprocess(results, hw)
        else:
            tracker.results.predictions = [o for p in parts for o in p]


# ---- fixed-capacity records for the gather ---------------------------------------------------------
class YoloRecords:
    """Detections of consecutive frames as one dense float32 block (frames, cap, 6 + K*D) + int32 counts: rows are
    [x1, y1, x2, y2, conf, cls, keypoints...] exactly as the engine's ResultBlock holds them."""

    def __init__(self, rows: torch.Tensor, counts: torch.Tensor, kpt_shape):
        self.rows, self.counts, self.kpt_shape = rows, counts, kpt_shape

    def to_results(self, tracker) -> ResultBlock:
        return ResultBlock(self.rows.numpy(), self.counts.numpy(), self.kpt_shape, tracker.model.names, None)


class BallRecords:
    """(x, y, visibility, present) int32 per frame of a contiguous range starting at `first`."""

    def __init__(self, first: int, data: torch.Tensor):
        self.first, self.data = first, data

    def to_dict(self) -> dict:
        return {self.first + i: (int(x), int(y), int(v)) for i, (x, y, v, p) in enumerate(self.data.tolist()) if p}


def _yolo_records(blocks: list, tracker) -> YoloRecords:
    """This rank's frames (a list of per-batch ResultBlocks, or of Results) as one padded block."""
    kpt_shape = tracker.model.kpt_shape
    if blocks and not isinstance(blocks[0], ResultBlock):  # plain Results: one frame each
        rowlen = 6 + (kpt_shape[0] * kpt_shape[1] if kpt_shape else 0)
        one = []
        for r in blocks:
            n = len(r.boxes)
            rows = np.zeros((1, max(n, 1), rowlen), dtype=np.float32)
            rows[0, :n, :6] = r.boxes.data.numpy()
            if kpt_shape and n:
                rows[0, :n, 6:] = r.keypoints.data.numpy().reshape(n, rowlen - 6)
            one.append(ResultBlock(rows, np.array([n], dtype=np.int32), kpt_shape, r.names, r.orig_shape))
        blocks = one
    rowlen = 6 + (kpt_shape[0] * kpt_shape[1] if kpt_shape else 0)
    empty = ResultBlock(np.zeros((0, 1, rowlen), np.float32), np.zeros((0,), np.int32), kpt_shape, None, None)
    blk = ResultBlock.concat(blocks, like=empty)
    return YoloRecords(torch.from_numpy(blk.rows), torch.from_numpy(blk.counts), kpt_shape)


def _ball_records(xyv: dict, lo: int, hi: int) -> BallRecords:
    data = np.zeros((hi - lo, 4), dtype=np.int32)
    own = [(n - lo, x, y, v) for n, (x, y, v) in xyv.items() if lo <= n < hi]
    if own:
        a = np.asarray(own, dtype=np.int64)
        data[a[:, 0], :3] = a[:, 1:]
        data[a[:, 0], 3] = 1
    return BallRecords(lo, torch.from_numpy(data))


def _comm_device() -> torch.device:
    import torch.distributed as dist

    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def _all_gather_yolo(rec: YoloRecords, total: int, rank: int, world: int) -> list[YoloRecords]:
    """all_gather of padded fixed-capacity blocks: capacity = the global maximum detections per frame (one MAX
    all-reduce), frames padded to the longest shard.  Every rank receives every block; rank 0 uses them."""
    import torch.distributed as dist

    dev = _comm_device()
    npad = max(shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world))
    cap = torch.tensor([rec.rows.shape[1]], dtype=torch.int64, device=dev)
    dist.all_reduce(cap, op=dist.ReduceOp.MAX)
    cap = int(cap.item())
    rowlen = rec.rows.shape[2]
    rows = torch.zeros((npad, cap, rowlen), dtype=torch.float32, device=dev)
    counts = torch.zeros((npad,), dtype=torch.int32, device=dev)
    n = rec.rows.shape[0]
    rows[:n, : rec.rows.shape[1]] = rec.rows.to(dev)
    counts[:n] = rec.counts.to(dev)
    all_rows = [torch.empty_like(rows) for _ in range(world)]
    all_counts = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(all_rows, rows)
    dist.all_gather(all_counts, counts)
    out = []
    for r in range(world):
        a, b = shard_range(total, r, world)
        out.append(YoloRecords(all_rows[r][: b - a].cpu(), all_counts[r][: b - a].cpu(), rec.kpt_shape))
    return out


def _all_gather_ball(rec: BallRecords, total: int, rank: int, world: int) -> list[BallRecords]:
    import torch.distributed as dist

    dev = _comm_device()
    npad = max(shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world))
    data = torch.zeros((npad, 4), dtype=torch.int32, device=dev)
    data[: rec.data.shape[0]] = rec.data.to(dev)
    parts = [torch.empty_like(data) for _ in range(world)]
    dist.all_gather(parts, data)
    out = []
    for r in range(world):
        a, b = shard_range(total, r, world)
        out.append(BallRecords(a, parts[r][: b - a].cpu()))
    return out
