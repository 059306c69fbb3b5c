"""BallTracker on the B200 engine — API of /root/reference/trackers/ball_tracker/ball_tracker.py (Ball :139-206,
BallTracker :208-711).  The TrackNet stage (:373-523) runs fully on device through engine.BallPipeline.

Documented deviations from reference quirks:
  q6  without an InpaintNet the reference dies with KeyError 'Frame' (:675-680); here the Ball list is built from the
      TrackNet x/y/visibility lists directly.
  q2  when `median` is None the reference buffers the first `median_max_sample_num` frames, converts them BGR->RGB
      twice and restarts the sliding window at the buffer boundary; here the median is computed from the same frames
      but every frame is converted once and the window never restarts.
  InpaintNet (:525-673) is not on the B200 path yet: passing inpainting_model_path raises.
"""
from __future__ import annotations

import math
from pathlib import Path
from typing import Iterable, Optional, Type

import numpy as np
import torch

from ..engine.tracknet_engine import BallPipeline, TrackNetEngine, bbox_to_xyv
from .tracker import NoPredictSample, Object, Tracker


class Ball(Object):
    def __init__(self, frame: int, xy: tuple[float, float], visibility: int,
                 projection: Optional[tuple[int, int]] = None):
        super().__init__()
        self.frame = frame
        self.xy = xy
        self.visibility = visibility
        self.projection = projection

    @classmethod
    def from_json(cls, x: dict):
        return cls(**x)

    def serialize(self) -> dict:
        return {"frame": self.frame, "xy": self.xy, "visibility": self.visibility, "projection": self.projection}

    def asint(self):
        return tuple(int(v) for v in self.xy)

    def draw(self, frame: np.ndarray) -> np.ndarray:
        import cv2

        cv2.circle(frame, self.asint(), 6, (0, 255, 0), -1)
        return frame

    def draw_projection(self, frame: np.ndarray) -> np.ndarray:
        import cv2

        cv2.circle(frame, self.projection, 6, (255, 255, 0), -1)
        return frame


def median_background(frames_bgr: list[np.ndarray], device="cuda") -> np.ndarray:
    """np.median(frames_rgb, 0).astype('uint8') (iterable.py:61-81) computed on device: per-pixel sort over the
    frame axis; even counts average the two middle values and truncate like the float64 -> uint8 cast."""
    n = len(frames_bgr)
    H, W, _ = frames_bgr[0].shape
    out = torch.empty((H, W, 3), dtype=torch.uint8, device=device)
    rows = max(1, (256 << 20) // (n * W * 3))
    for r0 in range(0, H, rows):
        chunk = torch.from_numpy(np.stack([f[r0:r0 + rows] for f in frames_bgr])).to(device)
        s, _ = torch.sort(chunk, dim=0)
        med = (s[(n - 1) // 2].to(torch.int32) + s[n // 2].to(torch.int32)) // 2
        out[r0:r0 + rows] = med.to(torch.uint8)
    return out.flip(-1).cpu().numpy()  # BGR -> RGB


class BallTracker(Tracker):
    EVAL_MODE: str = "weight"
    TRAJECTORY_LENGTH: int = 8
    HEIGHT: int = 288
    WIDTH: int = 512
    SIGMA: float = 2.5
    IMG_FORMAT = "png"

    def __init__(self, tracking_model_path, inpainting_model_path, batch_size: int,
                 median_max_sample_num: int = 1800, median: Optional[np.ndarray] = None,
                 load_path: Optional[str | Path] = None, save_path: Optional[str | Path] = None):
        super().__init__(load_path=load_path, save_path=save_path)
        self.DELTA_T: float = 1 / math.sqrt(self.HEIGHT ** 2 + self.WIDTH ** 2)
        self.COOR_TH = self.DELTA_T * 50
        ckpt = tracking_model_path if isinstance(tracking_model_path, dict) else \
            torch.load(tracking_model_path, map_location="cpu", weights_only=False)
        self.tracknet_seq_len = ckpt["param_dict"]["seq_len"]
        assert self.tracknet_seq_len == self.TRAJECTORY_LENGTH  # ball_tracker.py:256
        self.bg_mode = ckpt["param_dict"]["bg_mode"]
        if self.bg_mode != "concat":
            raise NotImplementedError("only bg_mode='concat' (what predict_frames hard-codes, :403) is supported")
        self.tracknet = TrackNetEngine(ckpt["model"], max_batch=batch_size, height=self.HEIGHT, width=self.WIDTH)
        if inpainting_model_path:
            raise NotImplementedError("InpaintNet stage is not on the B200 path yet (SURVEY §8f item 1)")
        self.inpaintnet = None
        self.batch_size = batch_size
        self.median_max_sample_num = median_max_sample_num
        self.median = median
        self._pipe = None

    def video_info_post_init(self, video_info) -> "BallTracker":
        self.video_info = video_info
        return self

    def object(self) -> Type[Object]:
        return Ball

    def draw_kwargs(self) -> dict:
        return {}

    def __str__(self) -> str:
        return "ball_tracker"

    def restart(self) -> None:
        self.results.restart()

    def to(self, device: str) -> None:
        self.tracknet.to(device)

    def predict_sample(self, sample, **kwargs):
        raise NoPredictSample()

    def _pipeline(self, frame_hw, median_rgb) -> BallPipeline:
        if self._pipe is None or (self._pipe.Hs, self._pipe.Ws) != tuple(frame_hw):
            self._pipe = BallPipeline(self.tracknet, frame_hw, median_rgb)
        return self._pipe

    def track_xyv(self, frame_generator: Iterable[np.ndarray], total_frames: int, first_frame: int = 0,
                  emit_range: Optional[tuple[int, int]] = None):
        """TrackNet stage on device.  Frames from the generator are absolute frames first_frame, first_frame+1, ...
        Returns dict frame_index -> (x, y, vis) for the frames emitted (restricted to emit_range if given)."""
        import itertools

        it = iter(frame_generator)
        B = self.batch_size
        pending: list[np.ndarray] = []
        median = self.median
        if median is None:  # iterable.py:58-73
            for f in it:
                pending.append(f)
                if len(pending) == self.median_max_sample_num:
                    break
            if not pending:
                return {}
            median = median_background(pending)
        stream = itertools.chain(pending, it)
        first = next(stream, None)
        if first is None:
            return {}
        pipe = self._pipeline(tuple(first.shape[-3:-1]), median)
        pipe.reset(base=first_frame)
        w_scaler, h_scaler = self.video_info.width / self.WIDTH, self.video_info.height / self.HEIGHT  # :379-384
        out: dict[int, tuple[int, int, int]] = {}
        total_windows = total_frames - 7

        def push(frames):
            pipe.push_frames(frames if isinstance(frames, torch.Tensor) else torch.from_numpy(np.stack(frames)))
            while True:  # run every window that became computable
                nb = min(B, pipe.windows_ready(), total_windows - (pipe.base + pipe.n_windows))
                if nb <= 0:
                    return
                f0, bbox = pipe.run_windows(nb, total_frames)
                xs, ys, vs = bbox_to_xyv(bbox, (w_scaler, h_scaler))
                for i in range(len(xs)):
                    n = f0 + i
                    if emit_range is None or emit_range[0] <= n < emit_range[1]:
                        out[n] = (xs[i], ys[i], vs[i])

        if isinstance(first, torch.Tensor) and first.dim() == 4:
            # batched frame source: items are uint8 (n,H,W,3) tensors (pinned host or device), n <= batch_size
            push(first)
            for batch in stream:
                push(batch)
            return out
        chunk = [first]
        for f in stream:
            if len(chunk) == B:
                push(chunk)
                chunk = []
            chunk.append(f)
        if chunk:
            push(chunk)
        return out

    # ---- streaming interface used by the fused single-pass runner (one batch at a time, device work asynchronous) ----
    def stream_begin(self, frame_hw, total_frames: int, first_frame: int = 0, emit_range=None, median=None):
        median = self.median if median is None else median
        if median is None:
            raise ValueError("stream_begin needs a background median (pass median= to BallTracker or here)")
        pipe = self._pipeline(tuple(frame_hw), median)
        pipe.reset(base=first_frame)
        self._stream = dict(total=total_frames, emit=emit_range,
                            scaler=(self.video_info.width / self.WIDTH, self.video_info.height / self.HEIGHT))
        return pipe

    def stream_push_async(self, frames: torch.Tensor):
        """frames: uint8 (n,H,W,3) BGR tensor (device or pinned host), n <= batch_size.  Enqueues resize + every
        window that became computable; returns a callable that waits and yields {frame: (x, y, vis)}."""
        pipe, s = self._pipe, self._stream
        pipe.push_frames(frames)
        fins = []
        while True:
            nb = min(self.batch_size, pipe.windows_ready(), s["total"] - 7 - (pipe.base + pipe.n_windows))
            if nb <= 0:
                break
            if fins:  # only one launch may be in flight per pipeline: resolve the previous one first
                res = fins[-1]()
                fins[-1] = (lambda r: (lambda: r))(res)
            fins.append(pipe.run_windows_async(nb, s["total"]))

        def finish():
            out = {}
            for fin in fins:
                f0, bbox = fin()
                xs, ys, vs = bbox_to_xyv(bbox, s["scaler"])
                for i in range(len(xs)):
                    n = f0 + i
                    if s["emit"] is None or s["emit"][0] <= n < s["emit"][1]:
                        out[n] = (xs[i], ys[i], vs[i])
            return out

        return finish

    def predict_frames(self, frame_generator: Iterable[np.ndarray], total_frames: int, **kwargs) -> list[Ball]:
        xyv = self.track_xyv(frame_generator, total_frames)
        balls = []
        for n in range(total_frames):  # ball_tracker.py:675-698 (missing frames -> (0,0), visibility 0)
            if n in xyv:
                x, y, v = xyv[n]
                balls.append(Ball(frame=n, xy=(x, y), visibility=v))
            else:
                balls.append(Ball(frame=n, xy=(0.0, 0.0), visibility=0))
        return balls
