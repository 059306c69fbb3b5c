"""Minimal local stand-ins for the `supervision` objects the reference trackers touch (supervision is a third-party
dependency of the reference, requirements.txt:8, absent from this image).  If the real package is importable it is
used instead.  Call sites mirrored: /root/reference/trackers/players_tracker/players_tracker.py:311,363-369,333 ;
main.py:64,108-119 ; runner.py:52,215-220.  ByteTrack ids are UNPINNED: this is a simplified
IoU tracker with ByteTrack's thresholds, not a bit-exact port.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Iterable, Optional

import numpy as np

try:  # pragma: no cover - not installed in the build image
    import types as _types

    import supervision as _sv  # type: ignore

    # a test harness may have planted a mock under this name: only a real module counts
    HAVE_SUPERVISION = isinstance(_sv, _types.ModuleType) and isinstance(getattr(_sv, "__version__", None), str)
except Exception:  # noqa: BLE001
    _sv = None
    HAVE_SUPERVISION = False


@dataclass
class VideoInfo:
    width: int
    height: int
    fps: float
    total_frames: Optional[int] = None

    @property
    def resolution_wh(self):
        return self.width, self.height

    @classmethod
    def from_video_path(cls, path: str) -> "VideoInfo":
        import cv2

        cap = cv2.VideoCapture(str(path))
        if not cap.isOpened():
            raise FileNotFoundError(path)
        info = cls(int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)),
                   cap.get(cv2.CAP_PROP_FPS), int(cap.get(cv2.CAP_PROP_FRAME_COUNT)))
        cap.release()
        return info


def get_video_frames_generator(source_path: str, stride: int = 1, start: int = 0, end: Optional[int] = None):
    import cv2

    cap = cv2.VideoCapture(str(source_path))
    if not cap.isOpened():
        raise FileNotFoundError(source_path)
    total = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    end = total if end is None else min(end, total)
    cap.set(cv2.CAP_PROP_POS_FRAMES, start)
    i = start
    while i < end:
        ok, frame = cap.read()
        if not ok:
            break
        if (i - start) % stride == 0:
            yield frame
        i += 1
    cap.release()


@dataclass
class Detections:
    xyxy: np.ndarray
    confidence: Optional[np.ndarray] = None
    class_id: Optional[np.ndarray] = None
    tracker_id: Optional[np.ndarray] = None
    data: dict = field(default_factory=dict)

    @classmethod
    def from_ultralytics(cls, result) -> "Detections":
        names = result.names
        cid = result.boxes.cls.cpu().numpy().astype(int)
        return cls(
            xyxy=result.boxes.xyxy.cpu().numpy().reshape(-1, 4),
            confidence=result.boxes.conf.cpu().numpy(),
            class_id=cid,
            tracker_id=result.boxes.id.int().cpu().numpy() if result.boxes.id is not None else None,
            data={"class_name": np.array([names[int(c)] for c in cid])},
        )

    @classmethod
    def empty(cls):
        return cls(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), np.zeros((0,), int))

    def __len__(self):
        return len(self.xyxy)

    def __getitem__(self, idx) -> "Detections":
        if isinstance(idx, (int, np.integer)):
            idx = [int(idx)]
        idx = np.asarray(idx)
        pick = lambda a: None if a is None else a[idx]
        return Detections(self.xyxy[idx], pick(self.confidence), pick(self.class_id), pick(self.tracker_id),
                          {k: v[idx] for k, v in self.data.items()})


class PolygonZone:
    """Bottom-centre-anchor-in-polygon test (supervision PolygonZone default triggering anchor)."""

    def __init__(self, polygon: np.ndarray, frame_resolution_wh: tuple[int, int] | None = None, **kw):
        import cv2

        self.polygon = np.asarray(polygon).astype(np.int32)
        if frame_resolution_wh is None:
            frame_resolution_wh = (int(self.polygon[:, 0].max()) + 2, int(self.polygon[:, 1].max()) + 2)
        w, h = frame_resolution_wh
        self.frame_resolution_wh = (w, h)
        self.mask = np.zeros((h + 1, w + 1), dtype=np.uint8)
        cv2.fillPoly(self.mask, [self.polygon], color=1)
        self.mask = self.mask.astype(bool)

    def trigger(self, detections: Detections) -> np.ndarray:
        if len(detections) == 0:
            return np.zeros((0,), dtype=bool)
        w, h = self.frame_resolution_wh
        x = np.ceil((detections.xyxy[:, 0] + detections.xyxy[:, 2]) / 2).astype(int)
        y = np.ceil(detections.xyxy[:, 3]).astype(int)
        x = np.clip(x, 0, w)
        y = np.clip(y, 0, h)
        return self.mask[y, x]


def _iou_matrix(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    if len(a) == 0 or len(b) == 0:
        return np.zeros((len(a), len(b)), np.float32)
    x1 = np.maximum(a[:, None, 0], b[None, :, 0])
    y1 = np.maximum(a[:, None, 1], b[None, :, 1])
    x2 = np.minimum(a[:, None, 2], b[None, :, 2])
    y2 = np.minimum(a[:, None, 3], b[None, :, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    aa = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    ab = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / (aa[:, None] + ab[None, :] - inter + 1e-9)


class ByteTrack:
    """Simplified ByteTrack: high/low score split, IoU association (Hungarian), constant-velocity box prediction,
    lost-track buffer, ids from a global counter starting at 1.  Stateful and sequential (runs on rank 0)."""

    def __init__(self, track_activation_threshold: float = 0.25, lost_track_buffer: int = 30,
                 minimum_matching_threshold: float = 0.8, frame_rate: float = 30, **kw):
        self.high = track_activation_threshold
        self.det_thresh = track_activation_threshold + 0.1
        self.match = minimum_matching_threshold
        self.max_lost = int(frame_rate / 30.0 * lost_track_buffer)
        self.reset()

    def reset(self):
        self.tracks = []  # dict(id, box, vel, lost, hits)
        self.next_id = 1
        self.frame = 0

    def _assign(self, tracks, boxes, thr):
        from scipy.optimize import linear_sum_assignment

        if not tracks or len(boxes) == 0:
            return [], list(range(len(tracks))), list(range(len(boxes)))
        pred = np.stack([t["box"] + t["vel"] for t in tracks])
        cost = 1.0 - _iou_matrix(pred, boxes)
        r, c = linear_sum_assignment(cost)
        pairs = [(i, j) for i, j in zip(r, c) if cost[i, j] <= thr]
        mi, mj = {i for i, _ in pairs}, {j for _, j in pairs}
        return pairs, [i for i in range(len(tracks)) if i not in mi], [j for j in range(len(boxes)) if j not in mj]

    def update_with_detections(self, detections: Detections) -> Detections:
        self.frame += 1
        boxes = detections.xyxy.astype(np.float32)
        conf = detections.confidence if detections.confidence is not None else np.ones(len(boxes), np.float32)
        hi = np.where(conf > self.high)[0]
        lo = np.where((conf > 0.1) & (conf <= self.high))[0]
        ids = np.full(len(boxes), -1, dtype=int)
        pairs, un_t, un_d = self._assign(self.tracks, boxes[hi], self.match)
        for ti, dj in pairs:
            self._hit(self.tracks[ti], boxes[hi[dj]])
            ids[hi[dj]] = self.tracks[ti]["id"]
        rem = [self.tracks[i] for i in un_t if self.tracks[i]["lost"] == 0]
        pairs2, _, _ = self._assign(rem, boxes[lo], 0.5)
        matched2 = set()
        for ti, dj in pairs2:
            self._hit(rem[ti], boxes[lo[dj]])
            ids[lo[dj]] = rem[ti]["id"]
            matched2.add(id(rem[ti]))
        for i in un_t:
            t = self.tracks[i]
            if id(t) not in matched2:
                t["lost"] += 1
        for dj in un_d:
            d = hi[dj]
            if conf[d] >= self.det_thresh:
                self.tracks.append(dict(id=self.next_id, box=boxes[d].copy(), vel=np.zeros(4, np.float32), lost=0,
                                        hits=1))
                if self.frame == 1:
                    ids[d] = self.next_id
                self.next_id += 1
        self.tracks = [t for t in self.tracks if t["lost"] <= self.max_lost]
        keep = ids >= 0
        out = detections[np.where(keep)[0]]
        out.tracker_id = ids[keep]
        return out

    @staticmethod
    def _hit(t, box):
        t["vel"] = 0.5 * t["vel"] + 0.5 * (box - t["box"])
        t["box"] = box.copy()
        t["lost"] = 0
        t["hits"] += 1


if HAVE_SUPERVISION:  # pragma: no cover
    VideoInfo = _sv.VideoInfo  # noqa: F811
    Detections = _sv.Detections  # noqa: F811
    PolygonZone = _sv.PolygonZone  # noqa: F811
    ByteTrack = _sv.ByteTrack  # noqa: F811
    get_video_frames_generator = _sv.get_video_frames_generator  # noqa: F811
