"""Minimal local stand-ins for the `supervision` objects the reference trackers touch (supervision is a third-party
dependency of the reference, requirements.txt:8, absent from this image).  If the real package is importable it is
used instead.  Call sites mirrored: /root/reference/trackers/players_tracker/players_tracker.py:311,363-369,333 ;
main.py:64,108-119 ; runner.py:52,215-220.  ByteTrack follows the published algorithm (Kalman xyah filter, two-stage
association, unconfirmed-track handling, duplicate pruning) but could not be checked against supervision's own code:
track ids stay UNPINNED.
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
    """Pairwise IoU of xyxy boxes (supervision box_iou_batch)."""
    a = np.asarray(a, dtype=np.float64).reshape(-1, 4)
    b = np.asarray(b, dtype=np.float64).reshape(-1, 4)
    if len(a) == 0 or len(b) == 0:
        return np.zeros((len(a), len(b)), np.float64)
    x1 = np.maximum(a[:, None, 0], b[None, :, 0])
    y1 = np.maximum(a[:, None, 1], b[None, :, 1])
    x2 = np.minimum(a[:, None, 2], b[None, :, 2])
    y2 = np.minimum(a[:, None, 3], b[None, :, 3])
    inter = np.clip(x2 - x1, 0, None) * np.clip(y2 - y1, 0, None)
    aa = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    ab = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    union = aa[:, None] + ab[None, :] - inter
    return np.where(union > 0, inter / np.where(union > 0, union, 1.0), 0.0)


# ---- ByteTrack (Zhang et al., ECCV 2022) as packaged by supervision: restated from the published algorithm ---------
# Kalman filter: the 8-state (x, y, aspect, height + velocities) constant-velocity filter of DeepSORT / ByteTrack.
class _KalmanXYAH:
    """Batched over tracks: the per-frame cost of the tracker is a handful of (N,8,8) numpy operations instead of N
    Python-level filter updates (ByteTrack runs on rank 0 for every frame of every shard)."""
    W_POS, W_VEL = 1.0 / 20, 1.0 / 160

    def __init__(self):
        self.F = np.eye(8)
        for i in range(4):
            self.F[i, 4 + i] = 1.0

    def initiate(self, m):
        mean = np.r_[m, np.zeros(4)]
        h = m[3]
        std = np.array([2 * self.W_POS * h, 2 * self.W_POS * h, 1e-2, 2 * self.W_POS * h,
                        10 * self.W_VEL * h, 10 * self.W_VEL * h, 1e-5, 10 * self.W_VEL * h])
        return mean, np.diag(std * std)

    def predict(self, means, covs):
        """means (N,8), covs (N,8,8) -> predicted."""
        h = means[:, 3]
        std = np.stack([self.W_POS * h, self.W_POS * h, np.full_like(h, 1e-2), self.W_POS * h, self.W_VEL * h,
                        self.W_VEL * h, np.full_like(h, 1e-5), self.W_VEL * h], 1)
        Q = np.zeros_like(covs)
        idx = np.arange(8)
        Q[:, idx, idx] = std * std
        return means @ self.F.T, self.F @ covs @ self.F.T + Q

    def update(self, means, covs, z):
        """means (N,8), covs (N,8,8), measurements z (N,4) -> corrected (H selects the first four states)."""
        h = means[:, 3]
        std = np.stack([self.W_POS * h, self.W_POS * h, np.full_like(h, 1e-1), self.W_POS * h], 1)
        S = covs[:, :4, :4].copy()
        idx = np.arange(4)
        S[:, idx, idx] += std * std
        PHt = covs[:, :, :4]  # (N,8,4)
        K = np.linalg.solve(S, PHt.transpose(0, 2, 1)).transpose(0, 2, 1)  # S symmetric: K = P H^T S^-1
        innov = z - means[:, :4]
        new_means = means + np.einsum("nij,nj->ni", K, innov)
        new_covs = covs - K @ S @ K.transpose(0, 2, 1)
        return new_means, new_covs


_NEW, _TRACKED, _LOST, _REMOVED = 0, 1, 2, 3


class _STrack:
    __slots__ = ("tlbr0", "score", "class_id", "det_index", "mean", "cov", "is_activated", "state", "track_id",
                 "frame_id", "start_frame", "tracklet_len")

    def __init__(self, tlbr, score, class_id, det_index):
        self.tlbr0 = np.asarray(tlbr, dtype=np.float64)
        self.score, self.class_id, self.det_index = float(score), class_id, det_index
        self.mean = self.cov = None
        self.is_activated = False
        self.state = _NEW
        self.track_id = 0
        self.frame_id = self.start_frame = 0
        self.tracklet_len = 0

    @property
    def tlbr(self):
        return _tlbr_of([self])[0]

    def xyah0(self):
        b = self.tlbr0
        w, h = b[2] - b[0], b[3] - b[1]
        return np.array([b[0] + w / 2, b[1] + h / 2, w / h, h])

    def activate(self, kf, frame_id, new_id):
        self.track_id = new_id
        self.mean, self.cov = kf.initiate(self.xyah0())
        self.tracklet_len = 0
        self.state = _TRACKED
        if frame_id == 1:
            self.is_activated = True
        self.frame_id = self.start_frame = frame_id


def _tlbr_of(tracks) -> np.ndarray:
    """(n,4) xyxy of tracks (from the filter state) / fresh detections (their own box)."""
    out = np.empty((len(tracks), 4))
    for i, t in enumerate(tracks):
        if t.mean is None:
            out[i] = t.tlbr0
        else:
            x, y, a, h = t.mean[:4]
            w = a * h
            out[i] = (x - w / 2, y - h / 2, x + w / 2, y + h / 2)
    return out


def _xyah_of(dets) -> np.ndarray:
    b = np.stack([d.tlbr0 for d in dets])
    w, h = b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]
    return np.stack([b[:, 0] + w / 2, b[:, 1] + h / 2, w / h, h], 1)


def _linear_assignment(cost: np.ndarray, thresh: float):
    if cost.size == 0:
        return [], list(range(cost.shape[0])), list(range(cost.shape[1]))
    from scipy.optimize import linear_sum_assignment

    c = cost.copy()
    c[c > thresh] = thresh + 1e-4
    rows, cols = linear_sum_assignment(c)
    matches = [(int(r), int(q)) for r, q in zip(rows, cols) if c[r, q] <= thresh]
    mr, mc = {r for r, _ in matches}, {q for _, q in matches}
    return matches, [i for i in range(cost.shape[0]) if i not in mr], [j for j in range(cost.shape[1]) if j not in mc]


def _iou_distance(a, b):
    return 1.0 - _iou_matrix(_tlbr_of(a), _tlbr_of(b)) if (a and b) else np.zeros((len(a), len(b)), np.float64)


def _fuse_score(cost, dets):
    if cost.size == 0:
        return cost
    return 1.0 - (1.0 - cost) * np.array([d.score for d in dets])[None, :]


def _joint(a, b):
    seen, out = set(), []
    for t in list(a) + list(b):
        if t.track_id not in seen:
            seen.add(t.track_id)
            out.append(t)
    return out


def _sub(a, b):
    drop = {t.track_id for t in b}
    return [t for t in a if t.track_id not in drop]


class ByteTrackPy:
    """ByteTrack with supervision's constructor and `update_with_detections` contract (players_tracker.py:311,367-369).

    Per frame: detections are split into high (score > track_activation_threshold) and low (0.1 < score <= threshold)
    sets; all tracked + lost tracks are Kalman-predicted; 1st association high detections <-> pool on 1 - IoU fused with
    the detection score (Hungarian, minimum_matching_threshold); 2nd association the still-unmatched TRACKED tracks <->
    low detections (0.5); the rest of those become lost; unconfirmed tracks (one frame old) <-> leftover high detections
    (0.7), unmatched ones are removed; leftover detections with score >= threshold + 0.1 start new tracks, which are
    confirmed at their next match (immediately on the very first frame); lost tracks older than lost_track_buffer are
    removed; tracked/lost duplicates with IoU > 0.85 are pruned (the younger one).  Output = input detections that
    match an active track (IoU > 0.5, Hungarian), with `tracker_id` set; ids count from 1.

    Restated from the paper / public description (supervision is absent here), NOT validated against supervision's
    own code: ids remain UNPINNED.  Stateful and sequential: runs on rank 0 over frame-ordered detections."""

    def __init__(self, track_activation_threshold: float = 0.25, lost_track_buffer: int = 30,
                 minimum_matching_threshold: float = 0.8, frame_rate: float = 30, **kw):
        self.track_activation_threshold = track_activation_threshold
        self.minimum_matching_threshold = minimum_matching_threshold
        self.det_thresh = track_activation_threshold + 0.1
        self.max_time_lost = int(frame_rate / 30.0 * lost_track_buffer)
        self.kf = _KalmanXYAH()
        self.reset()

    def reset(self):
        self.frame_id = 0
        self.tracked, self.lost = [], []
        self._next_id = 0

    def _new_id(self):
        self._next_id += 1
        return self._next_id

    def _apply(self, tracks, dets, matches, activated, refind):
        """Kalman-correct the matched tracks with their detections (one batched update) and mark them tracked."""
        if not matches:
            return
        ts = [tracks[i] for i, _ in matches]
        ds = [dets[j] for _, j in matches]
        M, Cv = self.kf.update(np.stack([t.mean for t in ts]), np.stack([t.cov for t in ts]), _xyah_of(ds))
        for k, (t, d) in enumerate(zip(ts, ds)):
            t.mean, t.cov = M[k], Cv[k]
            if t.state == _TRACKED:
                t.tracklet_len += 1
                activated.append(t)
            else:  # re-activation of a lost track
                t.tracklet_len = 0
                refind.append(t)
            t.state = _TRACKED
            t.is_activated = True
            t.frame_id = self.frame_id
            t.score = d.score

    def _update(self, boxes, scores, class_ids):
        self.frame_id += 1
        activated, refind, lost_now, removed_now = [], [], [], []
        high = scores > self.track_activation_threshold
        second = (scores > 0.1) & (scores < self.track_activation_threshold)
        dets = [_STrack(boxes[i], scores[i], class_ids[i], i) for i in np.where(high)[0]]
        dets2 = [_STrack(boxes[i], scores[i], class_ids[i], i) for i in np.where(second)[0]]
        unconfirmed = [t for t in self.tracked if not t.is_activated]
        tracked = [t for t in self.tracked if t.is_activated]
        pool = _joint(tracked, self.lost)
        if pool:  # multi_predict: lost tracks keep their height (velocity of h zeroed)
            M = np.stack([t.mean for t in pool])
            M[[t.state != _TRACKED for t in pool], 7] = 0
            M, Cv = self.kf.predict(M, np.stack([t.cov for t in pool]))
            for i, t in enumerate(pool):
                t.mean, t.cov = M[i], Cv[i]
        cost = _fuse_score(_iou_distance(pool, dets), dets)
        matches, u_track, u_det = _linear_assignment(cost, self.minimum_matching_threshold)
        self._apply(pool, dets, matches, activated, refind)
        r_tracked = [pool[i] for i in u_track if pool[i].state == _TRACKED]
        matches, u_track2, _ = _linear_assignment(_iou_distance(r_tracked, dets2), 0.5)
        self._apply(r_tracked, dets2, matches, activated, refind)
        for it in u_track2:
            t = r_tracked[it]
            if t.state != _LOST:
                t.state = _LOST
                lost_now.append(t)
        rest = [dets[i] for i in u_det]
        cost = _fuse_score(_iou_distance(unconfirmed, rest), rest)
        matches, u_unc, u_det = _linear_assignment(cost, 0.7)
        self._apply(unconfirmed, rest, matches, activated, activated)
        for it in u_unc:
            unconfirmed[it].state = _REMOVED
            removed_now.append(unconfirmed[it])
        for i in u_det:
            d = rest[i]
            if d.score < self.det_thresh:
                continue
            d.activate(self.kf, self.frame_id, self._new_id())
            activated.append(d)
        for t in self.lost:
            if self.frame_id - t.frame_id > self.max_time_lost:
                t.state = _REMOVED
                removed_now.append(t)
        self.tracked = [t for t in self.tracked if t.state == _TRACKED]
        self.tracked = _joint(_joint(self.tracked, activated), refind)
        self.lost = _sub(self.lost, self.tracked)
        self.lost.extend(lost_now)
        self.lost = [t for t in self.lost if t.state != _REMOVED]  # removed tracks leave at once (no zombie frame)
        # duplicate pruning: of a tracked/lost pair with IoU > 0.85 the one with the shorter history goes
        if self.tracked and self.lost:
            pd = _iou_distance(self.tracked, self.lost)
            da, db = set(), set()
            for p, q in zip(*np.where(pd < 0.15)):
                tp = self.tracked[p].frame_id - self.tracked[p].start_frame
                tq = self.lost[q].frame_id - self.lost[q].start_frame
                (db if tp > tq else da).add(int(q) if tp > tq else int(p))
            self.tracked = [t for i, t in enumerate(self.tracked) if i not in da]
            self.lost = [t for i, t in enumerate(self.lost) if i not in db]
        return [t for t in self.tracked if t.is_activated]

    def update_with_detections(self, detections: Detections) -> Detections:
        n = len(detections)
        boxes = np.asarray(detections.xyxy, dtype=np.float64).reshape(-1, 4)
        scores = np.asarray(detections.confidence if detections.confidence is not None else np.ones(n), dtype=np.float64)
        cls = np.asarray(detections.class_id if detections.class_id is not None else np.zeros(n, int))
        tracks = self._update(boxes, scores, cls)
        ids = np.full(n, -1, dtype=int)
        if tracks and n:
            cost = 1.0 - _iou_matrix(boxes, _tlbr_of(tracks))
            matches, _, _ = _linear_assignment(cost, 0.5)
            for i_det, i_trk in matches:
                ids[i_det] = tracks[i_trk].track_id
        keep = np.where(ids != -1)[0]
        out = detections[keep]
        out.tracker_id = ids[keep]
        return out


class ByteTrack:
    """The product tracker: the same algorithm as ByteTrackPy in C++ (csrc/bytetrack.cu, `pb_bytetrack_*`), a few
    microseconds per frame instead of ~0.3 ms -- at thousands of frames per second the sequential rank-0 stage must
    not be the slowest part of the pass.  tests/test_host_cpu.py checks id-for-id equality with ByteTrackPy."""

    def __init__(self, track_activation_threshold: float = 0.25, lost_track_buffer: int = 30,
                 minimum_matching_threshold: float = 0.8, frame_rate: float = 30, **kw):
        from .. import _lib as L

        self._L = L
        self._h = L.lib().pb_bytetrack_create(float(track_activation_threshold), int(lost_track_buffer),
                                              float(minimum_matching_threshold), float(frame_rate))

    def __del__(self):
        try:
            if self._h:
                self._L.lib().pb_bytetrack_destroy(self._h)
                self._h = None
        except Exception:  # noqa: BLE001
            pass

    def reset(self):
        self._L.lib().pb_bytetrack_reset(self._h)

    def update_with_detections(self, detections: Detections) -> Detections:
        n = len(detections)
        boxes = np.ascontiguousarray(detections.xyxy, dtype=np.float32).reshape(-1, 4)
        scores = np.ascontiguousarray(detections.confidence if detections.confidence is not None else np.ones(n),
                                      dtype=np.float32)
        ids = np.empty(max(n, 1), dtype=np.int32)
        self._L.check(self._L.lib().pb_bytetrack_update(self._h, boxes.ctypes.data, scores.ctypes.data, n,
                                                        ids.ctypes.data))
        keep = np.where(ids[:n] != -1)[0]
        out = detections[keep]
        out.tracker_id = ids[:n][keep].astype(int)
        return out

    def update_many(self, xyxy: np.ndarray, scores: np.ndarray, counts: np.ndarray) -> np.ndarray:
        """`len(counts)` consecutive frames in one native call: xyxy (total, 4) / scores (total) float32 are the frames'
        detections back to back.  Returns the track id of every detection (-1 = no active track), exactly what
        per-frame update_with_detections calls would assign."""
        xyxy = np.ascontiguousarray(xyxy, dtype=np.float32).reshape(-1, 4)
        scores = np.ascontiguousarray(scores, dtype=np.float32)
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        assert int(counts.sum()) == len(xyxy) == len(scores)
        ids = np.empty(max(len(xyxy), 1), dtype=np.int32)
        self._L.check(self._L.lib().pb_bytetrack_update_many(self._h, xyxy.ctypes.data, scores.ctypes.data,
                                                             counts.ctypes.data, len(counts), ids.ctypes.data))
        return ids[:len(xyxy)]


if HAVE_SUPERVISION:  # pragma: no cover
    VideoInfo = _sv.VideoInfo  # noqa: F811
    Detections = _sv.Detections  # noqa: F811
    PolygonZone = _sv.PolygonZone  # noqa: F811
    ByteTrack = _sv.ByteTrack  # noqa: F811
    get_video_frames_generator = _sv.get_video_frames_generator  # noqa: F811
