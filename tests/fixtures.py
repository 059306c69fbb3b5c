"""Shared test fixtures (TEST INFRASTRUCTURE): the committed rally.mp4 crops and the seeded checkpoints used by the
reference-glue goldens (tests/golden/make_golden.py::yolo_glue_golden) and by the tests that replay them."""
from __future__ import annotations

from pathlib import Path

import numpy as np

from oracle import weights as OW

GOLDEN = Path(__file__).resolve().parent / "golden"


def rally_frames():
    import cv2

    d = GOLDEN / "rally"
    return [cv2.imread(str(p)) for p in sorted(d.glob("rally_f*_crop640x360.png"))]


def court_conf_for_single_detection(net, frame, imgsz=640):
    """A confidence threshold under which the oracle finds exactly ONE court detection in `frame` (the reference's
    `result.keypoints.xy.squeeze(0)` at keypoints_tracker.py:249 only works for exactly one): midway between the two
    best post-NMS scores."""
    import cv2
    from PIL import Image
    from oracle import yolov8 as OY

    im = Image.fromarray(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB)).resize((imgsz, imgsz))
    res = OY.YOLO(net).predict([im], conf=0.05, iou=0.7, imgsz=imgsz, max_det=12)[0]
    c = res.boxes.conf
    assert len(c) >= 2, "court oracle found fewer than two detections"
    return float((c[0] + c[1]) / 2)


def glue_ckpt(kind):
    """Seeded checkpoints whose heads are standardised on the first rally crop (the default calibration scene is
    synthetic; natural frames saturate those heads)."""
    return OW.make_yolo(kind, calib=OW.calib_from_frame(rally_frames()[0]),
                        cls_mean={"detect": -3.6, "pose13": -3.8, "court12": -3.4}[kind])
