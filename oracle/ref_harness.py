"""Import the UNMODIFIED reference package from /root/reference with stubs for the third-party packages this image
lacks, to pin the oracle and generate golden vectors (TEST INFRASTRUCTURE; only works where /root/reference exists,
i.e. the build container — never imported by GPU-box tests).

Stubs: supervision -> MagicMock, parse -> MagicMock, ultralytics -> module exposing the oracle's
YOLO stand-in so the reference's own predict_sample() code drives it.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path
from unittest.mock import MagicMock

REFERENCE = Path("/root/reference")


def available() -> bool:
    return (REFERENCE / "trackers" / "tracker.py").exists()


def import_reference():
    """Returns the reference `trackers` package (imported once)."""
    if not available():
        raise RuntimeError("/root/reference is not present on this machine")
    if "trackers" in sys.modules and getattr(sys.modules["trackers"], "__file__", "").startswith(str(REFERENCE)):
        return sys.modules["trackers"]
    from . import yolov8

    sys.modules.setdefault("supervision", MagicMock())
    sys.modules.setdefault("parse", MagicMock())
    ul = types.ModuleType("ultralytics")
    ul.YOLO = yolov8.YOLO
    sys.modules.setdefault("ultralytics", ul)
    if str(REFERENCE) not in sys.path:
        sys.path.insert(0, str(REFERENCE))
    import trackers  # noqa: E402

    return trackers
