"""CPU oracle for the padel_analytics tracker hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import it; the product (padel_analytics_b200/) never does.

What it restates (plain PyTorch-CPU / numpy / the same cv2 + Pillow + torchvision calls the reference makes):
  * oracle.tracknet  — TrackNet forward, window assembly, temporal ensemble, heat-map -> (x, y, vis)
                       following /root/reference/trackers/ball_tracker/{models,iterable,predict,ball_tracker}.py
  * oracle.yolov8    — ultralytics YOLOv8 detect / pose model + predict() pipeline.  ultralytics is a third-party
                       dependency of the reference (requirements.txt:9, unpinned, snapshot Jan-2025 => 8.3.x) that is
                       NOT vendored under /root/reference and not installed in this image; restated from its
                       published architecture and anchored on the reference's call sites.
  * oracle.weights   — seeded synthetic checkpoints (no real weights exist offline).

Pinning: the reference ships no tests/golden vectors.  The ball path is pinned against the
reference's own code run in the build container (oracle/ref_harness.py -> tests/golden/*.npz).  The YOLO path has
no runnable reference here (ultralytics absent) => "parity unpinned" for YOLO arithmetic beyond the third-party
pieces that are present and called directly (cv2.resize, PIL resize, torchvision.ops.nms).
"""
