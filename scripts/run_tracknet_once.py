"""Run the TrackNet program (and optionally a YOLO program) a fixed number of times — target for ncu captures."""
import sys
import torch
sys.path.insert(0, ".")
from oracle import weights as OW
from padel_analytics_b200.engine.tracknet_engine import TrackNetEngine
from padel_analytics_b200.engine.yolo_engine import YoloEngine
from padel_analytics_b200 import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = TrackNetEngine(OW.make_tracknet()["model"], max_batch=B)
eng.x.copy_(torch.rand_like(eng.x.float()).half())
for _ in range(reps):
    eng.run_packed()
if len(sys.argv) > 3:
    y = YoloEngine(OW.make_yolo("pose13"), max_batch=B)
    fr = synth.make_frames(B, 1080, 1920, device="cuda")
    for _ in range(reps):
        y.predict_frames(fr, "pil_square", conf=0.25, iou=0.7, imgsz=1280, classes=[0])
torch.cuda.synchronize()
print("done")
