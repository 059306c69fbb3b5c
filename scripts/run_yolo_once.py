"""Fixed workload for `ncu --profile-from-start off`: one batch through one YOLO engine (pre-processing, the conv
program, decode, NMS) between cudaProfilerStart/Stop.  Usage: python scripts/run_yolo_once.py detect|pose13|court12 [B] [scale]"""
import sys

import torch

sys.path.insert(0, ".")
from oracle import weights as OW
from padel_analytics_b200 import synth
from padel_analytics_b200.engine.yolo_engine import YoloEngine

kind = sys.argv[1] if len(sys.argv) > 1 else "detect"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
scale = sys.argv[3] if len(sys.argv) > 3 else "n"
imgsz, prep, conf, classes, md = {"detect": (640, "letterbox_q1", 0.5, [0], 300), "pose13": (1280, "pil_square", 0.25, [0], 300),
                                  "court12": (640, "pil_square", 0.5, None, 12)}[kind]
eng = YoloEngine(OW.make_yolo(kind, scale=scale, cls_mean=-5.5), max_batch=B)
fr = synth.make_frames(B, 1080, 1920, device="cuda")
for _ in range(2):
    eng.predict_frames(fr, prep, conf=conf, iou=0.7, imgsz=imgsz, classes=classes, max_det=md)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.predict_frames(fr, prep, conf=conf, iou=0.7, imgsz=imgsz, classes=classes, max_det=md)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", kind, B, scale)
