"""Diagnostic: borderline-exclusion parity protocol (tests/parity.py) of the three YOLO engines vs the CPU oracle,
n and m scale, with full detail on every failure.  Usage: python scripts/diag_yolo_parity.py [scales] [B]"""
import sys
import time

import cv2
import torch
from PIL import Image

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import parity  # noqa: E402
from oracle import weights as OW  # noqa: E402
from oracle import yolov8 as OY  # noqa: E402
from padel_analytics_b200 import synth  # noqa: E402
from padel_analytics_b200.engine.yolo_engine import YoloEngine  # noqa: E402

scales = sys.argv[1].split(",") if len(sys.argv) > 1 else ["n", "m"]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
H, W = 1080, 1920
torch.set_num_threads(64)
frames = synth.make_frames(B, H, W, start=5)
fr_np = [f.numpy() for f in frames]
for scale in scales:
    for kind, imgsz, prep in (("detect", 640, "letterbox_q1"), ("pose13", 1280, "pil_square"),
                              ("court12", 640, "pil_square")):
        t0 = time.time()
        try:
            ck = OW.make_yolo(kind, scale=scale)
            net = OW.load_yolo(ck)
            conf = {"detect": 0.5, "pose13": 0.25, "court12": 0.5}[kind]
            classes = [0] if kind != "court12" else None
            max_det = 12 if kind == "court12" else 300
            eng = YoloEngine(ck, max_batch=B)
            res = eng.predict_frames(frames, prep, conf=conf, iou=0.7, imgsz=imgsz, classes=classes, max_det=max_det)
            yolo = OY.YOLO(net)
            if prep == "letterbox_q1":
                sample = [cv2.cvtColor(f, cv2.COLOR_BGR2RGB) for f in fr_np]
                img_hw, fs = (H, W), (1.0, 1.0)
            else:
                sample = [Image.fromarray(cv2.cvtColor(f, cv2.COLOR_BGR2RGB)).resize((imgsz, imgsz)) for f in fr_np]
                img_hw, fs = (imgsz, imgsz), (W / imgsz, H / imgsz)
            yolo.predict(sample, conf=conf, iou=0.7, imgsz=imgsz, classes=classes, max_det=max_det)
            x = yolo.last_preprocessed
            for eps_c, eps_i in ((0.02, 0.03), (0.04, 0.05)):
                reps = parity.check_batch(net, x, res, conf, 0.7, classes, max_det, img_hw, fs, eps_c, eps_i,
                                          tag=f"[{scale}/{kind} eps {eps_c}/{eps_i}]")
                ns = sum(r.n_sure for r in reps)
                print(f"== {scale}/{kind} eps {eps_c}/{eps_i}: sure {ns}, unmatched-sure "
                      f"{sum(len(r.sure_unmatched) for r in reps)}, extras {sum(len(r.extras) for r in reps)}, "
                      f"min IoU {min(r.min_iou_sure for r in reps):.4f}, max dconf {max(r.max_conf_err for r in reps):.4f}, "
                      f"max kpt {max(r.max_kpt_px for r in reps):.3f} px  ({time.time() - t0:.1f}s)", flush=True)
                for i, r in enumerate(reps):
                    for u in r.sure_unmatched:
                        print("   UNMATCHED img", i, u)
                    for u in r.extras:
                        print("   EXTRA img", i, u)
            del eng
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            import traceback

            traceback.print_exc()
            print(f"== {scale}/{kind}: FAILED {e!r}", flush=True)
