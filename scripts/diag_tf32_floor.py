"""Noise floor of the REFERENCE's own GPU numerics: the oracle network (= what ultralytics runs) executed by PyTorch
eager on the B200 with cuDNN TF32 convolutions (torch's default, i.e. the reference's GPU path) and in strict fp32,
each compared with the CPU fp32 oracle under the same borderline-exclusion protocol (tests/parity.py) that the engine
is held to.  Also reports the engine itself on the same inputs.  Usage: python scripts/diag_tf32_floor.py [B]"""
import sys

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

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H, W = 1080, 1920
torch.set_num_threads(64)
frames = synth.make_frames(B, H, W, start=5)
fr_np = [f.numpy() for f in frames]


def results_from_pred(net, pred, x, ims_hw, conf, classes, max_det):
    dets = OY.non_max_suppression(pred, conf, 0.7, classes, max_det, net.nc)
    out = []
    for det in dets:
        det = det.clone()
        det[:, :4] = OY.scale_boxes(x.shape[2:], det[:, :4], ims_hw)
        kp = None
        if net.kpt_shape:
            K, D = net.kpt_shape
            k = OY.scale_coords(x.shape[2:], det[:, 6:].view(-1, K, D), ims_hw)
            if D == 3:
                k[..., :2][k[..., 2] < 0.5] = 0
            kp = OY.Keypoints(k)
        out.append(OY.Result(OY.Boxes(det[:, :6]), kp, net.names, ims_hw))
    return out


def summary(tag, reps):
    ns = sum(r.n_sure for r in reps)
    print(f"  {tag:34s} sure {ns:4d} (ill-cond. {sum(r.n_reg for r in reps):3d})  unmatched-sure {sum(len(r.sure_unmatched) for r in reps):3d}  extras "
          f"{sum(len(r.extras) for r in reps):3d}  min IoU {min(r.min_iou_sure for r in reps):.4f}  max dconf "
          f"{max(r.max_conf_err for r in reps):.4f}  max kpt {max(r.max_kpt_px for r in reps):.3f} px", flush=True)


for dfl in ("random",):
    for scale, kind, imgsz, prep in (("n", "detect", 640, "letterbox_q1"), ("n", "pose13", 1280, "pil_square"),
                                     ("n", "court12", 640, "pil_square"), ("m", "detect", 640, "letterbox_q1"),
                                     ("m", "court12", 640, "pil_square")):
        ck = OW.make_yolo(kind, scale=scale)
        net = OW.load_yolo(ck)
        conf = {"detect": 0.5, "pose13": 0.25, "court12": 0.5}[kind]
        classes = [0] if kind != "court12" else None
        max_det = 12 if kind == "court12" else 300
        yolo = OY.YOLO(net)
        if prep == "letterbox_q1":
            sample = [cv2.cvtColor(f, cv2.COLOR_BGR2RGB) for f in fr_np]
            img_hw, fs = (H, W), (1.0, 1.0)
        else:
            sample = [Image.fromarray(cv2.cvtColor(f, cv2.COLOR_BGR2RGB)).resize((imgsz, imgsz)) for f in fr_np]
            img_hw, fs = (imgsz, imgsz), (W / imgsz, H / imgsz)
        yolo.predict(sample, conf=conf, iou=0.7, imgsz=imgsz, classes=classes, max_det=max_det)
        x = yolo.last_preprocessed
        print(f"== dfl={dfl} {scale}/{kind}")
        gnet = OW.load_yolo(ck).cuda()
        for name, tf32 in (("reference GPU path (cuDNN TF32)", True), ("eager CUDA strict fp32", False)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            with torch.no_grad():
                pred = gnet(x.cuda()).cpu()
            res = results_from_pred(net, pred, x, img_hw, conf, classes, max_det)
            summary(name, parity.check_batch(net, x, res, conf, 0.7, classes, max_det, img_hw, fs, verbose=False))
        try:
            eng = YoloEngine(ck, max_batch=B)
            res = eng.predict_frames(frames, prep, conf=conf, iou=0.7, imgsz=imgsz, classes=classes, max_det=max_det)
            summary("engine (fp16 storage, fp32 acc)", parity.check_batch(net, x, res, conf, 0.7, classes, max_det,
                                                                          img_hw, fs, verbose=False))
            del eng
        except Exception as e:  # noqa: BLE001
            print("  engine FAILED", repr(e))
        del gnet
        torch.cuda.empty_cache()
