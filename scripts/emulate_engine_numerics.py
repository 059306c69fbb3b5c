"""CPU emulation of the engine's numerics on the oracle network: BN folded, weights rounded to fp16, every conv output
rounded to fp16 (fp32 accumulate), final head convs fp32 out -- to study, without a GPU, how a head construction
behaves under the parity protocol.  Usage: python scripts/emulate_engine_numerics.py [dfl]"""
import sys

import cv2
import torch
import torch.nn as nn

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import parity  # noqa: E402
from oracle import weights as OW  # noqa: E402
from oracle import yolov8 as OY  # noqa: E402
from padel_analytics_b200 import synth  # noqa: E402


def emulated(net):
    import copy

    net = copy.deepcopy(net)
    h = lambda t: t.half().float()
    for m in net.modules():
        if isinstance(m, OY.ConvBnAct):
            scale = m.bn.weight / torch.sqrt(m.bn.running_var + m.bn.eps)
            w = h(m.conv.weight * scale.view(-1, 1, 1, 1))
            b = m.bn.bias - m.bn.running_mean * scale
            m.forward = (lambda w, b, mm: (lambda x: h(torch.nn.functional.silu(
                torch.nn.functional.conv2d(x, w, b, mm.conv.stride, mm.conv.padding)))))(w, b, m)
    for mod in (net.model[22].cv2, net.model[22].cv3, getattr(net.model[22], "cv4", [])):
        for br in mod:
            br[2].weight.data = h(br[2].weight.data)
    for m in net.modules():
        if isinstance(m, OY.Bottleneck):
            m.forward = (lambda mm: (lambda x: (h(x + mm.cv2(mm.cv1(x))) if mm.add else mm.cv2(mm.cv1(x)))))(m)
    return net


if __name__ == "__main__":
    dfls = ["random"]
    B, H, W = 3, 1080, 1920
    fr = [f.numpy() for f in synth.make_frames(B, H, W, start=5)]
    for dfl in dfls:
        for kind in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["detect"]):
            ck = OW.make_yolo(kind)
            net = OW.load_yolo(ck)
            yolo = OY.YOLO(net)
            yolo.predict([cv2.cvtColor(f, cv2.COLOR_BGR2RGB) for f in fr], conf=0.5, iou=0.7, imgsz=640, classes=[0])
            x = yolo.last_preprocessed
            em = emulated(net)
            with torch.no_grad():
                pred = em(h := x.half().float())
            dets = OY.non_max_suppression(pred, 0.5, 0.7, [0], 300, net.nc)
            res = []
            for det in dets:
                det = det.clone()
                det[:, :4] = OY.scale_boxes(x.shape[2:], det[:, :4], (H, W))
                res.append(OY.Result(OY.Boxes(det[:, :6]), None, net.names, (H, W)))
            reps = parity.check_batch(net, x, res, 0.5, 0.7, [0], 300, (H, W), verbose=False)
            print(f"{kind}: sure {sum(r.n_sure for r in reps)} (ill-conditioned {sum(r.n_reg for r in reps)}) unmatched {sum(len(r.sure_unmatched) for r in reps)} "
                  f"extras {sum(len(r.extras) for r in reps)} min IoU {min(r.min_iou_sure for r in reps):.4f} "
                  f"max dconf {max(r.max_conf_err for r in reps):.4f}")
            for r in reps:
                print("   ", r.sure_unmatched[:5])
