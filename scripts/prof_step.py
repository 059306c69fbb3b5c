import sys, time, torch, numpy as np
sys.path.insert(0,'.')
import bench
from oracle import weights as OW
from padel_analytics_b200 import synth
from padel_analytics_b200.engine.tracknet_engine import bbox_to_xyv
B=32; hw=(1080,1920)
ckpts={"detect":OW.make_yolo("detect",cls_mean=-5.0),"pose13":OW.make_yolo("pose13",cls_mean=-5.7),"court12":OW.make_yolo("court12")}; ckpts["tracknet"]=OW.make_tracknet()
tr,med=bench.build_trackers(B,hw,ckpts,'cuda')
fr=synth.make_frames(B,1080,1920,device='cuda')
pipe=tr["ball"]._pipeline(hw,med); pipe.reset(); pipe.push_frames(fr[:7])
def T(f,n=3):
    for _ in range(2): f()
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter()-t)/n*1e3
for k in ("players","pose","court"):
    t=tr[k]
    print(k,'predict_sample ms',T(lambda: t.predict_sample(fr)))
    print(k,'  detect_sample ms',T(lambda: t.detect_sample(fr)))
    m=t.model
    st=list(m._progs.values())[0]
    print(k,'  prog.run ms',T(lambda: st["prog"].run()))
    if k=="players":
        print(k,'  letterbox ms',T(lambda: m._letterbox(fr,640,(0,1,2))))
    else:
        sz=1280 if k=="pose" else 640
        print(k,'  pil_square ms',T(lambda: m._pil_square(fr,sz)))
    conf={"players":0.5,"pose":0.25,"court":0.5}[k]
    def det():
        return m._detect(st,B,conf,0.7,[0] if k!="court" else None,300 if k!="court" else 12)
    print(k,'  _detect(prog+decode+nms+d2h) ms',T(det))
    rows,counts=det()
    print(k,'  counts mean',counts.mean(), 'cand', st["cand_count"].float().mean().item())
    print(k,'  _results ms',T(lambda: m._results(rows,counts,B,(st["Hn"],st["Wn"]),hw)))
def ballstep():
    pipe.push_frames(fr); f0,bbox=pipe.run_windows(32,10**9); bbox_to_xyv(bbox,(3.75,3.75))
print('ball step ms',T(ballstep))
def bp(): pipe.push_frames(fr); pipe.n_frames_in-=32
print('ball push(resize) ms',T(bp))
print('tracknet prog ms',T(lambda: tr["ball"].tracknet.prog.run()))

import bench as _b
def fullstep():
    tr["players"].predict_sample(fr); tr["pose"].predict_sample(fr); tr["court"].predict_sample(fr); ballstep()
print('full step ms', T(fullstep, 5))
host=fr.cpu().pin_memory()
def fullstep_host():
    tr["players"].predict_sample(host); tr["pose"].predict_sample(host); tr["court"].predict_sample(host)
    pipe.push_frames(host); f0,bbox=pipe.run_windows(32,10**9); bbox_to_xyv(bbox,(3.75,3.75))
print('full step (pinned host frames) ms', T(fullstep_host, 5))
def h2d():
    tr["players"].model._upload(host)
print('one H2D upload of the batch ms', T(h2d, 5))

# ---- where does pose lose time when the trackers run back to back? ----
import collections
acc = collections.defaultdict(float)
def timed_call(name, f):
    torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize(); acc[name] += time.perf_counter() - t; return r
pm = tr["pose"].model
def pose_stages():
    frd = timed_call("pose.upload", lambda: pm._upload(fr))
    st, orig = timed_call("pose.pil_square", lambda: pm._pil_square(frd, 1280))
    rows, counts = timed_call("pose.detect", lambda: pm._detect(st, 32, 0.25, 0.7, [0], 300))
    res = timed_call("pose.results", lambda: pm._results(rows, counts, 32, (1280, 1280), orig))
    timed_call("pose.postprocess", lambda: tr["pose"].postprocess(res, (1080, 1920)))
for it in range(4):
    if it == 1: acc.clear()
    timed_call("players", lambda: tr["players"].predict_sample(fr))
    pose_stages()
    timed_call("court", lambda: tr["court"].predict_sample(fr))
    timed_call("ball", ballstep)
print({k: round(v / 3 * 1e3, 2) for k, v in acc.items()})
acc.clear()
for it in range(4):
    if it == 1: acc.clear()
    pose_stages()
print("pose alone", {k: round(v / 3 * 1e3, 2) for k, v in acc.items()})
