import sys, time, torch, numpy as np, collections
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
def ballstep():
    pipe.push_frames(fr); f0,bbox=pipe.run_windows(32,10**9); bbox_to_xyv(bbox,(3.75,3.75))
fns={"players":lambda: tr["players"].predict_sample(fr),"pose":lambda: tr["pose"].predict_sample(fr),"court":lambda: tr["court"].predict_sample(fr),"ball":ballstep,
     "players_detect_only":lambda: tr["players"].detect_sample(fr)}
def run(order, sync=False, reps=5):
    acc=collections.defaultdict(float)
    for it in range(reps+2):
        if it==2: acc.clear(); torch.cuda.synchronize(); t00=time.perf_counter()
        for k in order:
            t=time.perf_counter(); fns[k]()
            if sync: torch.cuda.synchronize()
            acc[k]+=time.perf_counter()-t
    torch.cuda.synchronize(); tot=(time.perf_counter()-t00)/reps*1e3
    print(order, "sync" if sync else "nosync", "total %.1f ms"%tot, {k: round(v/reps*1e3,2) for k,v in acc.items()})
run(["players","pose","court","ball"])
run(["players","pose","court","ball"], sync=True)
run(["ball","court","pose","players"])
run(["players_detect_only","pose","court","ball"])
run(["pose","court","ball"])
run(["pose"])
run(["players","pose"])
run(["court","pose"])
run(["ball","pose"])
print("---- per-iteration pose time, standard order, no sync")
for it in range(12):
    ts={}
    for k in ["players","pose","court","ball"]:
        t=time.perf_counter(); fns[k](); ts[k]=round((time.perf_counter()-t)*1e3,1)
    print(it, ts)
print("---- pose program only, event-timed, right after a ball step vs after idle")
st=list(tr["pose"].model._progs.values())[0]
for it in range(6):
    if it%2==0: ballstep()
    else: time.sleep(0.05)
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); st["prog"].run(); e1.record(); torch.cuda.synchronize()
    print("after ball" if it%2==0 else "after idle", round(e0.elapsed_time(e1),2),"ms")
