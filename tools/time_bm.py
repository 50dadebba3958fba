"""Times the block matcher alone on one config (device-resident), for tuning.
usage: python tools/time_bm.py C4 [reps] [frames_per_launch]      (environment: B200S_VH_* planner overrides, UNIQ, DISP12, CAP)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ros_gpu_stereo_processor_b200 as m
from ros_gpu_stereo_processor_b200 import _capi as capi
from tools import synth
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
nf = int(sys.argv[3]) if len(sys.argv) > 3 else 1
c = bench.CONFIGS[name]
W, H, nd = c["W"], c["H"], c["nd"]
frames = [synth.synth_pair(W, H, nd, 1234 + i) for i in range(min(nf, 4))]
dL = [torch.from_numpy(np.ascontiguousarray(f[0])).cuda() for f in frames]
dR = [torch.from_numpy(np.ascontiguousarray(f[1])).cuda() for f in frames]
proc = m.GpuStereoProcessor(0)
proc.setParams(numDisparities=nd, blockSize=c["block"], minDisparity=0, preFilterType=1, preFilterSize=9, preFilterCap=int(os.environ.get("CAP", "31")),
               textureThreshold=10, uniquenessRatio=int(os.environ.get("UNIQ", "15")), speckleWindowSize=0, speckleRange=0,
               disp12MaxDiff=int(os.environ.get("DISP12", "-1")))
proc.configureSlots(1, H, W, nf)
ios = (capi.FrameIO * nf)()
for k in range(nf):
    ios[k].want, ios[k].rectify, ios[k].inputs_on_device, ios[k].outputs_on_device = capi.OUT_DISPARITY16, 0, 1, 1
batch = proc.makeBatch([dL[k % len(dL)].data_ptr() for k in range(nf)], [dR[k % len(dR)].data_ptr() for k in range(nf)], ios)
proc.syncParams()
proc.enableTiming(True)
ts = []
for i in range(reps + 3):
    proc.processBatchRaw(0, batch); proc.waitSlot(0)
    t, ev = proc.lastBmTime(0)
    if i >= 3: ts.append(t)
ts = np.array(ts) / nf
st = proc.lastStageTimes(0)
print("%s x%d env=%s: bm %.1f us/frame (min %.1f)  %.3f Tevals/s  stages/frame %s" % (
    name, nf, {k: v for k, v in os.environ.items() if k.startswith("B200S_")}, 1e3 * ts.mean(), 1e3 * ts.min(), ev / nf / ts.mean() / 1e9,
    {k: round(1e3 * v / nf, 1) for k, v in st.items()}))
