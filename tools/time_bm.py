"""Times the block matcher alone on one config (device-resident), for tuning.  usage: python tools/time_bm.py C4 [reps]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ros_gpu_stereo_processor_b200 as m
from ros_gpu_stereo_processor_b200 import _capi as capi
from oracle import synth
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
c = bench.CONFIGS[name]
W, H, nd = c["W"], c["H"], c["nd"]
L, R = synth.synth_pair(W, H, nd, 1234)
proc = m.GpuStereoProcessor(0)
proc.setParams(numDisparities=nd, blockSize=c["block"], minDisparity=0, preFilterType=1, preFilterSize=9, preFilterCap=31,
               textureThreshold=10, uniquenessRatio=int(os.environ.get("UNIQ", "15")), speckleWindowSize=0, speckleRange=0, disp12MaxDiff=-1)
proc.configureSlots(1, H, W)
io = capi.FrameIO(); io.want = capi.OUT_DISPARITY16; io.rectify = 0
out = np.empty((H, W), np.int16); io.disparity16 = out.ctypes.data
proc.enableTiming(True)
ts = []
for i in range(reps + 3):
    proc.processPairAsync(0, L.ctypes.data, R.ctypes.data, io); proc.waitSlot(0)
    t, ev = proc.lastBmTime(0)
    if i >= 3: ts.append(t)
ts = np.array(ts)
print("%s env=%s: bm %.1f us (min %.1f)  %.1f Gevals/s  valid=%.2f" % (name, {k: v for k, v in os.environ.items() if k.startswith("B200S_")},
      1e3 * ts.mean(), 1e3 * ts.min(), ev / ts.mean() / 1e6, float((out != -16).mean())))
