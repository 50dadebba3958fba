"""Per-kernel device time of a full disparity chain with the speckle filter on (for ncu --metrics gpu__time_duration)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ros_gpu_stereo_processor_b200 as m
from ros_gpu_stereo_processor_b200 import _capi as capi
from oracle import synth
import bench
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
c = bench.CONFIGS[name]
W, H, nd = c["W"], c["H"], c["nd"]
L, R = synth.synth_pair(W, H, nd, 1234)
proc = m.GpuStereoProcessor(0)
proc.setParams(numDisparities=nd, blockSize=c["block"], minDisparity=0, preFilterType=1, preFilterSize=9, preFilterCap=31,
               textureThreshold=10, uniquenessRatio=15, speckleWindowSize=100, speckleRange=4, disp12MaxDiff=int(os.environ.get("D12", "-1")))
proc.configureSlots(1, H, W)
io = capi.FrameIO(); io.want = capi.OUT_DISPARITY16; io.rectify = 0
out = np.empty((H, W), np.int16); io.disparity16 = out.ctypes.data
import time
for i in range(5):
    proc.processPairAsync(0, L.ctypes.data, R.ctypes.data, io); proc.waitSlot(0)
t0 = time.perf_counter()
for i in range(20):
    proc.processPairAsync(0, L.ctypes.data, R.ctypes.data, io); proc.waitSlot(0)
print("%s: %.1f us per frame (host clock, incl. copies), valid=%.2f" % (name, (time.perf_counter() - t0) / 20 * 1e6, float((out != -16).mean())))
