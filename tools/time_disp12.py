"""Per-frame device time of the disparity chain with disp12MaxDiff >= 0 (border bands + validate), CUDA events via bench-style batch timing."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ros_gpu_stereo_processor_b200 as m
from ros_gpu_stereo_processor_b200 import _capi as capi
from oracle import synth
import bench
c = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C4"]
W, H, nd = c["W"], c["H"], c["nd"]
L, R = synth.synth_pair(W, H, nd, 77)
for d12 in (-1, 0):
    proc = m.GpuStereoProcessor(0)
    proc.setParams(numDisparities=nd, blockSize=c["block"], preFilterType=1, preFilterCap=31, textureThreshold=10, uniquenessRatio=15, disp12MaxDiff=d12)
    proc.configureSlots(2, H, W)
    io = capi.FrameIO(); io.want = capi.OUT_DISPARITY16; io.rectify = 0; io.outputs_on_device = 1
    for i in range(40): proc.processPairAsync(i % 2, L.ctypes.data, R.ctypes.data, io)
    proc.waitSlot(0); proc.waitSlot(1)
    proc.batchBegin()
    for i in range(100): proc.processPairAsync(i % 2, L.ctypes.data, R.ctypes.data, io)
    ms = proc.batchEnd()
    print("%s disp12MaxDiff %d: %.1f us/frame (incl. 2 x %d-byte H2D)" % (sys.argv[1] if len(sys.argv) > 1 else "C4", d12, ms * 10, W * H))
    proc.close()
