import sys, time, numpy as np
sys.path.insert(0, ".")
import ros_gpu_stereo_processor_b200 as m
from ros_gpu_stereo_processor_b200 import _capi as capi
from oracle import synth
import bench, os
NS = int(os.environ.get("SLOTS", "2"))
c = bench.CONFIGS["C4"]; W, H, nd = c["W"], c["H"], c["nd"]
frames, cal = bench.make_frames(c, 2, 4000)
for pft in (1, 0):
    proc = m.GpuStereoProcessor(0)
    info = lambda cc: dict(width=W, height=H, K=cc["K"], D=cc["D"], R=cc["R"], P=cc["P"])
    proc.initStereoModel(info(cal["left"]), info(cal["right"]))
    proc.setParams(numDisparities=nd, blockSize=11, preFilterType=pft, preFilterSize=9, preFilterCap=31, textureThreshold=10, uniquenessRatio=15)
    proc.configureSlots(NS, H, W)
    import torch
    dL = [torch.from_numpy(f[0]).cuda() for f in frames]; dR = [torch.from_numpy(f[1]).cuda() for f in frames]
    io = capi.FrameIO(); io.want = capi.OUT_DISPARITY32F | capi.OUT_POINTCLOUD2; io.rectify = 1; io.inputs_on_device = 1; io.outputs_on_device = 1
    for i in range(4 * NS): proc.processPairAsync(i % NS, dL[i % 2].data_ptr(), dR[i % 2].data_ptr(), io)
    for s_ in range(NS): proc.waitSlot(s_)
    t = time.perf_counter()
    for i in range(200): proc.processPairAsync(i % NS, dL[i % 2].data_ptr(), dR[i % 2].data_ptr(), io)
    for s_ in range(NS): proc.waitSlot(s_)
    print("preFilterType %d: %.1f us/frame" % (pft, (time.perf_counter() - t) / 200 * 1e6))
    proc.close()
