"""Per-frame time of the reference-shaped call sequence (uploadMat / rectifyImage / computeDisparity / filterSpeckles /
projectDisparityTo3DPoints / pack) through the named-buffer API, C4 size, host clock."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ros_gpu_stereo_processor_b200 as m
import bench
c = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C4"]
W, H, nd = c["W"], c["H"], c["nd"]
frames, cal = bench.make_frames(c, 2, 4000)
proc = m.GpuStereoProcessor(0)
info = lambda cc: dict(width=W, height=H, K=cc["K"], D=cc["D"], R=cc["R"], P=cc["P"])
proc.initStereoModel(info(cal["left"]), info(cal["right"]))
proc.setParams(numDisparities=nd, blockSize=c["block"], preFilterType=1, preFilterCap=31, textureThreshold=10, uniquenessRatio=15, disp12MaxDiff=-1)
L, R = m.SIDE_L, m.SIDE_R
def frame(i, pack):
    l, r = frames[i % 2]
    proc.uploadMat(m.SRC_RAW | L, l, "mono8"); proc.uploadMat(m.SRC_RAW | R, r, "mono8")
    proc.convertRawToMono(L); proc.convertRawToMono(R)
    proc.rectifyImage(m.SRC_MONO | L, m.SRC_RECT_MONO | L); proc.rectifyImage(m.SRC_MONO | R, m.SRC_RECT_MONO | R)
    proc.computeDisparity(m.SRC_RECT_MONO | L, m.SRC_RECT_MONO | R, m.SRC_DISPARITY | L)
    proc.filterSpeckles(m.SRC_DISPARITY | L)
    proc.projectDisparityTo3DPoints(m.SRC_DISPARITY | L, m.SRC_POINTS2 | L)
    if pack:
        return proc.enqueueSendPoints(m.SRC_POINTS2 | L, m.SRC_RECT_MONO | L)
    proc.waitForAllStreams()
for pack in (False, True):
    for i in range(10): frame(i, pack)
    t = time.perf_counter()
    n = 40
    for i in range(n): frame(i, pack)
    print("%s named-buffer API, %s: %.2f ms/frame" % (sys.argv[1] if len(sys.argv) > 1 else "C4", "with PointCloud2 payload to host" if pack else "device only", (time.perf_counter() - t) / n * 1e3))
