import time, numpy as np, sys, os
sys.path.insert(0, ".")
import ros_gpu_stereo_processor_b200 as m
from oracle import synth
p = m.GpuStereoProcessor(0)
for (W,H,nd,b) in ((640,553,128,19),(1920,1080,256,11)):
    L,R = synth.synth_pair(W,H,nd,1)
    p.setParams(numDisparities=nd, blockSize=b, preFilterType=1, textureThreshold=3)
    p.uploadMat(m.SRC_RECT_MONO|m.SIDE_L, L); p.uploadMat(m.SRC_RECT_MONO|m.SIDE_R, R)
    for i in range(5): p.computeDisparityCudaCompat(m.SRC_RECT_MONO|m.SIDE_L, m.SRC_RECT_MONO|m.SIDE_R, m.SRC_DISPARITY|m.SIDE_L)
    p.waitForAllStreams()
    t=time.perf_counter()
    for i in range(20): p.computeDisparityCudaCompat(m.SRC_RECT_MONO|m.SIDE_L, m.SRC_RECT_MONO|m.SIDE_R, m.SRC_DISPARITY|m.SIDE_L)
    p.waitForAllStreams()
    print("rpt=%s cuda-compat %dx%d nd%d b%d: %.1f us/frame" % (os.environ.get("B200S_CC_RPT"),W,H,nd,b,(time.perf_counter()-t)/20*1e6))
