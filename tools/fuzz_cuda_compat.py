"""Random parity sweep of the cuda-compat matcher (cv::cuda::StereoBM bytes) against the numpy restatement.
usage: python tools/fuzz_cuda_compat.py [n_cases] [seed]      (test infrastructure, like tests/)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ros_gpu_stereo_processor_b200 as m
from oracle import oracle as O, synth

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
proc = m.GpuStereoProcessor(0)
bad = 0
for case in range(n_cases):
    nd = int(rng.choice([16, 32, 48, 64, 96, 128, 256]))
    b = int(rng.choice([5, 7, 9, 15, 19, 21, 33, 51]))
    xs = bool(rng.integers(0, 2))
    cap = int(rng.choice([1, 15, 31, 63]))
    tex = int(rng.choice([0, 3, 10, 60]))
    W = int(rng.integers(nd + 2 * b + 10, nd + 2 * b + 400))
    H = int(rng.integers(2 * b + 5, 2 * b + 200))
    L, R = synth.synth_pair(W, H, max(nd, 16), seed=int(rng.integers(1 << 30)))
    if rng.random() < 0.5:
        L[H // 4:H // 2, W // 3:2 * W // 3] = int(rng.integers(0, 256))      # flat patch for the textureness filter
    proc.setParams(numDisparities=nd, blockSize=b, preFilterType=1 if xs else 0, preFilterCap=cap, textureThreshold=tex)
    got = proc.computeDisparityCudaCompat(L, R)
    want = O.cuda_stereobm(L, R, nd, b, xs, cap, tex)
    if not np.array_equal(got, want):
        bad += 1
        d = got != want
        ys, xs_ = np.nonzero(d)
        print("case %d %dx%d nd%d b%d xsobel%d cap%d tex%d: %d mismatches cols [%d,%d] rows [%d,%d]" % (
            case, W, H, nd, b, xs, cap, tex, d.sum(), xs_.min(), xs_.max(), ys.min(), ys.max()))
print("fuzz_cuda_compat: %d cases, %d bad" % (n_cases, bad))
sys.exit(1 if bad else 0)
