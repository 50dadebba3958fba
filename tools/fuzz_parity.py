"""Random-parameter parity sweep of the matcher (all kernels behind computeDisparityBare) against the CPU oracle.
usage: python tools/fuzz_parity.py [n_cases] [seed]      (test infrastructure, like tests/)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ros_gpu_stereo_processor_b200 as m
from oracle import oracle as O, synth

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
proc = m.GpuStereoProcessor(0)
bad = 0
for case in range(n_cases):
    nd = int(rng.choice([16, 32, 48, 64, 80, 96, 112, 128, 160, 192, 256]))
    b = int(rng.choice([5, 7, 9, 11, 13, 15, 17, 19, 21, 23, 31]))
    cap = int(rng.choice([1, 7, 15, 31, 31, 31, 40, 63]))
    kw = dict(numDisparities=nd, blockSize=b, preFilterCap=cap, preFilterType=int(rng.integers(0, 2)),
              preFilterSize=int(rng.choice([5, 9, 21])), textureThreshold=int(rng.choice([0, 3, 10, 50, 400])),
              uniquenessRatio=int(rng.choice([0, 5, 15, 30])), minDisparity=int(rng.choice([0, 0, 0, -16, -5, 8])),
              disp12MaxDiff=int(rng.choice([-1, -1, -1, 0, 1, 2])))
    if rng.random() < 0.3:
        kw.update(speckleWindowSize=int(rng.choice([20, 100, 800])), speckleRange=int(rng.choice([1, 4, 32])))
    W = int(rng.integers(nd + 2 * b + 40, nd + 900))
    H = int(rng.integers(2 * b + 20, 500))
    if rng.random() < 0.15:
        H = int(rng.integers(600, 1300))
    p = O.BMParams(**kw)
    L, R = synth.synth_pair(W, H, max(nd, 16), seed=int(rng.integers(1 << 30)))
    proc.setParams(**p.as_dict())
    try:
        got = proc.computeDisparityBare(L, R)
    except Exception as e:
        print("case %d %dx%d %s: EXC %s" % (case, W, H, kw, e)); bad += 1; continue
    want = O.stereobm_compute(L, R, p)
    lo = max(kw["minDisparity"], 0)      # parity domain for minD > 0
    ok = np.array_equal(got[:, lo:], want[:, lo:])
    if not ok:
        bad += 1
        d = got[:, lo:] != want[:, lo:]
        ys, xs = np.nonzero(d)
        print("case %d %dx%d %s: %d mismatches cols [%d,%d] rows [%d,%d]" % (case, W, H, kw, d.sum(), xs.min() + lo, xs.max() + lo, ys.min(), ys.max()))
print("fuzz: %d cases, %d bad" % (n_cases, bad))
sys.exit(1 if bad else 0)
