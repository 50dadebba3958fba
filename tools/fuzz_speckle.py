"""Random parity sweep of the speckle filter (connected-component kernels) against the CPU oracle.
usage: python tools/fuzz_speckle.py [n_cases] [seed]      (test infrastructure, like tests/)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ros_gpu_stereo_processor_b200 as m
from oracle import oracle as O

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
proc = m.GpuStereoProcessor(0)
bad = 0
for case in range(n_cases):
    W, H = int(rng.integers(5, 700)), int(rng.integers(5, 400))
    if rng.random() < 0.1:
        W, H = int(rng.integers(900, 2000)), int(rng.integers(500, 1100))
    newVal = int(rng.choice([-16, 0, -32]))
    kind = int(rng.integers(0, 6))
    yy, xx = np.mgrid[0:H, 0:W]
    if kind == 0:      # noise with a small value range: many tiny and some huge components
        img = rng.integers(0, int(rng.choice([3, 8, 40])), (H, W)) * 16
    elif kind == 1:    # smooth ramp with random holes
        img = (xx * 3 + yy * 2) // int(rng.choice([1, 2, 5]))
        img = np.where(rng.random((H, W)) < rng.choice([0.05, 0.3, 0.6]), newVal, img)
    elif kind == 2:    # one-pixel stripes and checkerboards (worst case for run merging)
        img = np.where((xx + yy * int(rng.integers(0, 2))) % int(rng.choice([2, 3])) == 0, 100, newVal)
        img = img + (rng.random((H, W)) < 0.02) * 50
    elif kind == 3:    # random rectangles of constant value
        img = np.full((H, W), newVal)
        for _ in range(int(rng.integers(5, 200))):
            x0, y0 = int(rng.integers(0, W)), int(rng.integers(0, H))
            w, h = int(rng.integers(1, 90)), int(rng.integers(1, 40))
            img[y0:y0 + h, x0:x0 + w] = int(rng.integers(0, 2000))
    elif kind == 4:    # long snakes crossing many tiles: every other row connected at alternating ends
        img = np.full((H, W), newVal)
        img[::2] = 500
        img[1::4, -1] = 500
        img[3::4, 0] = 500
        img = np.where(rng.random((H, W)) < 0.01, newVal, img)
    else:              # spiral-ish bands with a value gradient close to maxDiff
        img = ((np.hypot(xx - W / 2, yy - H / 2)).astype(np.int64) % 7 < 4) * (xx + yy) + newVal * ((np.hypot(xx - W / 2, yy - H / 2)).astype(np.int64) % 7 >= 4)
    img = np.ascontiguousarray(img, np.int16)
    maxSize = int(rng.choice([0, 1, 5, 50, 200, 1000, 100000]))
    maxDiff = int(rng.choice([0, 1, 2, 16, 100]))
    want = O.filter_speckles(img.copy(), newVal, maxSize, maxDiff)
    proc.setParams(minDisparity=newVal // 16 + 1, speckleWindowSize=maxSize, speckleRange=maxDiff)
    got = proc.filterSpeckles(img.copy())
    if not np.array_equal(got, want):
        bad += 1
        d = got != want
        ys, xs = np.nonzero(d)
        print("case %d kind %d %dx%d newVal %d maxSize %d maxDiff %d: %d mismatches cols [%d,%d] rows [%d,%d]" % (
            case, kind, W, H, newVal, maxSize, maxDiff, d.sum(), xs.min(), xs.max(), ys.min(), ys.max()))
print("fuzz_speckle: %d cases, %d bad" % (n_cases, bad))
sys.exit(1 if bad else 0)
