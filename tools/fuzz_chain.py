"""Random-calibration parity sweep of rectify -> disparity -> float / reproject / PointCloud2 pack against the CPU oracle.
usage: python tools/fuzz_chain.py [n_cases] [seed]      (test infrastructure, like tests/)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ros_gpu_stereo_processor_b200 as m
from oracle import oracle as O, synth

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0


def rot(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


for case in range(n_cases):
    W, H = int(rng.integers(160, 900)), int(rng.integers(120, 600))
    nd = int(rng.choice([16, 32, 64, 128]))
    if W < nd + 80:
        W = nd + 80 + int(rng.integers(0, 100))
    b = int(rng.choice([5, 9, 11, 15, 21]))
    f = float(rng.uniform(0.6, 1.4) * W)
    cal = {}
    T = float(rng.uniform(0.05, 0.3))
    fp = f * float(rng.uniform(0.9, 1.1))
    cxp, cyp = W / 2 + float(rng.uniform(-20, 20)), H / 2 + float(rng.uniform(-20, 20))
    for side in ("left", "right"):
        K = [f * float(rng.uniform(0.97, 1.03)), 0, W / 2 + float(rng.uniform(-15, 15)), 0, f * float(rng.uniform(0.97, 1.03)),
             H / 2 + float(rng.uniform(-15, 15)), 0, 0, 1]
        nD = int(rng.choice([5, 5, 8]))
        D = [float(rng.uniform(-0.3, 0.2)), float(rng.uniform(-0.1, 0.15)), float(rng.uniform(-2e-3, 2e-3)), float(rng.uniform(-2e-3, 2e-3)),
             float(rng.uniform(-0.05, 0.05))]
        if nD == 8:
            D += [float(rng.uniform(-0.05, 0.05)), float(rng.uniform(-0.02, 0.02)), float(rng.uniform(-0.01, 0.01))]
        Rm = rot(*rng.uniform(-0.02, 0.02, 3))
        P = [fp, 0, cxp, (-fp * T if side == "right" else 0.0), 0, fp, cyp, 0, 0, 0, 1, 0]
        cal[side] = dict(K=K, D=D, R=Rm.ravel().tolist(), P=P)
    Lraw = rng.integers(0, 256, (H, W), dtype=np.uint8)
    Lraw = np.ascontiguousarray(np.clip(Lraw.astype(np.int32) // 2 + np.roll(Lraw, 3, axis=1) // 2, 0, 255).astype(np.uint8))
    Rraw = np.ascontiguousarray(np.roll(Lraw, -int(rng.integers(2, max(3, nd // 2))), axis=1))
    p = O.BMParams(numDisparities=nd, blockSize=b, textureThreshold=0, uniquenessRatio=int(rng.choice([0, 10])),
                   preFilterType=int(rng.integers(0, 2)), preFilterSize=int(rng.choice([5, 9, 21, 31])))
    proc = m.GpuStereoProcessor(0)
    info = lambda c: dict(width=W, height=H, K=c["K"], D=c["D"], R=c["R"], P=c["P"])
    proc.initStereoModel(info(cal["left"]), info(cal["right"]))
    proc.setParams(**p.as_dict())
    if rng.random() < 0.5:
        proc._ck(proc._lib.b200s_set_rectify_mode(proc._h, 1))        # map evaluated on the fly
    out = proc.processPair(Lraw, Rraw, rectify=True, want=("rect_left", "rect_right", "disparity16", "disparity32f", "pointcloud2", "points_xyz"))
    proc.close()
    rl, rr = O.rectify(Lraw, **cal["left"]), O.rectify(Rraw, **cal["right"])
    d = O.stereobm_compute(rl, rr, p)
    cxd = cal["left"]["P"][2] - cal["right"]["P"][2]
    df = O.disparity_to_float(d, cxd)
    xyz = O.reproject(df, O.stereo_Q(cal["left"]["P"], cal["right"]["P"]))
    pc = O.pack_pointcloud2(xyz, rl)
    checks = dict(rect_left=np.array_equal(out["rect_left"], rl), rect_right=np.array_equal(out["rect_right"], rr),
                  disparity16=np.array_equal(out["disparity16"], d), disparity32f=np.array_equal(out["disparity32f"], df),
                  points_xyz=np.array_equal(out["points_xyz"].view(np.uint32), xyz.view(np.uint32)),
                  pointcloud2=np.array_equal(out["pointcloud2"], pc))
    if not all(checks.values()):
        bad += 1
        print("case %d %dx%d nd%d b%d nD%d: %s" % (case, W, H, nd, b, len(cal["left"]["D"]), {k: v for k, v in checks.items() if not v}))
print("fuzz_chain: %d cases, %d bad" % (n_cases, bad))
sys.exit(1 if bad else 0)
