"""Seeded synthetic stereo inputs and calibrations.  TEST/BENCH INPUT GENERATION.

Not an oracle: it only manufactures inputs (textured pairs with a known smooth disparity field, scaled
copies of the reference's 752x480 calibration).  Needs cv2 for GaussianBlur/remap/undistortPoints.
"""
import numpy as np
import cv2

# reference calibration fixtures, test/stereobm/test_data/left.yaml:1-20 / right.yaml:1-20 (numbers only)
CALIB_752x480 = dict(
    W=752, H=480,
    left=dict(
        K=[463.241134, 0.0, 385.760092, 0.0, 463.072782, 221.039882, 0.0, 0.0, 1.0],
        D=[-0.374450, 0.109252, 0.000282, -0.001548, 0.000000],
        R=[0.997310, 0.005745, 0.073072, -0.006100, 0.999971, 0.004637, -0.073043, -0.005070, 0.997316],
        P=[441.238411, 0.0, 322.811100, 0.0, 0.0, 441.238411, 230.623768, 0.0, 0.0, 0.0, 1.0, 0.0]),
    right=dict(
        K=[462.751523, 0.0, 347.278408, 0.0, 462.070418, 236.614124, 0.0, 0.0, 1.0],
        D=[-0.379317, 0.114524, -0.000156, -0.001440, 0.000000],
        R=[0.998747, 0.009973, 0.049045, -0.009735, 0.999940, -0.005097, -0.049093, 0.004613, 0.998784],
        P=[441.238411, 0.0, 322.811100, -44.133133, 0.0, 441.238411, 230.623768, 0.0, 0.0, 0.0, 1.0, 0.0]))


def scaled_calibration(W, H):
    """left.yaml/right.yaml scaled by S = diag(W/752, H/480, 1) applied to K and P (D, R unchanged)."""
    sx, sy = W / 752.0, H / 480.0
    out = dict(W=W, H=H)
    for side in ("left", "right"):
        c = CALIB_752x480[side]
        K = np.array(c["K"], np.float64).reshape(3, 3).copy()
        P = np.array(c["P"], np.float64).reshape(3, 4).copy()
        K[0] *= sx; K[1] *= sy
        P[0] *= sx; P[1] *= sy
        out[side] = dict(K=K.ravel().tolist(), D=list(c["D"]), R=list(c["R"]), P=P.ravel().tolist())
    return out


def synth_pair(W, H, nd, seed):
    """Ideal (already rectified) textured pair: R(x) = L(x + d(x, y)), d smooth in [0.15 nd, 0.70 nd]."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (H, W + nd), dtype=np.uint8)
    base = cv2.GaussianBlur(base, (0, 0), 1.2)
    xs, ys = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32))
    d = (0.15 * nd + 0.55 * nd * (0.5 + 0.5 * np.sin(3 * np.pi * xs / W) * np.cos(2 * np.pi * ys / H))).astype(np.float32)
    L = np.ascontiguousarray(base[:, nd:nd + W])
    R = cv2.remap(base, xs + nd + d, ys, cv2.INTER_LINEAR)
    return L, np.ascontiguousarray(R)


_UNRECT_CACHE = {}


def unrectify(ideal, cal_side, W, H):
    """Push an ideal rectified image through the inverse rectification so that rectifying it again
    gives (approximately) the ideal image back."""
    key = (W, H, tuple(cal_side["K"]), tuple(cal_side["D"]), tuple(cal_side["R"]), tuple(cal_side["P"]))
    rp = _UNRECT_CACHE.get(key)
    if rp is None:
        K = np.array(cal_side["K"], np.float64).reshape(3, 3)
        D = np.array(cal_side["D"], np.float64)
        R = np.array(cal_side["R"], np.float64).reshape(3, 3)
        P = np.array(cal_side["P"], np.float64).reshape(3, 4)
        xs, ys = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
        pts = np.stack([xs.ravel(), ys.ravel()], axis=1).reshape(-1, 1, 2)
        rp = cv2.undistortPoints(pts, K, D, R=R, P=P[:, :3]).reshape(H, W, 2).astype(np.float32)
        if len(_UNRECT_CACHE) > 4:
            _UNRECT_CACHE.clear()
        _UNRECT_CACHE[key] = rp
    return cv2.remap(ideal, rp[..., 0], rp[..., 1], cv2.INTER_LINEAR)


def synth_raw_pair(W, H, nd, seed):
    """Raw (unrectified) pair + calibration for the configs that exercise rectification."""
    cal = scaled_calibration(W, H)
    L, R = synth_pair(W, H, nd, seed)
    return unrectify(L, cal["left"], W, H), unrectify(R, cal["right"], W, H), cal
