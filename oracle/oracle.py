"""ctypes front-end of the CPU oracle (oracle/stereo_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package never does (tests/test_no_oracle_in_product.py checks).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")


class BMParams(C.Structure):
    """Mirror of orc_bm_params; field names are cv::StereoBM's."""
    _fields_ = [(n, C.c_int) for n in (
        "minDisparity", "numDisparities", "blockSize", "preFilterType", "preFilterSize", "preFilterCap",
        "textureThreshold", "uniquenessRatio", "speckleWindowSize", "speckleRange", "disp12MaxDiff")]

    def __init__(self, minDisparity=0, numDisparities=64, blockSize=21, preFilterType=1, preFilterSize=9,
                 preFilterCap=31, textureThreshold=10, uniquenessRatio=15, speckleWindowSize=0, speckleRange=0,
                 disp12MaxDiff=-1):
        super().__init__(minDisparity, numDisparities, blockSize, preFilterType, preFilterSize, preFilterCap,
                         textureThreshold, uniquenessRatio, speckleWindowSize, speckleRange, disp12MaxDiff)

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def build(force=False):
    src = os.path.join(_HERE, "stereo_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_num_threads.restype = C.c_int
        _lib.orc_stereobm_compute.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


def set_num_threads(n):
    lib().orc_set_num_threads(C.c_int(int(n)))


def num_threads():
    return lib().orc_num_threads()


def _d8(D):
    D = np.asarray(D, dtype=np.float64).ravel()
    out = np.zeros(8, dtype=np.float64)
    out[:min(8, D.size)] = D[:8]
    return out


def build_rect_map(K, D, R, P, W, H):
    K = np.ascontiguousarray(K, dtype=np.float64).reshape(9)
    R = np.ascontiguousarray(R, dtype=np.float64).reshape(9)
    P = np.ascontiguousarray(P, dtype=np.float64).reshape(12)
    D = _d8(D)
    mx = np.empty((H, W), np.float32)
    my = np.empty((H, W), np.float32)
    lib().orc_build_rect_map(_p(K, C.c_double), _p(D, C.c_double), _p(R, C.c_double), _p(P, C.c_double),
                             C.c_int(W), C.c_int(H), _p(mx, C.c_float), _p(my, C.c_float))
    return mx, my


def remap_linear(src, mapx, mapy):
    src = _u8(src)
    ch = 1 if src.ndim == 2 else src.shape[2]
    sH, sW = src.shape[:2]
    H, W = mapx.shape
    mapx = np.ascontiguousarray(mapx, np.float32)
    mapy = np.ascontiguousarray(mapy, np.float32)
    dst = np.empty((H, W) if ch == 1 else (H, W, ch), np.uint8)
    lib().orc_remap_linear(_p(src, C.c_uint8), C.c_int(sW), C.c_int(sH), C.c_int(ch), _p(mapx, C.c_float),
                           _p(mapy, C.c_float), C.c_int(W), C.c_int(H), _p(dst, C.c_uint8))
    return dst


def rectify(src, K, D, R, P):
    H, W = src.shape[:2]
    mx, my = build_rect_map(K, D, R, P, W, H)
    return remap_linear(src, mx, my)


def prefilter_xsobel(src, cap):
    src = _u8(src)
    H, W = src.shape
    dst = np.empty_like(src)
    lib().orc_prefilter_xsobel(_p(src, C.c_uint8), _p(dst, C.c_uint8), C.c_int(W), C.c_int(H), C.c_int(cap))
    return dst


def prefilter_norm(src, ps, cap, fast=True):
    src = _u8(src)
    H, W = src.shape
    dst = np.empty_like(src)
    fn = lib().orc_prefilter_norm_fast if fast else lib().orc_prefilter_norm
    fn(_p(src, C.c_uint8), _p(dst, C.c_uint8), C.c_int(W), C.c_int(H), C.c_int(ps), C.c_int(cap))
    return dst


def prefilter(src, p):
    if p.preFilterType == 1:
        return prefilter_xsobel(src, p.preFilterCap)
    return prefilter_norm(src, p.preFilterSize, p.preFilterCap)


def bm_core(Lp, Rp, p):
    """Matcher on prefiltered planes: returns (disp s16, cost s16) before validate/ROI mask/speckle."""
    Lp, Rp = _u8(Lp), _u8(Rp)
    H, W = Lp.shape
    disp = np.empty((H, W), np.int16)
    cost = np.empty((H, W), np.int16)
    lib().orc_bm_core(_p(Lp, C.c_uint8), _p(Rp, C.c_uint8), C.c_int(W), C.c_int(H), C.byref(p),
                      _p(disp, C.c_int16), _p(cost, C.c_int16))
    return disp, cost


def stereobm_post(disp, cost, p):
    disp = np.ascontiguousarray(disp, np.int16).copy()
    cost = np.ascontiguousarray(cost, np.int16)
    H, W = disp.shape
    lib().orc_stereobm_post(_p(disp, C.c_int16), _p(cost, C.c_int16), C.c_int(W), C.c_int(H), C.byref(p))
    return disp


def stereobm_compute(L, R, p, return_prefiltered=False):
    """cv::StereoBM::compute equivalent (prefilter + match + validate + ROI mask + speckle) -> s16 x16."""
    L, R = _u8(L), _u8(R)
    H, W = L.shape
    disp = np.empty((H, W), np.int16)
    Lp = np.empty((H, W), np.uint8)
    Rp = np.empty((H, W), np.uint8)
    rc = lib().orc_stereobm_compute(_p(L, C.c_uint8), _p(R, C.c_uint8), C.c_int(W), C.c_int(H), C.byref(p),
                                    _p(disp, C.c_int16), _p(Lp, C.c_uint8), _p(Rp, C.c_uint8))
    if rc != 0:
        raise ValueError("orc_stereobm_compute: invalid parameter (code %d)" % rc)
    return (disp, Lp, Rp) if return_prefiltered else disp


def validate_disp12(disp, cost, minD, nd, disp12MaxDiff, ya=0, yb=None):
    disp = np.ascontiguousarray(disp, np.int16).copy()
    cost = np.ascontiguousarray(cost, np.int16)
    H, W = disp.shape
    yb = H if yb is None else yb
    lib().orc_validate_disp12(_p(disp, C.c_int16), _p(cost, C.c_int16), C.c_int(W), C.c_int(H), C.c_int(minD),
                              C.c_int(nd), C.c_int(disp12MaxDiff), C.c_int(ya), C.c_int(yb))
    return disp


def filter_speckles(img, newVal, maxSize, maxDiff):
    img = np.ascontiguousarray(img, np.int16).copy()
    H, W = img.shape
    lib().orc_filter_speckles(_p(img, C.c_int16), C.c_int(W), C.c_int(H), C.c_int(int(newVal)), C.c_int(int(maxSize)),
                              C.c_int(int(maxDiff)))
    return img


def disparity_to_float(d16, cx_minus_cxr):
    d16 = np.ascontiguousarray(d16, np.int16)
    out = np.empty(d16.shape, np.float32)
    lib().orc_disparity_to_float(_p(d16, C.c_int16), C.c_size_t(d16.size), C.c_double(cx_minus_cxr), _p(out, C.c_float))
    return out


def reproject(df, Q, handle_missing=True):
    df = np.ascontiguousarray(df, np.float32)
    H, W = df.shape
    Q = np.ascontiguousarray(Q, np.float64).reshape(16)
    xyz = np.empty((H, W, 3), np.float32)
    lib().orc_reproject(_p(df, C.c_float), C.c_int(W), C.c_int(H), _p(Q, C.c_double), C.c_int(int(handle_missing)),
                        _p(xyz, C.c_float))
    return xyz


def pack_pointcloud2(xyz, color):
    xyz = np.ascontiguousarray(xyz, np.float32)
    color = _u8(color)
    H, W = xyz.shape[:2]
    ch = 1 if color.ndim == 2 else color.shape[2]
    out = np.empty((H, W, 32), np.uint8)
    lib().orc_pack_pointcloud2(_p(xyz, C.c_float), _p(color, C.c_uint8), C.c_int(ch), C.c_int(W), C.c_int(H),
                               _p(out, C.c_uint8))
    return out


# ---- stereo model helpers (image_geometry::StereoCameraModel::updateQ) -----------------

def stereo_Q(Pl, Pr):
    """Q of image_geometry's StereoCameraModel (newer closed form; both published forms agree to 1.8e-7)."""
    Pl = np.asarray(Pl, np.float64).reshape(3, 4)
    Pr = np.asarray(Pr, np.float64).reshape(3, 4)
    fx, fy, cx, cy = Pl[0, 0], Pl[1, 1], Pl[0, 2], Pl[1, 2]
    cxr = Pr[0, 2]
    Tx = Pr[0, 3] / Pr[0, 0]  # = -baseline
    Q = np.zeros((4, 4), np.float64)
    Q[0, 0] = fy * Tx
    Q[0, 3] = -fy * cx * Tx
    Q[1, 1] = fx * Tx
    Q[1, 3] = -fx * cy * Tx
    Q[2, 3] = fx * fy * Tx
    Q[3, 2] = -fy
    Q[3, 3] = fy * (cx - cxr)
    return Q


def valid_window(W, H, minD, nd, wsz):
    """DisparityImage.valid_window as GpuSenderDisparity.cpp:30-39 / stereo_image_proc compute it."""
    border = wsz // 2
    left = nd + minD + border - 1
    wtf = border + minD if minD >= 0 else max(border, -minD)
    right = W - 1 - wtf
    top = border
    bottom = H - 1 - border
    return dict(x_offset=left, y_offset=top, width=right - left, height=bottom - top)


def draw_color_disp(d16, nd):
    """cv::cuda::drawColorDisp restated (opencv_contrib cudastereo util.cu, cvtPixel) on d = clamp(d16 >> 4, 0, 255).
    PARITY UNPINNED: the upstream kernel is not in /root/reference and cv2 has no CUDA modules; no golden exists."""
    d = np.clip(np.asarray(d16, np.int32) >> 4, 0, 255)
    a = (nd - d).astype(np.int64) * 240
    H = (np.sign(a) * (np.abs(a) // nd)) & 0xFFFFFFFF   # C int division (truncating), then conversion to unsigned
    H = H.astype(np.uint32)
    hi = (H // 60) % 6
    f = (H.astype(np.float32) / np.float32(60.0)) - (H // 60).astype(np.float32)
    one, zero = np.ones_like(f), np.zeros_like(f)
    q, t = one - f, f
    x = np.select([hi == 0, hi == 1, hi == 2, hi == 3, hi == 4], [zero, zero, t, one, one], q)
    y = np.select([hi == 0, hi == 1, hi == 2, hi == 3, hi == 4], [t, one, one, q, zero], zero)
    z = np.select([hi == 0, hi == 1, hi == 2, hi == 3, hi == 4], [one, q, zero, zero, t], one)
    out = np.empty(d.shape + (4,), np.uint8)
    for c, v in enumerate((x, y, z)):
        out[..., c] = (np.clip(v, 0, 1).astype(np.float32) * np.float32(255.0)).astype(np.uint32).astype(np.uint8)
    out[..., 3] = 255
    return out


# ---- cv::cuda::StereoBM restated: what the reference's GPU path computes at
# GPUStereoProcessor.cpp:283 (block_matcher_gpu_->compute).  Upstream: opencv_contrib cudastereo stereobm.cu. -----------
def cuda_prefilter_xsobel(img, cap):
    """prefilter_kernel_xsobel: full 3x3 Sobel-x on clamp-addressed pixels, min(clip(v, -cap, cap) + cap, 255)."""
    p = np.pad(np.asarray(img, np.int32), 1, mode="edge")
    conv = (-p[:-2, :-2] + p[:-2, 2:] - 2 * p[1:-1, :-2] + 2 * p[1:-1, 2:] - p[2:, :-2] + p[2:, 2:])
    return np.minimum(np.clip(conv, -cap, cap) + cap, 255).astype(np.uint8)


def _box_sum(a, r):
    """exact integer box sum over (2r+1)^2, 'valid' region only: out[y, x] = sum a[y:y+2r+1, x:x+2r+1]"""
    c = np.cumsum(np.cumsum(np.pad(a.astype(np.int64), ((1, 0), (1, 0))), axis=0), axis=1)
    b = 2 * r + 1
    return c[b:, b:] - c[:-b, b:] - c[b:, :-b] + c[:-b, :-b]


def cuda_textureness_mask(img, wsz, avg_tex_threshold):
    """postfilter_textureness: pixels whose window sum of |Sobel-x| (clamp-addressed, 3x3) is below
    avergeTexThreshold * wsz * wsz.  Upstream sums normalised floats (value / 255) and scales by 255; this is the same
    quantity in exact integer arithmetic (PARITY UNPINNED at the rounding boundary: no golden removes a pixel)."""
    r = wsz // 2
    H, W = img.shape
    # the Sobel itself is evaluated at clamp-addressed coordinates: sobel(x, y) for x, y outside the image
    # uses tex2D clamping of every tap, which equals evaluating on the edge-padded image
    pp = np.pad(np.asarray(img, np.int32), r + 1, mode="edge")
    s = np.abs(-pp[:-2, :-2] + pp[:-2, 2:] - 2 * pp[1:-1, :-2] + 2 * pp[1:-1, 2:] - pp[2:, :-2] + pp[2:, 2:])
    win = _box_sum(s, r)                      # H x W
    return win < int(avg_tex_threshold) * wsz * wsz


def cuda_stereobm(L, R, nd, wsz, xsobel=False, cap=31, tex_threshold=3):
    """u8 integer disparity, 0 = not computed / filtered.  SSD over wsz^2; candidates d = 0..nd-1 scanned ascending in
    groups of 8: inside a group the HIGHEST index with the minimum wins, across groups strict '<' (earlier group wins).
    Computed for x in [nd + r, W - r), y in [r, H - r)."""
    L = np.ascontiguousarray(L, np.uint8)
    R = np.ascontiguousarray(R, np.uint8)
    if xsobel:
        L, R = cuda_prefilter_xsobel(L, cap), cuda_prefilter_xsobel(R, cap)
    H, W = L.shape
    r = wsz // 2
    out = np.zeros((H, W), np.uint8)
    x0, x1, y0, y1 = nd + r, W - r, r, H - r
    if x1 > x0 and y1 > y0:
        Li, Ri = L.astype(np.int64), R.astype(np.int64)
        best = np.full((y1 - y0, x1 - x0), np.iinfo(np.int64).max)
        bd = np.zeros((y1 - y0, x1 - x0), np.int64)
        for g in range(0, nd, 8):
            ssd = []
            for j in range(8):
                d = g + j
                sq = np.zeros((H, W), np.int64)
                sq[:, d:] = (Li[:, d:] - Ri[:, :W - d]) ** 2
                ssd.append(_box_sum(sq, r)[:, x0 - r:x1 - r])      # rows y0..y1, columns x0..x1
            ssd = np.stack(ssd)
            m = ssd.min(axis=0)
            idx = 7 - np.argmin(ssd[::-1], axis=0)                 # last index that holds the minimum
            upd = m < best
            best = np.where(upd, m, best)
            bd = np.where(upd, g + idx, bd)
        out[y0:y1, x0:x1] = bd.astype(np.uint8)
    if tex_threshold > 0:
        out[cuda_textureness_mask(L, wsz, tex_threshold)] = 0
    return out
