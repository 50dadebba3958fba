"""Expected outputs of the whole frame chain, computed on the CPU by the checker.  TEST INFRASTRUCTURE ONLY.

The chain is the reference's CPU path (StereoProcessor::imageCb order, src/StereoProcessor.cpp:157-298):
cv::remap x2 -> cv::StereoBM -> convertTo(CV_32F, 1/16, -(cx-cx')) -> reprojectImageTo3D -> PointCloud2 fill
(src/GpuSenderPc2.cpp:15-72).  `engine="cv2"` runs the real OpenCV (oracle/cv2_ref.py), `engine="oracle"` the C
restatement (oracle/stereo_oracle.c); the CPU tests pin both against each other and against the reference's goldens.
"""
import numpy as np

from oracle import oracle as O, cv2_ref as CV


def expected_chain(L, R, cal, p, rectify, engine="cv2", color=None):
    """Returns dict(rect_left, rect_right, disparity16, disparity32f, points_xyz, pointcloud2) for one raw pair.
    color: optional (H, W, 3) BGR raw left image; when given it is rectified too and colours the cloud
    (StereoProcessor.cpp:201-217 feeds L_RECT_COLOR to enqueueSendPoints)."""
    E = CV if engine == "cv2" else O
    out = {}
    if rectify:
        rl, rr = E.rectify(L, **cal["left"]), E.rectify(R, **cal["right"])
    else:
        rl, rr = np.ascontiguousarray(L), np.ascontiguousarray(R)
    out["rect_left"], out["rect_right"] = rl, rr
    d = E.stereobm_compute(rl, rr, p)
    out["disparity16"] = d
    cxd = cal["left"]["P"][2] - cal["right"]["P"][2]
    df = E.disparity_to_float(d, cxd)
    out["disparity32f"] = df
    xyz = E.reproject(df, O.stereo_Q(cal["left"]["P"], cal["right"]["P"]))
    out["points_xyz"] = xyz
    col = rl
    if color is not None:
        col = E.rectify(color, **cal["left"]) if rectify else np.ascontiguousarray(color)
        out["rect_color_left"] = col
    out["pointcloud2"] = O.pack_pointcloud2(xyz, col)
    return out


def first_mismatch(got, want):
    got, want = np.asarray(got), np.asarray(want)
    if got.shape != want.shape:
        return "shape %s vs %s" % (got.shape, want.shape)
    bad = got != want
    if got.dtype.kind == "f":
        bad &= ~(np.isnan(got) & np.isnan(want))
    n = int(bad.sum())
    if n == 0:
        return ""
    idx = tuple(int(v[0]) for v in np.nonzero(bad))
    return "%d mismatches, first at %s: got %s want %s" % (n, idx, got[idx], want[idx])
