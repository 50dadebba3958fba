"""Generates tests/golden/*.npz.  Run HERE (the build container), where /root/reference exists:

    python tests/golden/make_golden.py

fixtures.npz   the reference's own test data (test/stereobm/test_data/: left/right-0022.png raw pair,
               left/right-0022_rect.png rectification goldens asserted by test/UTest.cpp:247-256,
               left/right.yaml numbers, aloe pair converted to grey by cv2.imread(..., 0), aloe-disp.png)
               re-encoded as arrays, because /root/reference does not exist on the GPU box.
cv2_golden.npz outputs of the real OpenCV (cv2, version recorded inside) on those fixtures for a list of
               parameter sets: cv::StereoBM disparities, speckle-filtered planes, float disparity,
               reprojectImageTo3D points.  The oracle and the CUDA path are both checked against them.
"""
import json
import os
import sys

import cv2
import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle as O, cv2_ref as CV  # noqa: E402

REF = "/root/reference/test/stereobm/test_data/"

PARAM_SETS = {
    # name: (pair, BMParams kwargs)
    "c1_euroc_nd64_b21_xsobel": ("0022", dict(numDisparities=64, blockSize=21)),
    "nd128_b15_uniq_speckle": ("0022", dict(numDisparities=128, blockSize=15, speckleWindowSize=100, speckleRange=4)),
    "norm_ps9": ("0022", dict(preFilterType=0, preFilterSize=9)),
    "ref_cpu_matcher_state": ("0022", dict(preFilterType=0, preFilterSize=5, numDisparities=48, blockSize=19,
                                           textureThreshold=3, uniquenessRatio=0, disp12MaxDiff=0)),
    "neg_minD": ("0022", dict(minDisparity=-16, numDisparities=64, blockSize=11)),
    "nd256_b11": ("0022", dict(numDisparities=256, blockSize=11)),
    "aloe_default": ("aloe", dict()),
    "aloe_nd112_b7_disp12": ("aloe", dict(numDisparities=112, blockSize=7, uniquenessRatio=5, disp12MaxDiff=1, textureThreshold=0)),
    "aloe_b31_cap63": ("aloe", dict(blockSize=31, preFilterCap=63)),
    "aloe_b51_cap63": ("aloe", dict(numDisparities=32, blockSize=51, preFilterCap=63, uniquenessRatio=30)),
}


def load_yaml(f):
    y = yaml.safe_load(open(f))
    return dict(K=y["camera_matrix"]["data"], D=y["distortion_coefficients"]["data"],
                R=y["rectification_matrix"]["data"], P=y["projection_matrix"]["data"],
                W=y["image_width"], H=y["image_height"])


def main():
    fx = {}
    for side in ("left", "right"):
        fx[side + "_raw"] = cv2.imread(REF + side + "-0022.png", 0)
        fx[side + "_rect"] = cv2.imread(REF + side + "-0022_rect.png", 0)
        c = load_yaml(REF + side + ".yaml")
        for k in "KDRP":
            fx[side + "_" + k] = np.array(c[k], np.float64)
    fx["aloe_L"] = cv2.imread(REF + "aloe-L.png", 0)
    fx["aloe_R"] = cv2.imread(REF + "aloe-R.png", 0)
    fx["aloe_L_bgr"] = cv2.imread(REF + "aloe-L.png", 1)[:64, :96].copy()  # small colour crop for the bgr path
    fx["aloe_cuda_disp"] = cv2.imread(REF + "aloe-disp.png", 0)
    np.savez_compressed(os.path.join(HERE, "fixtures.npz"), **fx)

    out = {"cv2_version": np.array(cv2.__version__)}
    pairs = {"0022": (fx["left_rect"], fx["right_rect"]), "aloe": (fx["aloe_L"], fx["aloe_R"])}
    meta = {}
    for name, (pair, kw) in PARAM_SETS.items():
        p = O.BMParams(**kw)
        L, R = pairs[pair]
        d = CV.stereobm_compute(L, R, p)
        out["disp_" + name] = d
        meta[name] = dict(pair=pair, params=p.as_dict())
    # speckle + reprojection vectors on one plane
    d = out["disp_nd128_b15_uniq_speckle"]
    raw = CV.stereobm_compute(*pairs["0022"], O.BMParams(numDisparities=128, blockSize=15))
    out["speckle_in"] = raw
    out["speckle_out_100_4"] = CV.filter_speckles(raw, -16, 100, 4)
    out["speckle_out_800_80"] = CV.filter_speckles(raw, -16, 800, 80)
    Q = O.stereo_Q(fx["left_P"], fx["right_P"])
    df = CV.disparity_to_float(d, fx["left_P"][2] - fx["right_P"][2])
    out["Q"] = Q
    out["df_nd128"] = df
    out["xyz_nd128"] = CV.reproject(df, Q)[::4, ::4].copy()  # subsampled to keep the file small
    out["meta_json"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "cv2_golden.npz"), **out)
    for f in ("fixtures.npz", "cv2_golden.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
