"""The real OpenCV (cv2) run of the reference's CPU path.  TEST INFRASTRUCTURE ONLY.

The reference's own arithmetic for this path lives in OpenCV (cv::StereoBM, cv::remap,
cv::filterSpeckles, cv::reprojectImageTo3D -- call sites src/GPUStereoProcessor.cpp:252-262,
312-321, 332-346, 367-385; launch/ros_cpu_stereo_processing.launch:4-12 = stock stereo_image_proc).
The reference cannot be compiled here (catkin/roscpp/forked OpenCV-CUDA), but the same OpenCV
functions are importable as cv2 in this image (and on the GPU box -- same image), so this module is
  * the ground truth the C restatement (stereo_oracle.c) is pinned against, and
  * the CPU baseline timed by bench.py (`cpu_baseline` leg and `--impl reference`).
"""
import numpy as np
import cv2

from .oracle import BMParams, stereo_Q  # noqa: F401


def make_bm(p):
    bm = cv2.StereoBM_create(numDisparities=p.numDisparities, blockSize=p.blockSize)
    bm.setMinDisparity(p.minDisparity)
    bm.setPreFilterType(p.preFilterType)
    bm.setPreFilterSize(p.preFilterSize)
    bm.setPreFilterCap(p.preFilterCap)
    bm.setTextureThreshold(p.textureThreshold)
    bm.setUniquenessRatio(p.uniquenessRatio)
    bm.setSpeckleWindowSize(p.speckleWindowSize)
    bm.setSpeckleRange(p.speckleRange)
    bm.setDisp12MaxDiff(p.disp12MaxDiff)
    return bm


def stereobm_compute(L, R, p):
    return make_bm(p).compute(np.ascontiguousarray(L), np.ascontiguousarray(R))


def rect_maps(K, D, R, P, W, H):
    K = np.asarray(K, np.float64).reshape(3, 3)
    R = np.asarray(R, np.float64).reshape(3, 3)
    P = np.asarray(P, np.float64).reshape(3, 4)
    D = np.asarray(D, np.float64).ravel()
    return cv2.initUndistortRectifyMap(K, D, R, P, (W, H), cv2.CV_32FC1)


def rectify(src, K, D, R, P):
    H, W = src.shape[:2]
    mx, my = rect_maps(K, D, R, P, W, H)
    return cv2.remap(src, mx, my, cv2.INTER_LINEAR)


def disparity_to_float(d16, cx_minus_cxr):
    # cv::Mat::convertTo(CV_32F, 1/16., -(cx_l - cx_r))  (src/GPUStereoProcessor.cpp:320): computed in double
    return (np.asarray(d16, np.float64) * (1.0 / 16.0) + (-cx_minus_cxr)).astype(np.float32)


def reproject(df, Q):
    return cv2.reprojectImageTo3D(df, np.asarray(Q, np.float64).reshape(4, 4), handleMissingValues=True)


def filter_speckles(img, newVal, maxSize, maxDiff):
    img = img.copy()
    cv2.filterSpeckles(img, newVal, maxSize, maxDiff)
    return img
