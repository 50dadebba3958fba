"""CPU-side checks of the product library and host logic (no GPU, no compute calls)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    from ros_gpu_stereo_processor_b200 import build
    build.build_library()
    from ros_gpu_stereo_processor_b200 import _capi
    _capi.load()
    return _capi


def test_library_exports_every_declared_symbol(capi):
    hdr = open(os.path.join(ROOT, "include", "b200_stereo.h")).read()
    declared = set(re.findall(r"\b(b200s_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"b200s_error"}
    lib = C.CDLL(capi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)


def test_struct_layouts_match_header(capi):
    assert C.sizeof(capi.Params) == 12 * 4
    assert C.sizeof(capi.CamInfo) == 8 + 9 * 8 + 8 * 8 + 8 + 9 * 8 + 12 * 8
    assert C.sizeof(capi.Pc2Meta) == 40 and C.sizeof(capi.DisparityMeta) == 48
    assert C.sizeof(capi.FrameIO) == 16 + 6 * 8 + 8 + 4 * 4 + 8    # + color_left, color_encoding/rows/cols (+ pad), rect_color_left


def test_default_params_are_cv_stereobm_defaults(capi):
    p = capi.Params()
    assert capi.load().b200s_default_params(C.byref(p)) == 0
    got = {n: getattr(p, n) for n, _ in p._fields_}
    assert got == dict(pre_filter_type=1, pre_filter_size=9, pre_filter_cap=31, block_size=21, min_disparity=0,
                       num_disparities=64, texture_threshold=10, uniqueness_ratio=15, speckle_window_size=0,
                       speckle_range=0, disp12_max_diff=-1, refine_disparity=0)


def test_no_device_fails_loudly(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import ros_gpu_stereo_processor_b200 as m
    with pytest.raises(capi.B200StereoError) as e:
        m.GpuStereoProcessor(0)
    assert e.value.code == capi.ECUDA


def test_gpu_mat_source_ids_match_reference_enum():
    import ros_gpu_stereo_processor_b200 as m
    # include/gpuimageproc/GPUStereoProcessor.h:21-57
    assert (m.GPU_MAT_SIDE_L, m.GPU_MAT_SIDE_R) == (1, 2)
    assert m.GPU_MAT_SRC_L_RAW == 5 and m.GPU_MAT_SRC_R_RAW == 6
    assert m.GPU_MAT_SRC_L_RECT_MONO == (1 << 5) | 1 and m.GPU_MAT_SRC_R_DISPARITY == (1 << 7) | 2
    assert m.GPU_MAT_SRC_L_POINTS2 == (1 << 10) | 1 and m.GPU_MAT_SRC_DISPARITY_IMG == 1 << 9


def test_product_never_touches_the_oracle():
    """The product path must not import, link or execute anything under oracle/ (and has no CPU fallback)."""
    pkg = os.path.join(ROOT, "ros_gpu_stereo_processor_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in os.path.relpath(dirpath, pkg).split(os.sep)[:1] and dirpath != pkg:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no CPU fallback", ""), os.path.join(dirpath, f)
                assert "cv2" not in txt or f.endswith(".cu"), os.path.join(dirpath, f)
    out = subprocess.run(["ldd", os.path.join(pkg, "libb200stereo.so")], capture_output=True, text=True).stdout
    assert "liboracle" not in out and "opencv" not in out
