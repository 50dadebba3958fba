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


def test_ros_glue_compiles_against_stub_headers():
    """The ROS 1 glue (ros/src: StereoProcessor, node, nodelet) and the cv::Mat adapters of the C++ facade pass a syntax
    check against minimal stand-ins of the ROS / boost / OpenCV headers (tests/cpp/ros_stubs); ROS itself is absent here."""
    inc = ["-I" + os.path.join(ROOT, "tests", "cpp", "ros_stubs"), "-I" + os.path.join(ROOT, "ros", "include"), "-I" + os.path.join(ROOT, "include")]
    for src in ("ros/src/StereoProcessor.cpp", "ros/src/StereoProcessorNode.cpp", "ros/src/StereoProcessorNodelet.cpp"):
        out = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror"] + inc + [os.path.join(ROOT, src)], capture_output=True, text=True)
        assert out.returncode == 0, src + "\n" + out.stderr
    out = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-DB200S_WITH_OPENCV"] + inc +
                         [os.path.join(ROOT, "tests", "cpp", "opencv_adapters_check.cpp")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    # the topic and parameter surface of the reference (src/StereoProcessor.cpp:29-101) is all there
    txt = open(os.path.join(ROOT, "ros", "src", "StereoProcessor.cpp")).read()
    for name in ("left/image_mono", "right/image_mono", "left/image_color", "right/image_color", "left/rect_mono", "right/rect_mono",
                 "left/rect_color", "right/rect_color", "disparity", "disparity_vis", "pointcloud", "queue_size", "approximate_sync",
                 "camera_info_file_left", "camera_info_file_right", "publisher_queue_size"):
        assert '"%s"' % name in txt, name


def test_gpu_cfg_keeps_the_reference_parameters():
    """cfg/GPU.cfg: every parameter of the reference's cfg/GPU.cfg:12-35 with the same name, type and default."""
    txt = open(os.path.join(ROOT, "cfg", "GPU.cfg")).read()
    ref = {"xsobel": ("bool_t", "False"), "refine_disparity": ("bool_t", "False"), "correlation_window_size": ("int_t", "15"),
           "disparity_min": ("int_t", "0"), "disparity_range": ("int_t", "128"), "bilateral_filter": ("bool_t", "False"),
           "filter_ndisp": ("int_t", "64"), "filter_radius": ("int_t", "3"), "filter_iters": ("int_t", "1"),
           "filter_edge_threshold": ("double_t", "0.1"), "filter_max_disc_threshold": ("double_t", "0.2"), "filter_sigma_range": ("double_t", "10"),
           "texture_threshold": ("double_t", "10"), "max_speckle_size": ("int_t", "800"), "max_speckle_diff": ("double_t", "5")}
    for name, (typ, default) in ref.items():
        m = re.search(r'gen\.add\("%s",\s*(\w+),\s*0,\s*"[^"]*",\s*([^,\)]+)' % name, txt)
        assert m, name
        assert m.group(1) == typ and m.group(2).strip() == default, (name, m.groups())
