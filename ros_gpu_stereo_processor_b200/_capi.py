"""ctypes binding of libb200stereo.so (include/b200_stereo.h).  No fallback: a missing library raises."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# B200S_LIB_SUFFIX selects an experimental build variant (ros_gpu_stereo_processor_b200/build.py); the product is the plain name
LIB_PATH = os.path.join(HERE, "libb200stereo%s.so" % os.environ.get("B200S_LIB_SUFFIX", ""))

# error codes (b200s_error)
OK, EINVAL, ENOTINIT, ECUDA, ENOMEM, EUNSUPPORTED, EIO, ENOBUF = 0, -1, -2, -3, -4, -5, -6, -7
ERROR_NAMES = {EINVAL: "B200S_EINVAL", ENOTINIT: "B200S_ENOTINIT", ECUDA: "B200S_ECUDA", ENOMEM: "B200S_ENOMEM",
               EUNSUPPORTED: "B200S_EUNSUPPORTED", EIO: "B200S_EIO", ENOBUF: "B200S_ENOBUF"}

# GpuMatSource (include/gpuimageproc/GPUStereoProcessor.h:21-57)
SIDE_L, SIDE_R = 1, 2
SRC_RAW, SRC_MONO, SRC_COLOR, SRC_RECT_MONO, SRC_RECT_COLOR = 4, 8, 16, 32, 64
SRC_DISPARITY, SRC_DISPARITY_32F, SRC_DISPARITY_IMG, SRC_POINTS2 = 128, 256, 512, 1024

T_8UC1, T_16SC1, T_32FC1, T_8UC3, T_32FC3, T_8UC4 = 0, 3, 5, 16, 21, 24
INTER_NEAREST, INTER_LINEAR = 0, 1

OUT_RECT_L, OUT_RECT_R, OUT_DISPARITY16, OUT_DISPARITY32F, OUT_POINTCLOUD2, OUT_POINTS_XYZ, OUT_RECT_COLOR_L = 1, 2, 4, 8, 16, 32, 64
COLOR_NONE, COLOR_BGR8, COLOR_RGB8 = 0, 1, 2
STAGE_NAMES = ("rectify_prefilter", "match", "post", "to_float", "reproject_pack")
MAX_BATCH = 32


class CamInfo(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("K", C.c_double * 9), ("D", C.c_double * 8),
                ("n_D", C.c_int), ("R", C.c_double * 9), ("P", C.c_double * 12)]


class Params(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "pre_filter_type", "pre_filter_size", "pre_filter_cap", "block_size", "min_disparity", "num_disparities",
        "texture_threshold", "uniqueness_ratio", "speckle_window_size", "speckle_range", "disp12_max_diff",
        "refine_disparity")]


class DisparityMeta(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("step", C.c_int), ("f", C.c_float), ("T", C.c_float),
                ("min_disparity", C.c_float), ("max_disparity", C.c_float), ("delta_d", C.c_float),
                ("valid_x_offset", C.c_int), ("valid_y_offset", C.c_int), ("valid_width", C.c_int),
                ("valid_height", C.c_int)]


class Pc2Meta(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("width", "height", "point_step", "row_step", "is_bigendian", "is_dense",
                                       "off_x", "off_y", "off_z", "off_rgb")]


class FrameIO(C.Structure):
    _fields_ = [("want", C.c_uint32), ("rectify", C.c_int), ("inputs_on_device", C.c_int),
                ("outputs_on_device", C.c_int), ("rect_left", C.c_void_p), ("rect_right", C.c_void_p),
                ("disparity16", C.c_void_p), ("disparity32f", C.c_void_p), ("pointcloud2", C.c_void_p),
                ("points_xyz", C.c_void_p), ("color_left", C.c_void_p), ("color_encoding", C.c_int),
                ("rows", C.c_int), ("cols", C.c_int), ("rect_color_left", C.c_void_p)]


DONE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int)


# every symbol include/b200_stereo.h declares: name -> (restype, argtypes)
H = C.c_void_p
SYMBOLS = {
    "b200s_create": (C.c_int, [C.c_int, C.POINTER(H)]),
    "b200s_destroy": (C.c_int, [H]),
    "b200s_last_error_string": (C.c_char_p, [H]),
    "b200s_version": (C.c_char_p, []),
    "b200s_set_calibration": (C.c_int, [H, C.POINTER(CamInfo), C.POINTER(CamInfo)]),
    "b200s_load_calibration_files": (C.c_int, [H, C.c_char_p, C.c_char_p]),
    "b200s_is_model_initialised": (C.c_int, [H]),
    "b200s_get_model": (C.c_int, [H, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "b200s_default_params": (C.c_int, [C.POINTER(Params)]),
    "b200s_set_params": (C.c_int, [H, C.POINTER(Params)]),
    "b200s_get_params": (C.c_int, [H, C.POINTER(Params)]),
    "b200s_set_rectify_mode": (C.c_int, [H, C.c_int]),
    "b200s_upload": (C.c_int, [H, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_char_p]),
    "b200s_download": (C.c_int, [H, C.c_int, C.c_void_p, C.c_size_t]),
    "b200s_mat_info": (C.c_int, [H, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "b200s_device_ptr": (C.c_int, [H, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "b200s_convert_raw_to_mono": (C.c_int, [H, C.c_int]),
    "b200s_convert_raw_to_color": (C.c_int, [H, C.c_int]),
    "b200s_rectify": (C.c_int, [H, C.c_int, C.c_int, C.c_int]),
    "b200s_compute_disparity": (C.c_int, [H, C.c_int, C.c_int, C.c_int]),
    "b200s_compute_disparity_cuda_compat": (C.c_int, [H, C.c_int, C.c_int, C.c_int]),
    "b200s_filter_speckles": (C.c_int, [H, C.c_int]),
    "b200s_filter_speckles_host": (C.c_int, [H, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_int]),
    "b200s_compute_disparity_image": (C.c_int, [H, C.c_int, C.c_int]),
    "b200s_project_to_3d": (C.c_int, [H, C.c_int, C.c_int]),
    "b200s_wait": (C.c_int, [H, C.c_int]),
    "b200s_pack_image": (C.c_int, [H, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "b200s_pack_disparity": (C.c_int, [H, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(DisparityMeta)]),
    "b200s_pack_pointcloud2": (C.c_int, [H, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(Pc2Meta)]),
    "b200s_configure_slots": (C.c_int, [H, C.c_int, C.c_int, C.c_int]),
    "b200s_process_pair_async": (C.c_int, [H, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(FrameIO)]),
    "b200s_wait_slot": (C.c_int, [H, C.c_int]),
    "b200s_slot_device_ptr": (C.c_int, [H, C.c_int, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "b200s_process_pair": (C.c_int, [H, C.c_void_p, C.c_void_p, C.POINTER(FrameIO)]),
    "b200s_batch_begin": (C.c_int, [H]),
    "b200s_batch_end": (C.c_int, [H, C.POINTER(C.c_float)]),
    "b200s_kernel_launches": (C.c_uint64, [H]),
    "b200s_last_bm_time": (C.c_int, [H, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_double)]),
    "b200s_enable_timing": (C.c_int, [H, C.c_int]),
    "b200s_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "b200s_host_free": (C.c_int, [C.c_void_p]),
    "b200s_pool_create": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "b200s_pool_destroy": (C.c_int, [C.c_void_p]),
    "b200s_pool_size": (C.c_int, [C.c_void_p]),
    "b200s_pool_handle": (C.c_void_p, [C.c_void_p, C.c_int]),
    "b200s_pool_last_error_string": (C.c_char_p, [C.c_void_p]),
    "b200s_pool_set_calibration": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200s_pool_set_params": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200s_pool_submit": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "b200s_pool_wait": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "b200s_pool_wait_all": (C.c_int, [C.c_void_p]),
    "b200s_poll_slot": (C.c_int, [H, C.c_int, C.POINTER(C.c_int)]),
    "b200s_set_graph_mode": (C.c_int, [H, C.c_int]),
    "b200s_graph_replays": (C.c_uint64, [H]),
    "b200s_int_peak": (C.c_int, [H, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "b200s_convert_color": (C.c_int, [H, C.c_int, C.c_int, C.c_char_p, C.c_char_p]),
    "b200s_pack_image_async": (C.c_int, [H, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), DONE_FN, C.c_void_p]),
    "b200s_pack_disparity_async": (C.c_int, [H, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(DisparityMeta), DONE_FN, C.c_void_p]),
    "b200s_pack_pointcloud2_async": (C.c_int, [H, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(Pc2Meta), DONE_FN, C.c_void_p]),
    "b200s_set_pack_mode": (C.c_int, [H, C.c_int]),
    "b200s_configure_slots_batched": (C.c_int, [H, C.c_int, C.c_int, C.c_int, C.c_int]),
    "b200s_process_batch_async": (C.c_int, [H, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(FrameIO)]),
    "b200s_slot_frame_device_ptr": (C.c_int, [H, C.c_int, C.c_int, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "b200s_last_stage_times": (C.c_int, [H, C.c_int, C.POINTER(C.c_float)]),
    "b200s_host_alloc_mode": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t, C.c_int]),
    "b200s_copy_probe_ex": (C.c_int, [C.c_int, C.c_size_t, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "b200s_copy_probe": (C.c_int, [C.c_int, C.c_size_t, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "b200s_mat_stats": (C.c_int, [H, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]),
}

_lib = None


class B200StereoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (ERROR_NAMES.get(code, "B200S_E?"), code, msg))
        self.code = code


def load():
    """Loads the CUDA library.  There is no CPU fallback: if the extension was not built this raises."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libb200stereo.so is missing (%s): build it with `python -m ros_gpu_stereo_processor_b200.build`; "
                          "this package has no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
