"""Host-side mirror of gpuimageproc::GpuStereoProcessor over the C ABI (include/b200_stereo.h).

Same method names, argument meaning and error behaviour as the reference class
(include/gpuimageproc/GPUStereoProcessor.h:63-126, src/GPUStereoProcessor.cpp); cv::Mat becomes a numpy array,
sensor_msgs/CameraInfo becomes a dict (or a camera_info yaml path), ROS messages become plain dict payloads.
All computation happens in libb200stereo.so on the GPU; nothing here computes on the CPU.
"""
import ctypes as C

import numpy as np

from . import _capi as capi
from ._capi import (SIDE_L, SIDE_R, SRC_RAW, SRC_MONO, SRC_COLOR, SRC_RECT_MONO, SRC_RECT_COLOR, SRC_DISPARITY,  # noqa: F401
                    SRC_DISPARITY_32F, SRC_DISPARITY_IMG, SRC_POINTS2, INTER_LINEAR, INTER_NEAREST)

# GpuMatSource composite ids (GPUStereoProcessor.h:35-56)
GPU_MAT_SIDE_L, GPU_MAT_SIDE_R = SIDE_L, SIDE_R
for _n, _v in dict(RAW=SRC_RAW, MONO=SRC_MONO, COLOR=SRC_COLOR, RECT_MONO=SRC_RECT_MONO, RECT_COLOR=SRC_RECT_COLOR,
                   DISPARITY=SRC_DISPARITY, DISPARITY_32F=SRC_DISPARITY_32F, DISPARITY_IMG=SRC_DISPARITY_IMG,
                   POINTS2=SRC_POINTS2).items():
    globals()["GPU_MAT_SRC_" + _n] = _v
    globals()["GPU_MAT_SRC_L_" + _n] = _v | SIDE_L
    globals()["GPU_MAT_SRC_R_" + _n] = _v | SIDE_R

PREFILTER_NORMALIZED_RESPONSE, PREFILTER_XSOBEL = 0, 1

_NP_OF_TYPE = {capi.T_8UC1: (np.uint8, 1), capi.T_16SC1: (np.int16, 1), capi.T_32FC1: (np.float32, 1),
               capi.T_8UC3: (np.uint8, 3), capi.T_32FC3: (np.float32, 3), capi.T_8UC4: (np.uint8, 4)}


def _type_of(a):
    ch = 1 if a.ndim == 2 else a.shape[2]
    for t, (dt, c) in _NP_OF_TYPE.items():
        if a.dtype == dt and c == ch:
            return t
    raise capi.B200StereoError(capi.EUNSUPPORTED, "unsupported array dtype/channels %s x%d" % (a.dtype, ch))


def _fixed_point_plane_of(mat_source):
    """POINTS2 / DISPARITY_32F ids name the side whose CV_16SC1 plane the reprojection and pack kernels read."""
    m = int(mat_source)
    return (SRC_DISPARITY | (m & 3)) if (m & (SRC_POINTS2 | SRC_DISPARITY_32F)) else m


def _caminfo(d):
    ci = capi.CamInfo()
    ci.width, ci.height = int(d["width"]), int(d["height"])
    for name, n in (("K", 9), ("R", 9), ("P", 12)):
        v = np.asarray(d[name], np.float64).ravel()
        assert v.size == n, name
        getattr(ci, name)[:] = v.tolist()
    D = np.asarray(d.get("D", []), np.float64).ravel()[:8]
    ci.n_D = int(D.size)
    ci.D[:] = (D.tolist() + [0.0] * 8)[:8]
    return ci


class Sender(object):
    """Stand-in for GPUSenderImage/Disparity/Pc2 (src/GpuSender*.cpp): holds the packed message payload.  Like the
    reference's senders (src/GpuSenderIfc.cpp:13-26) an asynchronous one is completed -- and its publisher called -- from
    the CUDA stream callback once the payload has arrived in host memory."""

    def __init__(self, kind, message, pub=None, sent=True):
        self.kind, self.message, self._pub, self._sent = kind, message, pub, sent
        self._pinned = None      # (ptr, free function) of a pinned payload buffer owned by this sender
        self._cb = None          # keeps the ctypes callback alive until it has run

    def _complete(self, _user=None, _status=0):
        if self._pub is not None:
            self._pub(self.message)
        self._sent = True

    def wasDataSent(self):
        return self._sent

    def release(self):
        if self._pinned is not None:
            ptr, free = self._pinned
            self._pinned = None
            self.message = None
            free(ptr)


class GpuStereoProcessor(object):
    def __init__(self, device=0):
        self._lib = capi.load()
        self._h = capi.H()
        rc = self._lib.b200s_create(int(device), C.byref(self._h))
        if rc != 0:
            raise capi.B200StereoError(rc, "b200s_create failed (no usable CUDA device?)")
        self._p = capi.Params()
        self._ck(self._lib.b200s_get_params(self._h, C.byref(self._p)))
        self._senders = []
        self._slots = 0
        self._slot_shape = None
        self._keep = []

    # ---- plumbing -------------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise capi.B200StereoError(rc, self._lib.b200s_last_error_string(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200s_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _sync_params(self):
        self._ck(self._lib.b200s_set_params(self._h, C.byref(self._p)))

    # ---- calibration (src/GPUStereoProcessor.cpp:41-63) -----------------------------------------------
    def initStereoModel(self, left, right):
        if isinstance(left, str):
            self._ck(self._lib.b200s_load_calibration_files(self._h, left.encode(), right.encode()))
        else:
            l, r = _caminfo(left), _caminfo(right)
            self._ck(self._lib.b200s_set_calibration(self._h, C.byref(l), C.byref(r)))

    def isStereoModelInitialised(self):
        return bool(self._lib.b200s_is_model_initialised(self._h))

    def getModel(self):
        Q = (C.c_double * 16)()
        b, fx, cxd = C.c_double(), C.c_double(), C.c_double()
        self._ck(self._lib.b200s_get_model(self._h, Q, C.byref(b), C.byref(fx), C.byref(cxd)))
        return dict(Q=np.array(Q[:], np.float64).reshape(4, 4), baseline=b.value, fx=fx.value, cx_minus_cxr=cxd.value)

    def setRectifyOnTheFly(self, on):
        self._ck(self._lib.b200s_set_rectify_mode(self._h, int(bool(on))))

    # ---- buffers (src/GPUStereoProcessor.cpp:89-117) --------------------------------------------------
    def uploadMat(self, mat_source, cv_mat, encoding=""):
        a = np.asarray(cv_mat)
        t = _type_of(a)
        if not a.flags["C_CONTIGUOUS"]:
            a = np.ascontiguousarray(a)
        self._ck(self._lib.b200s_upload(self._h, int(mat_source), a.ctypes.data, a.shape[0], a.shape[1], t,
                                        a.strides[0], encoding.encode()))
        # the copy is asynchronous on that side's stream; keep the source alive until the next wait
        self._keep.append(a)

    def matInfo(self, mat_source):
        r, c, t = C.c_int(), C.c_int(), C.c_int()
        self._ck(self._lib.b200s_mat_info(self._h, int(mat_source), C.byref(r), C.byref(c), C.byref(t)))
        return r.value, c.value, t.value

    def downloadMat(self, mat_source, cv_mat=None):
        rows, cols, t = self.matInfo(mat_source)
        dt, ch = _NP_OF_TYPE[t]
        out = np.empty((rows, cols) if ch == 1 else (rows, cols, ch), dt) if cv_mat is None else cv_mat
        self._ck(self._lib.b200s_download(self._h, int(mat_source), out.ctypes.data, out.strides[0]))
        self._keep.clear()
        return out

    def convertRawToMono(self, side):
        self._ck(self._lib.b200s_convert_raw_to_mono(self._h, int(side)))

    def convertRawToColor(self, side):
        self._ck(self._lib.b200s_convert_raw_to_color(self._h, int(side)))

    # ---- the chain ------------------------------------------------------------------------------------
    def rectifyImage(self, source, dest, interpolation=INTER_LINEAR):
        self._ck(self._lib.b200s_rectify(self._h, int(source), int(dest), int(interpolation)))

    def _rectify_side(self, side, src, interpolation):
        src = np.asarray(src)
        raw_id = SRC_RAW | side
        dst_id = (SRC_RECT_MONO if src.ndim == 2 else SRC_RECT_COLOR) | side
        self.uploadMat(raw_id, src)
        self.rectifyImage(raw_id, dst_id, interpolation)
        return self.downloadMat(dst_id)

    def rectifyImageLeft(self, source, interpolation=INTER_LINEAR):
        return self._rectify_side(SIDE_L, source, interpolation)

    def rectifyImageRight(self, source, interpolation=INTER_LINEAR):
        return self._rectify_side(SIDE_R, source, interpolation)

    def computeDisparity(self, left, right, disparity=None):
        """computeDisparity(GpuMatSource, GpuMatSource, GpuMatSource) or computeDisparity(Mat, Mat) -> float32 Mat
        (src/GPUStereoProcessor.cpp:264-321)."""
        if isinstance(left, (int, np.integer)):
            self._sync_params()
            self._ck(self._lib.b200s_compute_disparity(self._h, int(left), int(right), int(disparity)))
            return None
        self.computeDisparityBare(left, right)
        return self.downloadMat(SRC_DISPARITY_32F | SIDE_L)

    def computeDisparityBare(self, left, right):
        """Matcher only, on host images: returns the CV_16SC1 x16 plane (src/GPUStereoProcessor.cpp:305-310)."""
        self.uploadMat(SRC_RECT_MONO | SIDE_L, left)
        self.uploadMat(SRC_RECT_MONO | SIDE_R, right)
        self._sync_params()
        self._ck(self._lib.b200s_compute_disparity(self._h, SRC_RECT_MONO | SIDE_L, SRC_RECT_MONO | SIDE_R,
                                                   SRC_DISPARITY | SIDE_L))
        return self.downloadMat(SRC_DISPARITY | SIDE_L)

    def computeDisparityCudaCompat(self, left, right, disparity=None):
        """cv::cuda::StereoBM compatibility mode = the bytes of the reference's GPU matcher
        (block_matcher_gpu_->compute, src/GPUStereoProcessor.cpp:283): CV_8UC1 integer disparity, 0 = invalid.
        Ids like computeDisparity, or two host images -> host u8 plane."""
        self._sync_params()
        if isinstance(left, (int, np.integer)):
            self._ck(self._lib.b200s_compute_disparity_cuda_compat(self._h, int(left), int(right), int(disparity)))
            return None
        self.uploadMat(SRC_RECT_MONO | SIDE_L, left)
        self.uploadMat(SRC_RECT_MONO | SIDE_R, right)
        self._ck(self._lib.b200s_compute_disparity_cuda_compat(self._h, SRC_RECT_MONO | SIDE_L, SRC_RECT_MONO | SIDE_R,
                                                               SRC_DISPARITY | SIDE_L))
        return self.downloadMat(SRC_DISPARITY | SIDE_L)

    def computeDisparityImage(self, disparity_src, disp_image_dest):
        self._sync_params()
        self._ck(self._lib.b200s_compute_disparity_image(self._h, int(disparity_src), int(disp_image_dest)))

    def projectDisparityTo3DPoints(self, disparity_src, points_src):
        """disparity_src: DISPARITY or, as the reference calls it (test/UTest.cpp:378), DISPARITY_32F of a side."""
        self._ck(self._lib.b200s_project_to_3d(self._h, int(disparity_src), int(points_src)))

    def filterSpeckles(self, disparity):
        """filterSpeckles(GpuMatSource) in place on the device, or filterSpeckles(array) on a host CV_16SC1 plane
        with newVal = FILTERED (src/GPUStereoProcessor.cpp:356-385)."""
        self._sync_params()
        if isinstance(disparity, (int, np.integer)):
            self._ck(self._lib.b200s_filter_speckles(self._h, int(disparity)))
            return None
        a = np.ascontiguousarray(disparity, np.int16)
        self._ck(self._lib.b200s_filter_speckles_host(self._h, a.ctypes.data, a.shape[0], a.shape[1], a.strides[0],
                                                      (self._p.min_disparity - 1) * 16, self._p.speckle_window_size,
                                                      self._p.speckle_range))
        return a

    def filterSpecklesRaw(self, img, new_val, max_size, max_diff):
        a = np.ascontiguousarray(img, np.int16).copy()
        self._ck(self._lib.b200s_filter_speckles_host(self._h, a.ctypes.data, a.shape[0], a.shape[1], a.strides[0],
                                                      int(new_val), int(max_size), int(max_diff)))
        return a

    def waitForStream(self, stream_source):
        self._ck(self._lib.b200s_wait(self._h, int(stream_source) & 3))
        self._keep.clear()

    def waitForAllStreams(self):
        self._ck(self._lib.b200s_wait(self._h, 0))
        self._keep.clear()

    def cleanSenders(self):
        """src/GPUStereoProcessor.cpp:228-234: drops the senders whose data went out (call after waitForAllStreams)."""
        keep = []
        for s in self._senders:
            if s.wasDataSent():
                s.release()
            else:
                keep.append(s)
        self._senders = keep

    # ---- senders (src/GPUStereoProcessor.cpp:210-234, src/GpuSender*.cpp) -----------------------------
    # asynchronous=True is the reference's behaviour: the call only enqueues work on the side's stream; the payload
    # lands in a pinned buffer owned by the sender and `pub(message)` is called from the stream callback
    # (src/GpuSenderIfc.cpp:13-26).  The default returns with the message complete, which the tests read directly.
    def _payload(self, nbytes, asynchronous):
        if not asynchronous:
            return np.empty(nbytes, np.uint8), None
        arr, ptr = self.hostAlloc(nbytes)
        return arr, (ptr, self.hostFree)

    def _finish_sender(self, snd, pinned, asynchronous, call):
        """call(done_fn) -> return code of the *_async entry point"""
        self._senders.append(snd)
        if not asynchronous:
            self._ck(call(capi.DONE_FN(0)))
            snd._complete()
            return snd
        snd._pinned = pinned
        snd._sent = False
        snd._cb = capi.DONE_FN(snd._complete)
        self._ck(call(snd._cb))
        return snd

    def enqueueSendImage(self, source, imagePattern=None, encoding="", pub=None, asynchronous=False):
        rows, cols, t = self.matInfo(source)
        dt, ch = _NP_OF_TYPE[t]
        data, pinned = self._payload(rows * cols * ch * np.dtype(dt).itemsize, asynchronous)
        r, c, s = C.c_int(), C.c_int(), C.c_int()
        msg = dict(header=imagePattern, height=rows, width=cols, step=cols * ch * np.dtype(dt).itemsize, encoding=encoding, data=data)
        snd = Sender("image", msg, pub)
        return self._finish_sender(snd, pinned, asynchronous, lambda cb: self._lib.b200s_pack_image_async(
            self._h, int(source), data.ctypes.data, data.size, C.byref(r), C.byref(c), C.byref(s), cb, None))

    def enqueueSendDisparity(self, source, imagePattern=None, pub=None, asynchronous=False):
        rows, cols, _ = self.matInfo(_fixed_point_plane_of(source))
        raw, pinned = self._payload(rows * cols * 4, asynchronous)
        data = raw.view(np.float32).reshape(rows, cols)
        meta = capi.DisparityMeta()
        self._sync_params()
        msg = dict(header=imagePattern, image=dict(height=rows, width=cols, step=cols * 4, encoding="32FC1", data=data))
        snd = Sender("disparity", msg, pub)

        def call(cb):
            rc = self._lib.b200s_pack_disparity_async(self._h, int(source), data.ctypes.data, data.nbytes, C.byref(meta), cb, None)
            msg.update(f=meta.f, T=meta.T, min_disparity=meta.min_disparity, max_disparity=meta.max_disparity, delta_d=meta.delta_d,
                       valid_window=dict(x_offset=meta.valid_x_offset, y_offset=meta.valid_y_offset, width=meta.valid_width,
                                         height=meta.valid_height))
            return rc
        return self._finish_sender(snd, pinned, asynchronous, call)

    def enqueueSendPoints(self, points_source, color_source, imagePattern=None, pub=None, asynchronous=False):
        """points_source: the reference passes its POINTS2 buffer (src/StereoProcessor.cpp:281); reprojection and packing
        are one kernel here, so that id -- or the DISPARITY id -- names the side whose fixed-point plane is read."""
        rows, cols, _ = self.matInfo(_fixed_point_plane_of(points_source))
        raw, pinned = self._payload(rows * cols * 32, asynchronous)
        data = raw.reshape(rows, cols, 32)
        meta = capi.Pc2Meta()
        msg = dict(header=imagePattern, height=rows, width=cols, data=data)
        snd = Sender("points2", msg, pub)

        def call(cb):
            rc = self._lib.b200s_pack_pointcloud2_async(self._h, int(points_source), int(color_source), data.ctypes.data, data.nbytes,
                                                        C.byref(meta), cb, None)
            fields = [dict(name="x", offset=meta.off_x, datatype=7, count=1), dict(name="y", offset=meta.off_y, datatype=7, count=1),
                      dict(name="z", offset=meta.off_z, datatype=7, count=1), dict(name="rgb", offset=meta.off_rgb, datatype=7, count=1)]
            msg.update(fields=fields, is_bigendian=bool(meta.is_bigendian), point_step=meta.point_step, row_step=meta.row_step,
                       is_dense=bool(meta.is_dense))
            return rc
        return self._finish_sender(snd, pinned, asynchronous, call)

    def convertColor(self, mat_source, mat_dst, src_encoding, dst_encoding):
        """src/GPUStereoProcessor.cpp:119-172 for the encodings on the hot path (mono8, bgr8, rgb8)."""
        self._ck(self._lib.b200s_convert_color(self._h, int(mat_source), int(mat_dst), src_encoding.encode(), dst_encoding.encode()))

    # ---- parameters (src/GPUStereoProcessor.cpp:202-208,389-419 + the cv::StereoBM ones GPU.cfg lacks) ---
    def setPreFilterType(self, filter_type): self._p.pre_filter_type = int(filter_type)
    def setPreFilterSize(self, v): self._p.pre_filter_size = int(v)
    def setPreFilterCap(self, v): self._p.pre_filter_cap = int(v)
    def setRefineDisparity(self, ref_disp): self._p.refine_disparity = int(bool(ref_disp))
    def setBlockSize(self, block_size): self._p.block_size = int(block_size)
    def setNumDisparities(self, numDisp): self._p.num_disparities = int(numDisp)
    def setMinDisparity(self, minDisp): self._p.min_disparity = int(minDisp)
    def setTextureThreshold(self, threshold): self._p.texture_threshold = int(threshold)
    def setUniquenessRatio(self, v): self._p.uniqueness_ratio = int(v)
    def setDisp12MaxDiff(self, v): self._p.disp12_max_diff = int(v)
    def setSpeckleRange(self, v): self._p.speckle_range = int(v)        # raw x16 units, as cv::StereoBM
    def getMaxSpeckleSize(self): return self._p.speckle_window_size
    def setMaxSpeckleSize(self, maxSpeckleSize): self._p.speckle_window_size = int(maxSpeckleSize)
    def getMaxSpeckleDiff(self): return self._p.speckle_range / 16.0
    def setMaxSpeckleDiff(self, maxSpeckleDiff): self._p.speckle_range = int(round(float(maxSpeckleDiff) * 16))  # integer-disparity units

    _PARAM_NAMES = dict(minDisparity="min_disparity", numDisparities="num_disparities", blockSize="block_size",
                        preFilterType="pre_filter_type", preFilterSize="pre_filter_size", preFilterCap="pre_filter_cap",
                        textureThreshold="texture_threshold", uniquenessRatio="uniqueness_ratio",
                        speckleWindowSize="speckle_window_size", speckleRange="speckle_range", disp12MaxDiff="disp12_max_diff")

    def setParams(self, **kw):
        """Bulk setter with cv::StereoBM names: numDisparities, blockSize, minDisparity, preFilterType, ..."""
        for k, v in kw.items():
            setattr(self._p, self._PARAM_NAMES[k], int(v))

    def getParams(self):
        return {n: getattr(self._p, n) for n, _ in self._p._fields_}

    def matStats(self, mat_source):
        """Per-channel (min, max, mean) of a named buffer, reduced on the GPU."""
        mn, mx, mean, ch = (C.c_double * 4)(), (C.c_double * 4)(), (C.c_double * 4)(), C.c_int()
        self._ck(self._lib.b200s_mat_stats(self._h, int(mat_source), mn, mx, mean, C.byref(ch)))
        return [(mn[i], mx[i], mean[i]) for i in range(ch.value)]

    def printStats(self, name, mat):
        """src/GPUStereoProcessor.cpp:421-435; mat is a GpuMatSource id (reduced on the GPU) or a host array (uploaded first)."""
        if not isinstance(mat, (int, np.integer)):
            self.uploadMat(SRC_DISPARITY_IMG | SIDE_R, np.asarray(mat))
            mat = SRC_DISPARITY_IMG | SIDE_R
        lines = ["ARRAY STATS:%s; channel:%d; min:%f; max:%f; mean:%f;" % (name, i, a, b, m) for i, (a, b, m) in enumerate(self.matStats(mat))]
        print("\n".join(lines))
        return lines

    # ---- fused frame path (StereoProcessor::imageCb chain, src/StereoProcessor.cpp:157-298) -----------
    def configureSlots(self, n_slots, rows, cols, frames_per_slot=1):
        self._ck(self._lib.b200s_configure_slots_batched(self._h, int(n_slots), int(rows), int(cols), int(frames_per_slot)))
        self._slots = n_slots
        self._slot_shape = (int(rows), int(cols))

    def processPairAsync(self, slot, left, right, io):
        self._sync_params()
        self._ck(self._lib.b200s_process_pair_async(self._h, int(slot), left, right, C.byref(io)))

    def processBatchAsync(self, slot, lefts, rights, ios):
        """One batch of frames on a slot (b200s_process_batch_async): lefts / rights are sequences of host or device
        addresses, ios a ctypes array (FrameIO * n) or a sequence of FrameIO."""
        n = len(lefts)
        la = (C.c_void_p * n)(*[int(p) if p else None for p in lefts])
        ra = (C.c_void_p * n)(*[int(p) for p in rights])
        if not isinstance(ios, C.Array):
            ios = (capi.FrameIO * n)(*ios)
        self._sync_params()
        self._ck(self._lib.b200s_process_batch_async(self._h, int(slot), n, la, ra, ios))

    def makeBatch(self, lefts, rights, ios):
        """Pre-built argument arrays of a batch for processBatchRaw (keeps the per-call host work minimal)."""
        n = len(rights)
        la = (C.c_void_p * n)(*[int(p) if p else None for p in lefts])
        ra = (C.c_void_p * n)(*[int(p) for p in rights])
        if not isinstance(ios, C.Array):
            ios = (capi.FrameIO * n)(*ios)
        return n, la, ra, ios

    def processBatchRaw(self, slot, batch):
        n, la, ra, ios = batch
        rc = self._lib.b200s_process_batch_async(self._h, slot, n, la, ra, ios)
        if rc != 0:
            self._ck(rc)

    def syncParams(self):
        self._sync_params()

    def setPackMode(self, direct):
        self._ck(self._lib.b200s_set_pack_mode(self._h, int(bool(direct))))

    def lastStageTimes(self, slot):
        ms = (C.c_float * len(capi.STAGE_NAMES))()
        self._ck(self._lib.b200s_last_stage_times(self._h, int(slot), ms))
        return dict(zip(capi.STAGE_NAMES, [float(v) for v in ms]))

    def slotFrameDevicePtr(self, slot, frame, which):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self._lib.b200s_slot_frame_device_ptr(self._h, int(slot), int(frame), int(which), C.byref(p), C.byref(n)))
        return p.value, n.value

    def waitSlot(self, slot):
        self._ck(self._lib.b200s_wait_slot(self._h, int(slot)))

    def slotDone(self, slot):
        """Non-blocking completion test of the slot's last frame."""
        d = C.c_int()
        self._ck(self._lib.b200s_poll_slot(self._h, int(slot), C.byref(d)))
        return bool(d.value)

    def processPair(self, left, right, rectify=True, want=("disparity16",), color=None, color_encoding="bgr8"):
        """Synchronous convenience on slot 0 with host arrays; returns a dict of numpy outputs.  color: optional raw
        left colour image (H, W, 3) that colours the point cloud; left=None derives the grey image from it."""
        R = np.ascontiguousarray(right, np.uint8)
        L = None if left is None else np.ascontiguousarray(left, np.uint8)
        Cimg = None if color is None else np.ascontiguousarray(color, np.uint8)
        rows, cols = R.shape
        for a in (L, Cimg):
            if a is not None and a.shape[:2] != (rows, cols):
                raise capi.B200StereoError(capi.EINVAL, "left / right / colour images differ in size")
        if self._slots == 0 or self._slot_shape != (rows, cols):
            self.configureSlots(max(self._slots, 1), rows, cols)     # the slots follow the image size
        io = capi.FrameIO()
        io.rectify = int(bool(rectify))
        io.rows, io.cols = rows, cols
        if Cimg is not None:
            io.color_left = Cimg.ctypes.data
            io.color_encoding = capi.COLOR_RGB8 if color_encoding == "rgb8" else capi.COLOR_BGR8
        out = {}
        spec = dict(rect_left=(capi.OUT_RECT_L, (rows, cols), np.uint8), rect_right=(capi.OUT_RECT_R, (rows, cols), np.uint8),
                    disparity16=(capi.OUT_DISPARITY16, (rows, cols), np.int16), disparity32f=(capi.OUT_DISPARITY32F, (rows, cols), np.float32),
                    pointcloud2=(capi.OUT_POINTCLOUD2, (rows, cols, 32), np.uint8), points_xyz=(capi.OUT_POINTS_XYZ, (rows, cols, 3), np.float32),
                    rect_color_left=(capi.OUT_RECT_COLOR_L, (rows, cols, 3), np.uint8))
        for name in want:
            bit, shape, dt = spec[name]
            io.want |= bit
            out[name] = np.empty(shape, dt)
            setattr(io, name, out[name].ctypes.data)
        self.processPairAsync(0, None if L is None else L.ctypes.data, R.ctypes.data, io)
        self.waitSlot(0)
        return out

    def batchBegin(self):
        self._ck(self._lib.b200s_batch_begin(self._h))

    def batchEnd(self):
        ms = C.c_float()
        self._ck(self._lib.b200s_batch_end(self._h, C.byref(ms)))
        return ms.value

    def hostAlloc(self, nbytes, write_combined=False):
        """Pinned host buffer as a numpy uint8 array (cudaHostAlloc)."""
        p = C.c_void_p()
        rc = self._lib.b200s_host_alloc_mode(C.byref(p), int(nbytes), int(bool(write_combined)))
        if rc != 0:
            raise capi.B200StereoError(rc, "cudaHostAlloc failed")
        buf = (C.c_uint8 * int(nbytes)).from_address(p.value)
        return np.frombuffer(buf, np.uint8), p.value

    def hostFree(self, ptr):
        self._lib.b200s_host_free(C.c_void_p(ptr))

    def slotDevicePtr(self, slot, which):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self._lib.b200s_slot_device_ptr(self._h, int(slot), int(which), C.byref(p), C.byref(n)))
        return p.value, n.value

    def devicePtr(self, mat_source):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self._lib.b200s_device_ptr(self._h, int(mat_source), C.byref(p), C.byref(n)))
        return p.value, n.value

    # ---- instrumentation ------------------------------------------------------------------------------
    def kernelLaunches(self):
        return int(self._lib.b200s_kernel_launches(self._h))

    def enableTiming(self, on=True):
        self._ck(self._lib.b200s_enable_timing(self._h, int(bool(on))))

    def lastBmTime(self, slot=-1):
        ms, ev = C.c_float(), C.c_double()
        self._ck(self._lib.b200s_last_bm_time(self._h, int(slot), C.byref(ms), C.byref(ev)))
        return ms.value, ev.value

    def setGraphMode(self, on):
        """CUDA-graph replay of the per-slot frame chain (default on); results are identical either way."""
        self._ck(self._lib.b200s_set_graph_mode(self._h, int(bool(on))))

    def graphReplays(self):
        return int(self._lib.b200s_graph_replays(self._h))

    def intPeak(self, which):
        ops, mhz = C.c_double(), C.c_double()
        self._ck(self._lib.b200s_int_peak(self._h, int(which), C.byref(ops), C.byref(mhz)))
        return ops.value, mhz.value


class GpuStereoPool(object):
    """Several GPUs inside one process (b200s_pool_*): one GpuStereoProcessor-equivalent handle per GPU, independent
    frames sharded round-robin (frame k -> GPU k mod N, slot (k // N) mod slots), no exchange between GPUs."""

    def __init__(self, n_gpus, rows, cols, slots_per_gpu=2, devices=None):
        self._lib = capi.load()
        self._p = C.c_void_p()
        dev = (C.c_int * n_gpus)(*devices) if devices is not None else None
        rc = self._lib.b200s_pool_create(int(n_gpus), dev, int(slots_per_gpu), int(rows), int(cols), C.byref(self._p))
        if rc != 0:
            raise capi.B200StereoError(rc, "b200s_pool_create failed")
        self.n_gpus, self.slots = int(n_gpus), int(slots_per_gpu)

    def _ck(self, rc):
        if rc != 0:
            raise capi.B200StereoError(rc, self._lib.b200s_pool_last_error_string(self._p).decode())

    def initStereoModel(self, left, right):
        l, r = _caminfo(left), _caminfo(right)
        self._ck(self._lib.b200s_pool_set_calibration(self._p, C.byref(l), C.byref(r)))

    def setParams(self, **kw):
        prm = capi.Params()
        self._lib.b200s_default_params(C.byref(prm))
        for k, v in kw.items():
            setattr(prm, GpuStereoProcessor._PARAM_NAMES[k], int(v))
        self._ck(self._lib.b200s_pool_set_params(self._p, C.byref(prm)))

    def submit(self, frame_index, left_ptr, right_ptr, io):
        g, s = C.c_int(), C.c_int()
        self._ck(self._lib.b200s_pool_submit(self._p, int(frame_index), left_ptr, right_ptr, C.byref(io), C.byref(g), C.byref(s)))
        return g.value, s.value

    def wait(self, gpu, slot):
        self._ck(self._lib.b200s_pool_wait(self._p, int(gpu), int(slot)))

    def waitAll(self):
        self._ck(self._lib.b200s_pool_wait_all(self._p))

    def close(self):
        if getattr(self, "_p", None):
            self._lib.b200s_pool_destroy(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
