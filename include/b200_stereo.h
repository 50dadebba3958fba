/*
 * b200_stereo.h -- C ABI of libb200stereo.so: the B200-native (sm_100a) stereo hot path behind the
 * method surface of gpuimageproc::GpuStereoProcessor (reference: include/gpuimageproc/GPUStereoProcessor.h:63-126).
 *
 * Plain C: opaque handle, POD structs, pointers + sizes, int return codes (0 = OK, negative = b200s_error).
 * Nothing here throws or aborts; the text of the last failure is available from b200s_last_error_string().
 * One handle = one camera pair on one GPU; a handle is thread-compatible (one caller at a time), exactly like
 * the reference object (include/gpuimageproc/StereoProcessor.h:92, src/StereoProcessor.cpp:160-161).
 *
 * There is NO CPU fallback: every entry point that computes does so with the CUDA kernels in this library
 * and fails with B200S_ECUDA when no usable device exists.
 */
#ifndef B200_STEREO_H
#define B200_STEREO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200s_handle b200s_handle;

typedef enum {
    B200S_OK = 0,
    B200S_EINVAL = -1,       /* bad argument / parameter rejected (cv::StereoBM would raise cv::Exception)      */
    B200S_ENOTINIT = -2,     /* stereo model not initialised (reference: assert(model_.initialized()))          */
    B200S_ECUDA = -3,        /* CUDA runtime / kernel failure, or no device                                     */
    B200S_ENOMEM = -4,
    B200S_EUNSUPPORTED = -5, /* encoding / interpolation / type outside the hot path (no CPU fallback offered)  */
    B200S_EIO = -6,          /* calibration file unreadable / malformed                                         */
    B200S_ENOBUF = -7        /* named buffer is empty or has the wrong type                                     */
} b200s_error;

/* Buffer ids: verbatim bit layout of enum GpuMatSource (include/gpuimageproc/GPUStereoProcessor.h:21-57).
 * id = one B200S_SRC_* flag | one side bit.  The side bit selects the stream (src/GPUStereoProcessor.cpp:190-200). */
enum {
    B200S_SIDE_L = 1 << 0,
    B200S_SIDE_R = 1 << 1,
    B200S_SIDE_MASK = 3,
    B200S_SRC_RAW = 1 << 2,
    B200S_SRC_MONO = 1 << 3,          /* CV_8UC1                                                         */
    B200S_SRC_COLOR = 1 << 4,         /* CV_8UC3 BGR                                                     */
    B200S_SRC_RECT_MONO = 1 << 5,
    B200S_SRC_RECT_COLOR = 1 << 6,
    B200S_SRC_DISPARITY = 1 << 7,     /* CV_16SC1 fixed point x16, cv::StereoBM semantics    */
    B200S_SRC_DISPARITY_32F = 1 << 8, /* CV_32FC1 = d16/16 - (cx_l - cx_r)  (GPUStereoProcessor.cpp:320) */
    B200S_SRC_DISPARITY_IMG = 1 << 9, /* CV_8UC4 BGRA colour-coded disparity                             */
    B200S_SRC_POINTS2 = 1 << 10       /* CV_32FC3 xyz                                                    */
};

/* element types use OpenCV's numeric codes so a cv::Mat::type() can be passed through unchanged */
enum { B200S_8UC1 = 0, B200S_16SC1 = 3, B200S_32FC1 = 5, B200S_8UC3 = 16, B200S_32FC3 = 21, B200S_8UC4 = 24 };

enum { B200S_INTER_NEAREST = 0, B200S_INTER_LINEAR = 1 }; /* cv::InterpolationFlags values */

/* sensor_msgs/CameraInfo subset consumed by initStereoModel (src/GPUStereoProcessor.cpp:41-53) */
typedef struct {
    int width, height;
    double K[9];
    double D[8]; /* k1 k2 p1 p2 k3 k4 k5 k6 (plumb_bob uses the first 5, rational_polynomial all 8) */
    int n_D;
    double R[9];
    double P[12];
} b200s_caminfo;

/* cv::StereoBM state + the speckle filter (cfg/GPU.cfg:12-35; setters src/GPUStereoProcessor.cpp:389-419).
 * Validation is cv::StereoBM's: invalid values are rejected with B200S_EINVAL, never clamped. */
typedef struct {
    int pre_filter_type; /* 0 = NORMALIZED_RESPONSE, 1 = XSOBEL   (GPU.cfg `xsobel`)                       */
    int pre_filter_size; /* odd, 5..255                                                                    */
    int pre_filter_cap;  /* 1..63                                                                          */
    int block_size;      /* odd, 5..255, < min(W,H)            (GPU.cfg `correlation_window_size`)         */
    int min_disparity;   /*                                    (GPU.cfg `disparity_min`)                   */
    int num_disparities; /* > 0, multiple of 16                (GPU.cfg `disparity_range`)                 */
    int texture_threshold;   /* >= 0                           (GPU.cfg `texture_threshold`)               */
    int uniqueness_ratio;    /* >= 0                                                                       */
    int speckle_window_size; /* 0 = off                        (GPU.cfg `max_speckle_size`)                */
    int speckle_range;       /* raw x16 units, as cv::StereoBM (GPU.cfg `max_speckle_diff` * 16)           */
    int disp12_max_diff;     /* -1 = off                                                                   */
    int refine_disparity;    /* fork-only switch of the reference (setRefineDisparity); accepted, no effect */
} b200s_params;

/* stereo_msgs/DisparityImage metadata (src/GpuSenderDisparity.cpp:18-48, intended stereo_image_proc semantics) */
typedef struct {
    int width, height, step; /* 32FC1 payload: step = width*4 */
    float f, T, min_disparity, max_disparity, delta_d;
    int valid_x_offset, valid_y_offset, valid_width, valid_height;
} b200s_disparity_meta;

/* sensor_msgs/PointCloud2 metadata (src/GpuSenderPc2.cpp:25-34): fields x@0 y@4 z@8 rgb@16, all FLOAT32 */
typedef struct {
    int width, height, point_step, row_step;
    int is_bigendian, is_dense;
    int off_x, off_y, off_z, off_rgb;
} b200s_pc2_meta;

/* ---- lifecycle ----------------------------------------------------------------------------------------------- */
/* GpuStereoProcessor::GpuStereoProcessor()  (src/GPUStereoProcessor.cpp:12-39).  device = CUDA ordinal.
 * Initial parameters are the reference constructor's: numDisparities 48, blockSize 19, preFilterSize 5,
 * PREFILTER_XSOBEL (:27-30), texture threshold 3, uniqueness 0, disp12MaxDiff 0 (copied from the cuda matcher's getters).
 * Every entry point makes the handle's device current for the call and restores the caller's current device. */
int b200s_create(int device, b200s_handle** out);
int b200s_destroy(b200s_handle* h);
const char* b200s_last_error_string(const b200s_handle* h);
const char* b200s_version(void);

/* ---- calibration: initStereoModel (src/GPUStereoProcessor.cpp:41-61), isStereoModelInitialised (:63) -------- */
int b200s_set_calibration(b200s_handle* h, const b200s_caminfo* left, const b200s_caminfo* right);
int b200s_load_calibration_files(b200s_handle* h, const char* left_yaml, const char* right_yaml);
int b200s_is_model_initialised(const b200s_handle* h);
/* reads back the stereo model: Q (16 doubles, row major), baseline, fx of the right camera, cx_l - cx_r */
int b200s_get_model(const b200s_handle* h, double* Q16, double* baseline, double* fx, double* cx_minus_cxr);

/* ---- parameters: setPreFilterType/.../setMaxSpeckleDiff (src/GPUStereoProcessor.cpp:202-208,389-419) -------- */
int b200s_default_params(b200s_params* p);                        /* cv::StereoBM defaults */
int b200s_set_params(b200s_handle* h, const b200s_params* p);
int b200s_get_params(const b200s_handle* h, b200s_params* p);
/* rectification mode: 0 = cached fixed-point map (built once per calibration by the GPU), 1 = map evaluated
 * on the fly in FP64 inside the rectify kernel.  Both are bit-identical. */
int b200s_set_rectify_mode(b200s_handle* h, int on_the_fly);

/* ---- named buffers: uploadMat / downloadMat (src/GPUStereoProcessor.cpp:89-117) ----------------------------- */
int b200s_upload(b200s_handle* h, int mat_id, const void* data, int rows, int cols, int type, size_t step,
                 const char* encoding);
int b200s_download(b200s_handle* h, int mat_id, void* dst, size_t dst_step); /* syncs that side's stream */
int b200s_mat_info(const b200s_handle* h, int mat_id, int* rows, int* cols, int* type);
/* device address of a named buffer (tightly packed rows); for callers that keep data in HBM (benchmarks, torch) */
int b200s_device_ptr(b200s_handle* h, int mat_id, void** dptr, size_t* bytes);

/* ---- the chain, one call per reference method ---------------------------------------------------------------- */
/* convertRawToMono / convertRawToColor (src/GPUStereoProcessor.cpp:65-88): side = B200S_SIDE_L or _R.
 * Supported raw encodings: mono8, bgr8, rgb8 (others: B200S_EUNSUPPORTED). */
int b200s_convert_raw_to_mono(b200s_handle* h, int side);
int b200s_convert_raw_to_color(b200s_handle* h, int side);
/* convertColor(src, dst, src_encoding, dst_encoding) (src/GPUStereoProcessor.cpp:119-172): mono8 / bgr8 / rgb8 sources,
 * mono8 / bgr8 destinations; anything else B200S_EUNSUPPORTED (the reference throws on unknown encodings) */
int b200s_convert_color(b200s_handle* h, int src_id, int dst_id, const char* src_encoding, const char* dst_encoding);
/* rectifyImage (src/GPUStereoProcessor.cpp:236-250); CPU-semantics result (bit-exact to cv::remap fixed point) */
int b200s_rectify(b200s_handle* h, int src_id, int dst_id, int interpolation);
/* computeDisparity (src/GPUStereoProcessor.cpp:264-303) with cv::StereoBM semantics -> CV_16SC1 x16 in disp_id;
 * also fills the matching DISPARITY_32F buffer of the same side. Includes validate + ROI mask + speckle exactly
 * as cv::StereoBM::compute does. */
int b200s_compute_disparity(b200s_handle* h, int left_id, int right_id, int disp_id);
/* "cuda-compat" mode: the bytes the reference's GPU matcher produces (block_matcher_gpu_->compute = cv::cuda::StereoBM,
 * src/GPUStereoProcessor.cpp:283; known answer test_data/aloe-disp.png): SSD, integer disparity, CV_8UC1 in disp_id,
 * 0 = invalid.  Uses num_disparities (multiple of 8, <= 256), block_size (3..51), pre_filter_type (1 = x-Sobel
 * prefilter, else none), pre_filter_cap, texture_threshold (avergeTexThreshold).  minDisparity, uniqueness, disp12 and
 * the speckle parameters have no effect there, exactly like the upstream setters. */
int b200s_compute_disparity_cuda_compat(b200s_handle* h, int left_id, int right_id, int disp_id);
/* filterSpeckles(GpuMatSource) (src/GPUStereoProcessor.cpp:356-385), in place.  CV_16SC1 plane: newVal = FILTERED,
 * maxSize = speckle_window_size, maxDiff = speckle_range (raw x16 units, as cv::StereoBM).  CV_8UC1 plane (cuda-compat):
 * the reference's own flow, newVal = 0 and maxDiff = speckle_range / 16 in integer disparities. */
int b200s_filter_speckles(b200s_handle* h, int disp_id);
/* stand-alone cv::filterSpeckles on a host CV_16SC1 plane (filterSpeckles(InputOutputArray), :367-385) */
int b200s_filter_speckles_host(b200s_handle* h, int16_t* img, int rows, int cols, size_t step, int new_val,
                               int max_size, int max_diff);
/* computeDisparityImage (src/GPUStereoProcessor.cpp:323-330): colour-coded BGRA8 */
int b200s_compute_disparity_image(b200s_handle* h, int disp_id, int img_id);
/* projectDisparityTo3DPoints (src/GPUStereoProcessor.cpp:332-346): CV_32FC3, missing -> Z = 10000.  disp_id may be the
 * DISPARITY or, like the reference's call (test/UTest.cpp:378), the DISPARITY_32F id of a side: both read that side's
 * fixed-point plane. */
int b200s_project_to_3d(b200s_handle* h, int disp_id, int points_id);
/* waitForStream / waitForAllStreams (src/GPUStereoProcessor.cpp:348-354); side 0 = all */
int b200s_wait(b200s_handle* h, int side);

/* ---- message payload packing (replaces GpuSender*::fillInData, src/GpuSender{Image,Disparity,Pc2}.cpp) ------ */
/* Synchronous forms: return with the payload complete in dst.  b200s_pack_pointcloud2 accepts the DISPARITY id or, like
 * the reference (src/StereoProcessor.cpp:281, test/UTest.cpp:382), the POINTS2 id of the side.  When dst is page-locked
 * host memory (b200s_host_alloc, cudaHostAlloc, cudaHostRegister) the float-disparity and PointCloud2 kernels store their
 * records straight into it ("pack straight into pinned host buffers"); pageable dst goes through a double-buffered pinned
 * staging copy. */
int b200s_pack_image(b200s_handle* h, int mat_id, void* dst, size_t cap_bytes, int* rows, int* cols, int* step);
int b200s_pack_disparity(b200s_handle* h, int disp_id, void* dst, size_t cap_bytes, b200s_disparity_meta* meta);
int b200s_pack_pointcloud2(b200s_handle* h, int disp_id, int color_id, void* dst, size_t cap_bytes,
                           b200s_pc2_meta* meta);
/* Asynchronous forms = the reference's senders (src/GpuSenderIfc.cpp:13-26: publish from a stream callback): the work
 * is enqueued on the side's stream and `done(user, status)` runs on the CUDA runtime's callback thread once the payload
 * is complete in dst.  No CUDA / b200s calls inside `done`.  meta is filled before the call returns.  done = NULL makes
 * the call synchronous. */
typedef void (*b200s_done_fn)(void* user, int status);
int b200s_pack_image_async(b200s_handle* h, int mat_id, void* dst, size_t cap_bytes, int* rows, int* cols, int* step,
                           b200s_done_fn done, void* user);
int b200s_pack_disparity_async(b200s_handle* h, int disp_id, void* dst, size_t cap_bytes, b200s_disparity_meta* meta,
                               b200s_done_fn done, void* user);
int b200s_pack_pointcloud2_async(b200s_handle* h, int disp_id, int color_id, void* dst, size_t cap_bytes,
                                 b200s_pc2_meta* meta, b200s_done_fn done, void* user);
/* 1 (default): pack kernels write straight into page-locked destinations; 0: always pack into HBM and copy (also
 * environment B200S_PACK_DIRECT=0).  The bytes are identical. */
int b200s_set_pack_mode(b200s_handle* h, int direct);

/* ---- fused frame path: the whole StereoProcessor::imageCb chain (src/StereoProcessor.cpp:157-298) ----------- */
enum {
    B200S_OUT_RECT_L = 1 << 0,      /* u8  H*W            */
    B200S_OUT_RECT_R = 1 << 1,      /* u8  H*W            */
    B200S_OUT_DISPARITY16 = 1 << 2, /* s16 H*W            */
    B200S_OUT_DISPARITY32F = 1 << 3,/* f32 H*W            */
    B200S_OUT_POINTCLOUD2 = 1 << 4, /* 32 B * H*W         */
    B200S_OUT_POINTS_XYZ = 1 << 5,  /* f32 3 * H*W        */
    B200S_OUT_RECT_COLOR_L = 1 << 6 /* u8  3 * H*W (BGR); needs color_left                                  */
};
enum { B200S_COLOR_NONE = 0, B200S_COLOR_BGR8 = 1, B200S_COLOR_RGB8 = 2 };
typedef struct {
    uint32_t want;          /* which products to compute (B200S_OUT_*)                                      */
    int rectify;            /* 1: inputs are raw images, rectify first; 0: inputs are already rectified     */
    int inputs_on_device;   /* 1: left/right/color_left are device pointers                                 */
    int outputs_on_device;  /* 1: the destination pointers below are device pointers                        */
    void* rect_left;        /* destinations; NULL = leave the product in the slot's device buffer           */
    void* rect_right;
    void* disparity16;
    void* disparity32f;
    void* pointcloud2;
    void* points_xyz;
    /* colour camera (src/StereoProcessor.cpp:201-217,239-256: L_RECT_COLOR feeds enqueueSendPoints): optional raw LEFT
     * colour image, 8UC3 of the slot size.  It is rectified like the mono image and colours the PointCloud2 records; when
     * the `left` argument of the call is NULL the matcher's grey image is derived from it (convertRawToMono). */
    const void* color_left;
    int color_encoding;     /* B200S_COLOR_BGR8 / B200S_COLOR_RGB8 when color_left is set                   */
    int rows, cols;         /* size of the caller's images; 0 = unchecked, otherwise must equal the slot size */
    void* rect_color_left;  /* destination of B200S_OUT_RECT_COLOR_L                                        */
} b200s_frame_io;

/* Frame slots: independent stream + device buffers each, so that copy-in, compute and copy-out of consecutive
 * frames overlap.  b200s_process_pair_async enqueues one frame on a slot and returns; b200s_wait_slot blocks until
 * that slot's outputs are complete.  Image size = calibration size (or rows/cols when no calibration is needed). */
int b200s_configure_slots(b200s_handle* h, int n_slots, int rows, int cols);
int b200s_process_pair_async(b200s_handle* h, int slot, const void* left, const void* right, const b200s_frame_io* io);
/* Batches ("batched stereo stream"): a slot can hold up to frames_per_slot (<= 32) frames and process them together --
 * every kernel of the chain takes the whole batch in one launch (frame = a grid dimension), so small images still fill
 * all SMs and the matcher works on tall row bands.  left / right / ios are arrays of n_frames entries; all frames of a
 * batch must ask for the same products and flags, each has its own destinations.  Results are identical to n_frames
 * single calls. */
int b200s_configure_slots_batched(b200s_handle* h, int n_slots, int rows, int cols, int frames_per_slot);
int b200s_process_batch_async(b200s_handle* h, int slot, int n_frames, const void* const* left, const void* const* right,
                              const b200s_frame_io* ios);
int b200s_wait_slot(b200s_handle* h, int slot);
/* non-blocking: *done = 1 when the slot's last frame (outputs included) is complete; the polled counterpart of the
 * reference's stream callback that publishes a message (src/GpuSenderIfc.cpp:13-26) */
int b200s_poll_slot(b200s_handle* h, int slot, int* done);
int b200s_slot_device_ptr(b200s_handle* h, int slot, uint32_t which /* one B200S_OUT_* */, void** dptr, size_t* bytes);
int b200s_slot_frame_device_ptr(b200s_handle* h, int slot, int frame, uint32_t which, void** dptr, size_t* bytes);
/* synchronous convenience: slot 0, process + wait */
int b200s_process_pair(b200s_handle* h, const void* left, const void* right, const b200s_frame_io* io);
/* The frame chain of a slot is captured into a CUDA graph the second time it runs with the same parameters, products
 * and output addresses, and replayed with one cudaGraphLaunch afterwards (inputs are copied into the slot first); a
 * slot caches a few graphs, so alternating destination buffers keep replaying.
 * on = 0 launches every kernel individually (also: environment B200S_GRAPH=0).  Results are identical. */
int b200s_set_graph_mode(b200s_handle* h, int on);
uint64_t b200s_graph_replays(const b200s_handle* h);

/* ---- several GPUs inside one process (the nodelet case): one handle per GPU, frames sharded round-robin ------------ */
/* Independent stereo frames need no exchange between GPUs: frame k runs on GPU k mod N in slot
 * (k div N) mod slots_per_gpu.  Every call only enqueues work, so a single host thread (the reference's one callback
 * thread, src/StereoProcessor.h:92) keeps all GPUs busy.  devices = NULL selects ordinals 0..n_gpus-1. */
typedef struct b200s_pool b200s_pool;
int b200s_pool_create(int n_gpus, const int* devices, int slots_per_gpu, int rows, int cols, b200s_pool** out);
int b200s_pool_destroy(b200s_pool* p);
int b200s_pool_size(const b200s_pool* p);
b200s_handle* b200s_pool_handle(b200s_pool* p, int gpu);                       /* for per-GPU calls of the API above */
const char* b200s_pool_last_error_string(const b200s_pool* p);
int b200s_pool_set_calibration(b200s_pool* p, const b200s_caminfo* left, const b200s_caminfo* right);   /* all GPUs */
int b200s_pool_set_params(b200s_pool* p, const b200s_params* prm);                                      /* all GPUs */
/* waits for the previous frame of the chosen slot, then enqueues this one; *gpu / *slot (optional) tell where it went.
 * left/right/io follow b200s_process_pair_async; device pointers must belong to that GPU. */
int b200s_pool_submit(b200s_pool* p, uint64_t frame_index, const void* left, const void* right, const b200s_frame_io* io,
                      int* gpu, int* slot);
int b200s_pool_wait(b200s_pool* p, int gpu, int slot);
int b200s_pool_wait_all(b200s_pool* p);

/* Device-side timing of a batch of frames spread over the slots (CUDA events on the slot streams): begin() syncs
 * the device, records a start event and makes every slot stream wait on it; end() records one event per slot
 * stream, waits for all of them and returns the longest start->end span in ms. */
int b200s_batch_begin(b200s_handle* h);
int b200s_batch_end(b200s_handle* h, float* ms);

/* ---- instrumentation ------------------------------------------------------------------------------------------ */
/* number of kernels this library has launched on this handle since creation (bench.py's gpu_launches) */
uint64_t b200s_kernel_launches(const b200s_handle* h);
/* device time in ms of the most recent block-matching kernel sequence on slot `slot` (CUDA events on its stream);
 * valid after b200s_wait_slot.  evals_effective receives the (pixel, disparity) evaluations it performed. */
int b200s_last_bm_time(b200s_handle* h, int slot, float* ms, double* evals_effective);
int b200s_enable_timing(b200s_handle* h, int on);
/* device time of each stage of the slot's last (timed) frame chain: rectify + prefilter, matcher, validate + speckle,
 * disparity -> float, reproject + PointCloud2 pack */
enum { B200S_STAGE_RECTIFY = 0, B200S_STAGE_MATCH = 1, B200S_STAGE_POST = 2, B200S_STAGE_TOFLOAT = 3, B200S_STAGE_PACK = 4,
       B200S_STAGE_COUNT = 5 };
int b200s_last_stage_times(b200s_handle* h, int slot, float* ms /* [B200S_STAGE_COUNT] */);
/* printStats (src/GPUStereoProcessor.cpp:421-435): per-channel min / max / mean of a named buffer, reduced on the GPU.
 * mn / mx / mean receive one value per channel (up to 4); *channels the channel count.  Syncs that side's stream. */
int b200s_mat_stats(b200s_handle* h, int mat_id, double* mn, double* mx, double* mean, int* channels);
/* Bare copy probe (no kernels): streams device->host copies of `bytes` round-robin into four page-locked host buffers on
 * two streams for about `seconds` and returns the achieved GB/s -- the ceiling of the end-to-end path on this host, per
 * GPU; run it from one process per GPU at the same time for the aggregate (tools/d2h_probe.py).  host_mode: 0 =
 * cudaHostAlloc default, 1 = cudaHostAllocWriteCombined, 2 = malloc + cudaHostRegister.  with_h2d also streams a
 * 4 MB host->device copy per D2H copy (the raw pair going in). */
int b200s_copy_probe(int device, size_t bytes, double seconds, int host_mode, int with_h2d, double* d2h_gbs);
/* same with n_streams (1..8) copy streams and n_buffers (1..16) host buffers instead of 2 and 4 */
int b200s_copy_probe_ex(int device, size_t bytes, double seconds, int host_mode, int with_h2d, int n_streams, int n_buffers,
                        double* d2h_gbs);
/* pinned host memory helpers (cudaHostAlloc / cudaFreeHost) for callers without a CUDA runtime of their own */
int b200s_host_alloc(void** p, size_t bytes);
/* mode 0 = cudaHostAllocDefault, 1 = cudaHostAllocWriteCombined (device writes bypass the CPU caches; CPU reads are slow) */
int b200s_host_alloc_mode(void** p, size_t bytes, int mode);
int b200s_host_free(void* p);
/* integer-ALU micro-benchmark used for the roofline denominator: runs `which` (0 IADD3, 1 VABSDIFF4, 2 VIADD.16x2,
 * 3 VIMNMX.U16x2, 4 IMAD, 5 PRMT, 6 LOP3, 7 IADD3+IMAD mixed) and returns lane-ops per second */
int b200s_int_peak(b200s_handle* h, int which, double* lane_ops_per_s, double* sm_clock_mhz);

#ifdef __cplusplus
}
#endif
#endif /* B200_STEREO_H */
