// Internal launcher interface between the host pipeline (api.cu) and the sm_100a kernels.
// Every function enqueues work on `st` and returns the number of kernels it launched (>= 0) or a
// negative value when the launch configuration is impossible (caller maps it to B200S_E*).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200s {

struct CamModel {          // one camera: K, D and ir = inv(P[:, :3] * R) (FP64, host-computed)
    double fx, fy, cx, cy;
    double k1, k2, p1, p2, k3, k4, k5, k6;
    double ir[9];
};

struct BMConfig {          // cv::StereoBM state actually used by the matcher
    int minD, nd, wsz, cap, textureThreshold, uniquenessRatio, disp12MaxDiff;
};

// ---- batches ----------------------------------------------------------------------------------------
// Several equally sized frames can go through one launch (frame index = a grid dimension): frame f of a plane starts
// `stride` BYTES after frame f - 1.  nf = 1 with zero strides is the single-frame case of the named-buffer API.
// Input planes may instead be named by a device-resident table of per-frame addresses (`tab*`): the caller's own
// device images are read in place, and a captured CUDA graph stays valid when the addresses change (only the table
// is rewritten).
constexpr int MAX_BATCH = 32;
struct PtrList {           // per-frame destinations that are not one strided buffer (caller-owned pinned host memory)
    void* p[MAX_BATCH];
};

// ---- rectify.cu -------------------------------------------------------------------------------------
// Rectification map of cv::initUndistortRectifyMap in x32 fixed point, (sx, sy) = (rint(u*32), rint(v*32)):
//   MAP_DELTA16: 4 B/px, (sx - 32 x, sy - 32 y) as int16 pairs -- every practical calibration (|shift| < 1024 px)
//   MAP_ABS32:   8 B/px int2, used when a delta does not fit 16 bits
//   MAP_FLY:     no table, the map is evaluated in FP64 inside the consuming kernel
//   MAP_NONE:    identity (the source is already rectified)
enum MapMode { MAP_NONE = 0, MAP_ABS32 = 1, MAP_FLY = 2, MAP_DELTA16 = 3 };
// builds both table formats' content into `map` (MAP_ABS32: W*H int2, MAP_DELTA16: W*H short2); *overflow (device int,
// preset to 0 by this call) becomes 1 when MAP_DELTA16 was requested and some delta does not fit
int launch_build_map(const CamModel& cm, int W, int H, void* map, MapMode mode, int* overflow, cudaStream_t st);
// cv::remap INTER_LINEAR / BORDER_CONSTANT(0); ch = 1, 3 or 4 interleaved
int launch_remap(const uint8_t* src, int sW, int sH, int ch, const void* map, MapMode mode, const CamModel& cm,
                 uint8_t* dst, int W, int H, cudaStream_t st, int nf = 1, size_t src_stride = 0, size_t dst_stride = 0,
                 const uint8_t* const* src_tab = nullptr);
// fused (rectify +) x-Sobel prefilter: writes the rectified plane (mode != MAP_NONE) and the prefiltered plane of both
// sides in one launch (blockIdx.z = 2 * frame + side)
int launch_rectify_xsobel_pair(const uint8_t* srcL, const uint8_t* srcR, int sW, int sH, MapMode mode, const void* mapL,
                               const void* mapR, const CamModel& cmL, const CamModel& cmR, uint8_t* rectL, uint8_t* rectR,
                               uint8_t* preL, uint8_t* preR, size_t pre_pitch, int W, int H, int cap, cudaStream_t st,
                               int nf = 1, size_t src_stride = 0, size_t rect_stride = 0, size_t pre_stride = 0,
                               const uint8_t* const* tabL = nullptr, const uint8_t* const* tabR = nullptr);
// (rectify +) normalised-response prefilter of both sides in one tiled kernel (preFilterSize <= 21; returns 0 otherwise)
int launch_norm_prefilter_pair(const uint8_t* srcL, const uint8_t* srcR, int sW, int sH, MapMode mode, const void* mapL,
                               const void* mapR, const CamModel& cmL, const CamModel& cmR, uint8_t* rectL, uint8_t* rectR,
                               uint8_t* preL, uint8_t* preR, size_t pre_pitch, int W, int H, int ps, int cap, cudaStream_t st,
                               int nf = 1, size_t src_stride = 0, size_t rect_stride = 0, size_t pre_stride = 0,
                               const uint8_t* const* tabL = nullptr, const uint8_t* const* tabR = nullptr);
int launch_remap_nearest(const uint8_t* src, int sW, int sH, int ch, const CamModel& cm, uint8_t* dst, int W, int H,
                         cudaStream_t st);

// ---- prefilter.cu -----------------------------------------------------------------------------------
// dst rows are `dst_pitch` bytes apart (src is tightly packed)
int launch_prefilter_xsobel(const uint8_t* src, uint8_t* dst, size_t dst_pitch, int W, int H, int cap, cudaStream_t st);  // (unfused; tools only)
// scratch: W*H int32
int launch_prefilter_norm(const uint8_t* src, uint8_t* dst, size_t dst_pitch, int W, int H, int ps, int cap, int* scratch, cudaStream_t st);
int launch_bgr_to_gray(const uint8_t* src, uint8_t* dst, int n, int rgb_order, cudaStream_t st);
int launch_gray_to_bgr(const uint8_t* src, uint8_t* dst, int n, cudaStream_t st);
int launch_swap_rb(const uint8_t* src, uint8_t* dst, int n, cudaStream_t st);
// per-channel (min, max, sum) partials, 3 doubles per block and channel; elem_kind 0 = u8, 1 = s16, 2 = f32
int launch_mat_stats(const void* src, int elem_kind, size_t npix, int ch, double* partials, int nblocks, cudaStream_t st);

// ---- bm_sad.cu --------------------------------------------------------------------------------------
struct BMScratch {         // device scratch owned by the caller
    int* vol;              // generic-path cost volume
    size_t vol_bytes;
};
// Full matcher on prefiltered planes: fills disp (s16, every pixel) and, if cost != nullptr, the s16 cost plane.
// When cfg.disp12MaxDiff < 0 the valid-ROI mask is applied directly (pixels outside are FILTERED);
// otherwise all columns [lofs, lofs+width1) of the ROI rows are produced and the caller runs validate + mask.
// evals (optional) receives the number of (pixel, disparity) evaluations inside the valid ROI.
// Lp/Rp: prefiltered planes with row pitch `pitch` (multiple of 16) and at least PLANE_LEAD bytes of readable
// slack before row 0 and PLANE_TAIL bytes after the last row (the tile loaders read whole aligned words).
// nf > 1: a batch; prefiltered planes pre_stride bytes apart, disp / cost planes disp_stride bytes apart.  *evals is per frame.
int launch_block_match(const uint8_t* Lp, const uint8_t* Rp, size_t pitch, int W, int H, const BMConfig& cfg,
                       int16_t* disp, int16_t* cost, BMScratch* scratch, cudaStream_t st, double* evals,
                       int nf = 1, size_t pre_stride = 0, size_t disp_stride = 0, bool border_is_filled = false);
// border_is_filled: the caller guarantees that everything outside the rectangle the matcher writes already holds FILTERED
// (a slot's disparity planes keep their border from the previous frame with the same geometry), so the fill is skipped
constexpr size_t PLANE_LEAD = 256, PLANE_TAIL = 4096;
inline size_t plane_pitch(int W) { return ((size_t)W + 15) / 16 * 16; }
inline size_t plane_bytes(int W, int H) { return PLANE_LEAD + plane_pitch(W) * H + PLANE_TAIL; }
size_t bm_scratch_bytes(int W, int H, const BMConfig& cfg);

// ---- bm_cuda_compat.cu ------------------------------------------------------------------------------
// cv::cuda::StereoBM compatibility mode (SSD, CV_8UC1 integer disparity, 0 = invalid); tmpL/tmpR: W*H scratch (xsobel only)
int launch_cuda_compat_bm(const uint8_t* L, const uint8_t* R, uint8_t* tmpL, uint8_t* tmpR, int W, int H, int nd, int wsz,
                          bool xsobel, int cap, int tex_threshold, uint8_t* disp, cudaStream_t st);
int launch_u8_to_s16(const uint8_t* a, int16_t* b, size_t n, cudaStream_t st);
int launch_s16_to_u8(const int16_t* a, uint8_t* b, size_t n, cudaStream_t st);

// ---- post.cu ----------------------------------------------------------------------------------------
// cv::validateDisparity on the ROI rows followed by the valid-ROI column mask of those rows
int launch_validate_disp12(int16_t* disp, const int16_t* cost, int W, int H, const BMConfig& cfg, cudaStream_t st);
int launch_roi_mask(int16_t* disp, int W, int H, const BMConfig& cfg, cudaStream_t st);
// scratch: 3 * W*H int32 (parents, sizes, roots)
// batches: img planes img_stride bytes apart, scratch blocks scratch_stride bytes apart
int launch_filter_speckles(int16_t* img, int W, int H, int newVal, int maxSize, int maxDiff, int* scratch, cudaStream_t st,
                           int nf = 1, size_t img_stride = 0, size_t scratch_stride = 0);

// ---- reproject.cu -----------------------------------------------------------------------------------
// d16 -> f32 (d/16 - cxd); also reduces min(d16) of frame f into min_d16[f] (device ints, preset by this call).
// df_list (optional) overrides df with one destination per frame (pinned host memory written by the kernel).
int launch_disparity_to_float(const int16_t* d16, float* df, int n, double cxd, int* min_d16, cudaStream_t st,
                              int nf = 1, size_t d_stride = 0, size_t df_stride = 0, const PtrList* df_list = nullptr);
// cv::reprojectImageTo3D(handleMissingValues=true) fused with PointCloud2 packing.
// xyz (f32 x3, optional) and pc2 (32 B records, optional); color: ch = 1 (mono replicated) or 3 (BGR)
// qmask: bit (4*r + c) set when Q[r][c] != 0 (zero terms are skipped; the result is bit-identical)
// pc2_list (optional) overrides pc2 with one destination per frame ("pack straight into pinned host buffers").
// min_d16 == nullptr: the missing value is the constant dmin_const (a plane that comes out of the matcher always holds
// FILTERED = (minD - 1) * 16 as its minimum: the border columns), no reduction pass needed.
// df / df_list (optional): also write the float disparity plane (convertTo) from the same pass over d16.
struct ReprojectExtras {
    int dmin_const = 0;
    float* df = nullptr;
    size_t df_stride = 0;
    const PtrList* df_list = nullptr;
    const uint8_t* const* color_tab = nullptr;      // per-frame colour planes (see `tab*` above)
    // table over the disparity values the chain can produce (launch_reproject_lut): entry dv - dmin_const
    const void* lut = nullptr;
    int lut_n = 0;
};
// Everything of a PointCloud2 record that depends on the disparity alone -- the reciprocal of W, float(Z / W) with the
// missing-value and isValidPoint rules applied, the float disparity -- for dv = dmin .. dmin + n - 1 (16 bytes each).
// returns 1 when built (image_geometry's Q sparsity), 0 when this Q has no table (the kernel computes per pixel)
int launch_reproject_lut(void* lut, int n, int dmin, double cxd, const double* Q, unsigned qmask, cudaStream_t st);
constexpr int REPROJECT_LUT_MAX = 1 << 16;          // entries a slot keeps room for (numDisparities up to 4094)
int launch_reproject_pack(const int16_t* d16, int W, int H, double cxd, const double* Q, unsigned qmask, const int* min_d16,
                          const uint8_t* color, int ch, float* xyz, uint8_t* pc2, cudaStream_t st, int nf = 1,
                          size_t d_stride = 0, size_t color_stride = 0, size_t xyz_stride = 0, size_t pc2_stride = 0,
                          const PtrList* pc2_list = nullptr, const ReprojectExtras* extra = nullptr);
int launch_disparity_color(const int16_t* d16, uint8_t* bgra, int n, int nd, cudaStream_t st);

// ---- intpeak.cu -------------------------------------------------------------------------------------
int run_int_peak(int which, double* lane_ops_per_s, double* sm_mhz, cudaStream_t st);

}  // namespace b200s
