// Named-buffer API + extern "C" layer of libb200stereo.so (declared in include/b200_stereo.h).
// Mirrors the state and method set of gpuimageproc::GpuStereoProcessor
// (reference: include/gpuimageproc/GPUStereoProcessor.h:63-126, src/GPUStereoProcessor.cpp) with the
// OpenCV / image_geometry calls replaced by the sm_100a kernels in this directory.  The fused frame path lives in slots.cu.
#include "handle.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

using namespace b200s;

namespace b200s {

int fail(b200s_handle* h, int code, const std::string& msg)
{
    if (h) h->err = msg;
    return code;
}

int check_kernels(b200s_handle* h, const char* what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, B200S_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return B200S_OK;
}

void* mapped_alias(const void* p)
{
    if (!p) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

// Device -> host copy of a packed payload followed by a stream sync.  Pinned (or registered) destinations take one DMA.
// Pageable destinations -- a freshly allocated message buffer, the usual case with the reference-shaped calls -- go
// through two pinned staging chunks so that the DMA of chunk k+1 overlaps the host memcpy of chunk k, instead of the
// driver's slow pageable path.
constexpr size_t STAGE_CHUNK = 8u << 20;

int d2h_sync(b200s_handle* h, void* dst, const void* src, size_t bytes, cudaStream_t st)
{
    cudaPointerAttributes at;
    const bool pageable = cudaPointerGetAttributes(&at, dst) != cudaSuccess || at.type == cudaMemoryTypeUnregistered;
    cudaGetLastError();
    if (!pageable || bytes < (1u << 20)) {
        CUDA_OK(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
        CUDA_OK(h, cudaStreamSynchronize(st));
        return B200S_OK;
    }
    for (int i = 0; i < 2; ++i) {
        if (!h->stage_host[i]) CUDA_OK(h, cudaHostAlloc(&h->stage_host[i], STAGE_CHUNK, cudaHostAllocDefault));
        if (!h->stage_ev[i]) CUDA_OK(h, cudaEventCreateWithFlags(&h->stage_ev[i], cudaEventDisableTiming));
    }
    const size_t nchunks = (bytes + STAGE_CHUNK - 1) / STAGE_CHUNK;
    auto issue = [&](size_t k) -> cudaError_t {
        const size_t off = k * STAGE_CHUNK, len = std::min(STAGE_CHUNK, bytes - off);
        cudaError_t e = cudaMemcpyAsync(h->stage_host[k & 1], (const char*)src + off, len, cudaMemcpyDeviceToHost, st);
        return e != cudaSuccess ? e : cudaEventRecord(h->stage_ev[k & 1], st);
    };
    CUDA_OK(h, issue(0));
    for (size_t k = 0; k < nchunks; ++k) {
        if (k + 1 < nchunks) CUDA_OK(h, issue(k + 1));          // its staging chunk was drained in iteration k - 1
        CUDA_OK(h, cudaEventSynchronize(h->stage_ev[k & 1]));
        const size_t off = k * STAGE_CHUNK, len = std::min(STAGE_CHUNK, bytes - off);
        memcpy((char*)dst + off, h->stage_host[k & 1], len);
    }
    return B200S_OK;
}

// ---- parameter validation: cv::StereoBM::compute's rules -------------------------------
int validate_params(b200s_handle* h, const b200s_params& p)
{
    if (p.pre_filter_type != 0 && p.pre_filter_type != 1) return fail(h, B200S_EINVAL, "preFilterType must be 0 (NORMALIZED_RESPONSE) or 1 (XSOBEL)");
    if (p.pre_filter_size < 5 || p.pre_filter_size > 255 || p.pre_filter_size % 2 == 0) return fail(h, B200S_EINVAL, "preFilterSize must be odd and within 5..255");
    if (p.pre_filter_cap < 1 || p.pre_filter_cap > 63) return fail(h, B200S_EINVAL, "preFilterCap must be within 1..63");
    if (p.block_size < 5 || p.block_size > 255 || p.block_size % 2 == 0) return fail(h, B200S_EINVAL, "SADWindowSize must be odd, be within 5..255");
    if (p.num_disparities <= 0 || p.num_disparities % 16 != 0) return fail(h, B200S_EINVAL, "numDisparities must be positive and divisible by 16");
    if (p.texture_threshold < 0) return fail(h, B200S_EINVAL, "texture threshold must be non-negative");
    if (p.uniqueness_ratio < 0) return fail(h, B200S_EINVAL, "uniqueness ratio must be non-negative");
    return B200S_OK;
}

BMConfig bm_config(const b200s_params& p)
{
    BMConfig c;
    c.minD = p.min_disparity;
    c.nd = p.num_disparities;
    c.wsz = p.block_size;
    c.cap = p.pre_filter_cap;
    c.textureThreshold = p.texture_threshold;
    c.uniquenessRatio = p.uniqueness_ratio;
    c.disp12MaxDiff = p.disp12_max_diff;
    return c;
}

// ---- camera model ----------------------------------------------------------------------------------------
void inv3x3(const double* m, double* o)
{
    // closed-form 3x3 inverse, the same expression tree cv::invert uses for n = 3 (determinant, then cofactors * 1/det)
    double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], hh = m[7], i = m[8];
    double A = e * i - f * hh, B = -(d * i - f * g), C = d * hh - e * g;
    double det = a * A + b * B + c * C;
    double id = 1.0 / det;
    o[0] = A * id;  o[1] = -(b * i - c * hh) * id; o[2] = (b * f - c * e) * id;
    o[3] = B * id;  o[4] = (a * i - c * g) * id;   o[5] = -(a * f - c * d) * id;
    o[6] = C * id;  o[7] = -(a * hh - b * g) * id; o[8] = (a * e - b * d) * id;
}

void make_cam_model(const b200s_caminfo& ci, CamModel& cm)
{
    cm.fx = ci.K[0]; cm.fy = ci.K[4]; cm.cx = ci.K[2]; cm.cy = ci.K[5];
    double D[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < ci.n_D && i < 8; ++i) D[i] = ci.D[i];
    cm.k1 = D[0]; cm.k2 = D[1]; cm.p1 = D[2]; cm.p2 = D[3]; cm.k3 = D[4]; cm.k4 = D[5]; cm.k5 = D[6]; cm.k6 = D[7];
    double pr[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += ci.P[r * 4 + k] * ci.R[k * 3 + c];
            pr[r * 3 + c] = s;
        }
    inv3x3(pr, cm.ir);
}

// camera_calibration_parsers-style YAML (test/stereobm/test_data/left.yaml:1-20): flat keys, matrices as
// `name:` followed by rows/cols/data lines.  Hand-rolled: no yaml-cpp in this image.
bool parse_matrix(const std::string& txt, const char* name, double* out, int n_expected, int* n_found)
{
    size_t p = txt.find(std::string(name) + ":");
    if (p == std::string::npos) return false;
    size_t d = txt.find("data:", p);
    if (d == std::string::npos) return false;
    size_t lb = txt.find('[', d), rb = txt.find(']', d);
    if (lb == std::string::npos || rb == std::string::npos || rb < lb) return false;
    std::string body = txt.substr(lb + 1, rb - lb - 1);
    for (char& c : body) if (c == ',') c = ' ';
    std::istringstream is(body);
    int n = 0;
    double v;
    while (is >> v) {
        if (n < n_expected) out[n] = v;
        ++n;
    }
    if (n_found) *n_found = n;
    return n > 0 && (n_found || n == n_expected);
}

bool parse_int_key(const std::string& txt, const char* name, int* out)
{
    size_t p = txt.find(std::string(name) + ":");
    if (p == std::string::npos) return false;
    *out = atoi(txt.c_str() + p + strlen(name) + 1);
    return true;
}

int load_caminfo_yaml(b200s_handle* h, const char* path, b200s_caminfo* ci)
{
    std::ifstream f(path);
    if (!f) return fail(h, B200S_EIO, std::string("cannot open calibration file ") + path);
    std::stringstream ss;
    ss << f.rdbuf();
    std::string txt = ss.str();
    memset(ci, 0, sizeof(*ci));
    int nD = 0;
    if (!parse_int_key(txt, "image_width", &ci->width) || !parse_int_key(txt, "image_height", &ci->height) ||
        !parse_matrix(txt, "camera_matrix", ci->K, 9, nullptr) ||
        !parse_matrix(txt, "distortion_coefficients", ci->D, 8, &nD) ||
        !parse_matrix(txt, "rectification_matrix", ci->R, 9, nullptr) ||
        !parse_matrix(txt, "projection_matrix", ci->P, 12, nullptr))
        return fail(h, B200S_EIO, std::string("malformed calibration file ") + path);
    ci->n_D = nD > 8 ? 8 : nD;
    return B200S_OK;
}

}  // namespace b200s

namespace {

Mat* find_mat(b200s_handle* h, int id)
{
    auto it = h->mats.find(id);
    if (it == h->mats.end() || it->second.empty()) return nullptr;
    return &it->second;
}

int alloc_mat(b200s_handle* h, int id, int rows, int cols, int type, const char* enc, Mat** out)
{
    Mat& m = h->mats[id];
    m.rows = rows; m.cols = cols; m.type = type;
    if (enc) m.enc = enc;
    if (m.buf.ensure(m.bytes() + 64) != 0) return fail(h, B200S_ENOMEM, "cudaMalloc failed for a named buffer");
    *out = &m;
    return B200S_OK;
}

cudaStream_t stream_of(b200s_handle* h, int id) { return (id & B200S_SIDE_R) && !(id & B200S_SIDE_L) ? h->r_strm : h->l_strm; }

// The left stream reads right-side planes (matcher, point cloud); work that overwrites them on the right stream waits
// for the last such reader (the reverse dependency is ev_r; the reference has neither).
cudaError_t order_after_left_readers(b200s_handle* h, cudaStream_t st)
{
    return st == h->r_strm ? cudaStreamWaitEvent(st, h->ev_l, 0) : cudaSuccess;
}

// The reference hands its POINTS2 / DISPARITY_32F ids to projectDisparityTo3DPoints and enqueueSendPoints
// (test/UTest.cpp:378-382, src/StereoProcessor.cpp:281); here reprojection and packing read the fixed-point plane of
// the same side directly.
int fixed_point_plane_of(int id)
{
    return (id & (B200S_SRC_POINTS2 | B200S_SRC_DISPARITY_32F)) ? (B200S_SRC_DISPARITY | (id & B200S_SIDE_MASK)) : id;
}

}  // namespace

namespace b200s {

int ensure_map(b200s_handle* h, int side /*0 L, 1 R*/, cudaStream_t st)
{
    Camera& c = h->cam[side];
    if (c.map_valid) return B200S_OK;
    const size_t n = (size_t)c.info.width * c.info.height;
    if (h->flagdev.ensure(256)) return fail(h, B200S_ENOMEM, "cudaMalloc failed (flags)");
    // 4 B/px delta table first; a calibration whose shifts exceed +-1024 px falls back to the 8 B/px absolute table
    static const int force_abs = getenv("B200S_MAP_ABS32") ? atoi(getenv("B200S_MAP_ABS32")) : 0;
    for (int attempt = force_abs ? 1 : 0; attempt < 2; ++attempt) {
        const MapMode mode = attempt == 0 ? MAP_DELTA16 : MAP_ABS32;
        if (c.map.ensure(n * (mode == MAP_DELTA16 ? 4 : 8))) return fail(h, B200S_ENOMEM, "cudaMalloc failed for the rectification map");
        h->launches += launch_build_map(c.cm, c.info.width, c.info.height, c.map.p, mode, (int*)h->flagdev.p, st);
        int overflow = 0;
        if (mode == MAP_DELTA16) CUDA_OK(h, cudaMemcpyAsync(&overflow, h->flagdev.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        // the map may be consumed on another stream: make it visible everywhere once
        CUDA_OK(h, cudaStreamSynchronize(st));
        c.map_mode = mode;
        if (!overflow) break;
    }
    c.map_valid = true;
    return check_kernels(h, "build_map");
}

int map_for(b200s_handle* h, int side, cudaStream_t st, MapMode* mode, const void** map)
{
    if (h->rect_fly) {
        *mode = MAP_FLY;
        *map = nullptr;
        return B200S_OK;
    }
    int rc = ensure_map(h, side, st);
    if (rc) return rc;
    *mode = h->cam[side].map_mode;
    *map = h->cam[side].map.p;
    return B200S_OK;
}

// pitched prefiltered planes with readable slack on both sides (zeroed once so that stray reads are defined)
int ensure_pre_planes(b200s_handle* h, Work& w, int rows, int cols, int nf)
{
    const size_t need = plane_stride(cols, rows) * (size_t)nf;
    for (DevBuf* b : {&w.preL, &w.preR}) {
        if (b->cap < need) {
            if (b->ensure(need)) return fail(h, B200S_ENOMEM, "cudaMalloc failed (prefilter planes)");
            if (cudaMemset(b->p, 0, need) != cudaSuccess) return fail(h, B200S_ECUDA, "cudaMemset failed (prefilter planes)");
        }
    }
    return B200S_OK;
}

// prefilter + match + post-filters on rectified device planes; disp must hold rows*cols int16 per frame
int run_disparity(b200s_handle* h, Work& w, const uint8_t* L, const uint8_t* R, bool already_prefiltered, int rows,
                  int cols, int16_t* disp, cudaStream_t st, int nf, size_t src_stride, size_t disp_stride, const uint8_t* const* tabL,
                  const uint8_t* const* tabR, bool keep_border)
{
    const b200s_params& p = h->prm;
    int rc = validate_params(h, p);
    if (rc) return rc;
    if (p.block_size >= (rows < cols ? rows : cols))
        return fail(h, B200S_EINVAL, "SADWindowSize must be odd, be within 5..255 and be not larger than image width or height");
    size_t n = (size_t)rows * cols;
    BMConfig cfg = bm_config(p);
    const size_t pitch = plane_pitch(cols), pstride = plane_stride(cols, rows);
    if (nf <= 1 && disp_stride == 0) disp_stride = n * 2;
    const int cap_nf = nf > w.depth ? nf : w.depth;     // buffers are sized for the slot's whole batch once
    if (!already_prefiltered) {
        int rc2 = ensure_pre_planes(h, w, rows, cols, cap_nf);
        if (rc2) return rc2;
        uint8_t* pl = (uint8_t*)w.preL.p + PLANE_LEAD;
        uint8_t* pr = (uint8_t*)w.preR.p + PLANE_LEAD;
        if (p.pre_filter_type == 1) {
            // tiled x-Sobel of both sides (and all frames) in one launch: the fused kernel with the identity map
            h->launches += launch_rectify_xsobel_pair(L, R, cols, rows, MAP_NONE, nullptr, nullptr, h->cam[0].cm, h->cam[1].cm, nullptr,
                                                      nullptr, pl, pr, pitch, cols, rows, p.pre_filter_cap, st, nf, src_stride, 0, pstride, tabL, tabR);
        } else {
            const int one = launch_norm_prefilter_pair(L, R, cols, rows, MAP_NONE, nullptr, nullptr, h->cam[0].cm, h->cam[1].cm, nullptr, nullptr,
                                                       pl, pr, pitch, cols, rows, p.pre_filter_size, p.pre_filter_cap, st, nf, src_stride, 0, pstride, tabL, tabR);
            h->launches += one;
            if (!one) {     // preFilterSize > 21: two passes through a scratch plane, frame by frame
                if (w.normtmp.ensure(n * sizeof(int))) return fail(h, B200S_ENOMEM, "cudaMalloc failed (prefilter scratch)");
                for (int f = 0; f < nf; ++f) {
                    // sources named by the slot's address table: its host copy tells where frame f is
                    const uint8_t* lf = tabL ? (const uint8_t*)w.in_tab[f] : L + f * src_stride;
                    const uint8_t* rf = tabR ? (const uint8_t*)w.in_tab[MAX_BATCH + f] : R + f * src_stride;
                    h->launches += launch_prefilter_norm(lf, pl + f * pstride, pitch, cols, rows, p.pre_filter_size, p.pre_filter_cap, (int*)w.normtmp.p, st);
                    h->launches += launch_prefilter_norm(rf, pr + f * pstride, pitch, cols, rows, p.pre_filter_size, p.pre_filter_cap, (int*)w.normtmp.p, st);
                }
            }
        }
    }
    // the matcher always reads the pitched, slack-padded prefiltered planes of this Work
    const uint8_t* Lp = (const uint8_t*)w.preL.p + PLANE_LEAD;
    const uint8_t* Rp = (const uint8_t*)w.preR.p + PLANE_LEAD;
    int16_t* cost = nullptr;
    if (cfg.disp12MaxDiff >= 0) {
        if (w.cost.ensure(disp_stride * cap_nf)) return fail(h, B200S_ENOMEM, "cudaMalloc failed (cost plane)");
        cost = (int16_t*)w.cost.p;
    }
    if (w.vol.ensure(bm_scratch_bytes(cols, rows, cfg))) return fail(h, B200S_ENOMEM, "cudaMalloc failed (matcher scratch)");
    BMScratch sc{(int*)w.vol.p, w.vol.cap};
    if (h->timing) {
        if (!w.ev_bm0) { cudaEventCreate(&w.ev_bm0); cudaEventCreate(&w.ev_bm1); }
        cudaEventRecord(w.ev_bm0, st);
    }
    // A slot's disparity planes keep the FILTERED border of the previous batch: the matcher and the post-filters only ever
    // write FILTERED there, so the fill is needed again only when the geometry (or FILTERED itself) changes.
    bool border_ok = false;
    if (keep_border) {
        std::string key;
        const int kv[8] = {rows, cols, p.min_disparity, p.num_disparities, p.block_size, p.disp12_max_diff >= 0, nf > w.depth ? nf : w.depth, 0};
        key.assign((const char*)kv, sizeof kv);
        key.append((const char*)&disp, sizeof disp);
        border_ok = key == w.border_key;
        w.border_key = key;
        if (nf < w.depth && !border_ok) w.border_key.clear();      // a partial batch fills only its own frames
    }
    int l = launch_block_match(Lp, Rp, pitch, cols, rows, cfg, disp, cost, &sc, st, &w.last_evals, nf, pstride, disp_stride, border_ok);
    w.last_evals *= nf;
    if (h->timing) { cudaEventRecord(w.ev_bm1, st); w.timed = true; }
    if (l < 0) return fail(h, B200S_ECUDA, "block matcher launch failed (code " + std::to_string(l) + "): " + cudaGetErrorString(cudaGetLastError()));
    h->launches += l;
    if (cfg.disp12MaxDiff >= 0) {
        for (int f = 0; f < nf; ++f) {
            int v = launch_validate_disp12((int16_t*)((uint8_t*)disp + f * disp_stride), (const int16_t*)((const uint8_t*)cost + f * disp_stride),
                                           cols, rows, cfg, st);
            if (v < 0) return fail(h, B200S_EUNSUPPORTED, "disp12MaxDiff needs image width <= 65535");
            h->launches += v;     // the valid-ROI mask of the rows is applied by the same kernel
        }
    }
    if (p.speckle_window_size > 0 && p.speckle_range >= 0) {
        const size_t cstride = align256(3 * n * sizeof(int) + 64);
        if (w.ccl.ensure(cstride * cap_nf)) return fail(h, B200S_ENOMEM, "cudaMalloc failed (speckle scratch)");
        h->launches += launch_filter_speckles(disp, cols, rows, (p.min_disparity - 1) * 16, p.speckle_window_size,
                                              p.speckle_range, (int*)w.ccl.p, st, nf, disp_stride, cstride);
    }
    ++h->stats_frames;
    return check_kernels(h, "disparity chain");
}

int ensure_misc(b200s_handle* h, Work& w)
{
    if (w.misc.ensure(MISC_BYTES)) return fail(h, B200S_ENOMEM, "cudaMalloc failed (misc)");
    return B200S_OK;
}

void fill_disparity_meta(const b200s_handle* h, int rows, int cols, b200s_disparity_meta* m)
{
    const b200s_params& p = h->prm;
    m->width = cols; m->height = rows; m->step = cols * 4;
    m->f = (float)h->fx_right;
    m->T = (float)h->baseline;
    m->min_disparity = (float)p.min_disparity;
    m->max_disparity = (float)(p.min_disparity + p.num_disparities - 1);
    m->delta_d = 1.0f / 16.0f;
    int border = p.block_size / 2;
    int left = p.num_disparities + p.min_disparity + border - 1;
    int wtf = p.min_disparity >= 0 ? border + p.min_disparity : (border > -p.min_disparity ? border : -p.min_disparity);
    int right = cols - 1 - wtf, top = border, bottom = rows - 1 - border;
    m->valid_x_offset = left; m->valid_y_offset = top; m->valid_width = right - left; m->valid_height = bottom - top;
}

void fill_pc2_meta(int rows, int cols, b200s_pc2_meta* m)
{
    m->width = cols; m->height = rows; m->point_step = 32; m->row_step = 32 * cols;
    m->is_bigendian = 0; m->is_dense = 0;
    m->off_x = 0; m->off_y = 4; m->off_z = 8; m->off_rgb = 16;
}

}  // namespace b200s

// ==========================================================================================================
extern "C" {

const char* b200s_version(void) { return "b200-stereo 0.2 (sm_100a)"; }

int b200s_default_params(b200s_params* p)
{
    if (!p) return B200S_EINVAL;
    p->pre_filter_type = 1; p->pre_filter_size = 9; p->pre_filter_cap = 31;
    p->block_size = 21; p->min_disparity = 0; p->num_disparities = 64;
    p->texture_threshold = 10; p->uniqueness_ratio = 15;
    p->speckle_window_size = 0; p->speckle_range = 0; p->disp12_max_diff = -1; p->refine_disparity = 0;
    return B200S_OK;
}

int b200s_create(int device, b200s_handle** out)
{
    if (!out) return B200S_EINVAL;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return B200S_ECUDA;
    DeviceGuard g(device);      // the caller's current device is restored when b200s_create returns
    {
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess || cur != device) return B200S_ECUDA;
    }
    b200s_handle* h = new (std::nothrow) b200s_handle();
    if (!h) return B200S_ENOMEM;
    h->device = device;
    // state of the reference's CPU matcher after its constructor (src/GPUStereoProcessor.cpp:18-38):
    // createStereoBM(48, 19) mirrored into cv::StereoBM, preFilterSize 5, PREFILTER_XSOBEL (:27-30), texture threshold /
    // uniqueness / disp12MaxDiff copied from the cuda matcher's getters (3 / 0 / 0)
    if (const char* e = getenv("B200S_GRAPH")) h->use_graphs = atoi(e) != 0;
    if (const char* e = getenv("B200S_PACK_DIRECT")) h->pack_direct = atoi(e);
    b200s_default_params(&h->prm);
    h->prm.pre_filter_type = 1; h->prm.pre_filter_size = 5; h->prm.num_disparities = 48; h->prm.block_size = 19;
    h->prm.texture_threshold = 3; h->prm.uniqueness_ratio = 0; h->prm.disp12_max_diff = 0;
    if (cudaStreamCreateWithFlags(&h->l_strm, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->r_strm, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_r, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_l, cudaEventDisableTiming) != cudaSuccess) {
        delete h;
        return B200S_ECUDA;
    }
    h->w0.st = h->l_strm;
    *out = h;
    return B200S_OK;
}

int b200s_destroy(b200s_handle* h)
{
    if (!h) return B200S_OK;
    DeviceGuard g(h->device);
    cudaDeviceSynchronize();
    for (auto& kv : h->mats) kv.second.buf.release();
    for (Work& w : h->slots) w.release();
    h->w0.release();
    for (int s = 0; s < 2; ++s) h->cam[s].map.release();
    h->Qdev.release();
    h->flagdev.release();
    if (h->ev_r) cudaEventDestroy(h->ev_r);
    if (h->ev_l) cudaEventDestroy(h->ev_l);
    for (int i = 0; i < 2; ++i) {
        if (h->stage_host[i]) cudaFreeHost(h->stage_host[i]);
        if (h->stage_ev[i]) cudaEventDestroy(h->stage_ev[i]);
    }
    if (h->batch_start) cudaEventDestroy(h->batch_start);
    for (cudaEvent_t e : h->batch_end) cudaEventDestroy(e);
    if (h->l_strm) cudaStreamDestroy(h->l_strm);
    if (h->r_strm) cudaStreamDestroy(h->r_strm);
    delete h;
    return B200S_OK;
}

const char* b200s_last_error_string(const b200s_handle* h) { return h ? h->err.c_str() : "null handle"; }

int b200s_set_calibration(b200s_handle* h, const b200s_caminfo* left, const b200s_caminfo* right)
{
    if (!h || !left || !right) return B200S_EINVAL;
    DeviceGuard g(h->device);
    if (left->width <= 0 || left->height <= 0 || left->width != right->width || left->height != right->height)
        return fail(h, B200S_EINVAL, "left/right camera_info sizes must be positive and equal");
    h->cam[0].info = *left;
    h->cam[1].info = *right;
    for (int s = 0; s < 2; ++s) {
        make_cam_model(h->cam[s].info, h->cam[s].cm);
        h->cam[s].map_valid = false;
    }
    // image_geometry::StereoCameraModel::updateQ
    const double* Pl = left->P;
    const double* Pr = right->P;
    double fx = Pl[0], fy = Pl[5], cx = Pl[2], cy = Pl[6], cxr = Pr[2];
    double Tx = Pr[3] / Pr[0];
    memset(h->Q, 0, sizeof(h->Q));
    h->Q[0] = fy * Tx;       h->Q[3] = -fy * cx * Tx;
    h->Q[5] = fx * Tx;       h->Q[7] = -fx * cy * Tx;
    h->Q[11] = fx * fy * Tx;
    h->Q[14] = -fy;          h->Q[15] = fy * (cx - cxr);
    h->qmask = 0;
    for (int i = 0; i < 16; ++i)
        if (h->Q[i] != 0.0) h->qmask |= 1u << i;
    // a NaN/Inf coordinate cannot occur (x, y are pixel indices, d is a finite float), so skipping zero terms is exact
    h->baseline = -Pr[3] / Pr[0];
    h->fx_right = Pr[0];
    h->cxd = cx - cxr;
    if (h->Qdev.ensure(sizeof(h->Q))) return fail(h, B200S_ENOMEM, "cudaMalloc failed (Q)");
    // frames in flight on the (non-blocking) slot streams still read the old Q and maps: drain the device first
    CUDA_OK(h, cudaDeviceSynchronize());
    for (Work& w : h->slots) w.drop_graphs();
    CUDA_OK(h, cudaMemcpy(h->Qdev.p, h->Q, sizeof(h->Q), cudaMemcpyHostToDevice));
    h->model_ok = true;
    ++h->model_version;
    return B200S_OK;
}

int b200s_load_calibration_files(b200s_handle* h, const char* left_yaml, const char* right_yaml)
{
    if (!h || !left_yaml || !right_yaml) return B200S_EINVAL;
    b200s_caminfo l, r;
    int rc = load_caminfo_yaml(h, left_yaml, &l);
    if (rc) return rc;
    rc = load_caminfo_yaml(h, right_yaml, &r);
    if (rc) return rc;
    return b200s_set_calibration(h, &l, &r);
}

int b200s_is_model_initialised(const b200s_handle* h) { return h && h->model_ok ? 1 : 0; }

int b200s_get_model(const b200s_handle* h, double* Q16, double* baseline, double* fx, double* cxd)
{
    if (!h) return B200S_EINVAL;
    if (!h->model_ok) return B200S_ENOTINIT;
    if (Q16) memcpy(Q16, h->Q, sizeof(h->Q));
    if (baseline) *baseline = h->baseline;
    if (fx) *fx = h->fx_right;
    if (cxd) *cxd = h->cxd;
    return B200S_OK;
}

int b200s_set_params(b200s_handle* h, const b200s_params* p)
{
    if (!h || !p) return B200S_EINVAL;
    int rc = validate_params(h, *p);
    if (rc) return rc;
    h->prm = *p;
    return B200S_OK;
}

int b200s_get_params(const b200s_handle* h, b200s_params* p)
{
    if (!h || !p) return B200S_EINVAL;
    *p = h->prm;
    return B200S_OK;
}

int b200s_set_rectify_mode(b200s_handle* h, int on_the_fly)
{
    if (!h) return B200S_EINVAL;
    h->rect_fly = on_the_fly != 0;
    ++h->model_version;
    return B200S_OK;
}

int b200s_enable_timing(b200s_handle* h, int on)
{
    if (!h) return B200S_EINVAL;
    h->timing = on != 0;
    return B200S_OK;
}

// ---- named buffers ---------------------------------------------------------------------------------------
int b200s_upload(b200s_handle* h, int mat_id, const void* data, int rows, int cols, int type, size_t step, const char* encoding)
{
    if (!h || !data || rows <= 0 || cols <= 0) return B200S_EINVAL;
    DeviceGuard g(h->device);
    int es = elem_size(type);
    if (!es) return fail(h, B200S_EUNSUPPORTED, "unsupported element type");
    if (step == 0) step = (size_t)cols * es;
    if (step < (size_t)cols * es) return fail(h, B200S_EINVAL, "step smaller than a row");
    Mat* m;
    int rc = alloc_mat(h, mat_id, rows, cols, type, encoding ? encoding : "", &m);
    if (rc) return rc;
    cudaStream_t st = stream_of(h, mat_id);
    CUDA_OK(h, order_after_left_readers(h, st));
    CUDA_OK(h, cudaMemcpy2DAsync(m->buf.p, (size_t)cols * es, data, step, (size_t)cols * es, rows, cudaMemcpyHostToDevice, st));
    return B200S_OK;
}

int b200s_download(b200s_handle* h, int mat_id, void* dst, size_t dst_step)
{
    if (!h || !dst) return B200S_EINVAL;
    DeviceGuard g(h->device);
    Mat* m = find_mat(h, mat_id);
    if (!m) return fail(h, B200S_ENOBUF, "named buffer is empty");
    int es = elem_size(m->type);
    if (dst_step == 0) dst_step = (size_t)m->cols * es;
    cudaStream_t st = stream_of(h, mat_id);
    if (dst_step == (size_t)m->cols * es) return d2h_sync(h, dst, m->buf.p, m->bytes(), st);
    CUDA_OK(h, cudaMemcpy2DAsync(dst, dst_step, m->buf.p, (size_t)m->cols * es, (size_t)m->cols * es, m->rows, cudaMemcpyDeviceToHost, st));
    CUDA_OK(h, cudaStreamSynchronize(st));
    return B200S_OK;
}

int b200s_mat_info(const b200s_handle* h, int mat_id, int* rows, int* cols, int* type)
{
    if (!h) return B200S_EINVAL;
    auto it = h->mats.find(mat_id);
    if (it == h->mats.end() || it->second.empty()) return B200S_ENOBUF;
    if (rows) *rows = it->second.rows;
    if (cols) *cols = it->second.cols;
    if (type) *type = it->second.type;
    return B200S_OK;
}

int b200s_device_ptr(b200s_handle* h, int mat_id, void** dptr, size_t* bytes)
{
    if (!h || !dptr) return B200S_EINVAL;
    Mat* m = find_mat(h, mat_id);
    if (!m) return fail(h, B200S_ENOBUF, "named buffer is empty");
    *dptr = m->buf.p;
    if (bytes) *bytes = m->bytes();
    return B200S_OK;
}

// ---- colour conversion (convertColor, src/GPUStereoProcessor.cpp:119-172; mono8 / bgr8 / rgb8 only) -------
static int convert_raw(b200s_handle* h, int side, bool to_color)
{
    if (!h || (side != B200S_SIDE_L && side != B200S_SIDE_R)) return B200S_EINVAL;
    DeviceGuard g(h->device);
    Mat* src = find_mat(h, B200S_SRC_RAW | side);
    if (!src) return fail(h, B200S_ENOBUF, "raw buffer is empty");
    const std::string& e = src->enc;
    bool mono = (e == "mono8" || (e.empty() && src->type == B200S_8UC1));
    bool bgr = (e == "bgr8" || (e.empty() && src->type == B200S_8UC3)), rgb = (e == "rgb8");
    if (!((mono && src->type == B200S_8UC1) || ((bgr || rgb) && src->type == B200S_8UC3)))
        return fail(h, B200S_EUNSUPPORTED, "raw encoding '" + e + "' is outside the hot path (mono8, bgr8, rgb8 are supported)");
    cudaStream_t st = stream_of(h, side);
    CUDA_OK(h, order_after_left_readers(h, st));
    int n = src->rows * src->cols;
    Mat* dst;
    int rc = alloc_mat(h, (to_color ? B200S_SRC_COLOR : B200S_SRC_MONO) | side, src->rows, src->cols,
                       to_color ? B200S_8UC3 : B200S_8UC1, to_color ? "bgr8" : "mono8", &dst);
    if (rc) return rc;
    src = find_mat(h, B200S_SRC_RAW | side);   // the map may have rehashed
    if (to_color) {
        if (mono) h->launches += launch_gray_to_bgr((const uint8_t*)src->buf.p, (uint8_t*)dst->buf.p, n, st);
        else if (rgb) h->launches += launch_swap_rb((const uint8_t*)src->buf.p, (uint8_t*)dst->buf.p, n, st);
        else CUDA_OK(h, cudaMemcpyAsync(dst->buf.p, src->buf.p, src->bytes(), cudaMemcpyDeviceToDevice, st));
    } else {
        if (mono) CUDA_OK(h, cudaMemcpyAsync(dst->buf.p, src->buf.p, src->bytes(), cudaMemcpyDeviceToDevice, st));
        else h->launches += launch_bgr_to_gray((const uint8_t*)src->buf.p, (uint8_t*)dst->buf.p, n, rgb ? 1 : 0, st);
    }
    return check_kernels(h, "convert_raw");
}

int b200s_convert_raw_to_mono(b200s_handle* h, int side) { return convert_raw(h, side, false); }
int b200s_convert_raw_to_color(b200s_handle* h, int side) { return convert_raw(h, side, true); }

// ---- rectify ---------------------------------------------------------------------------------------------
int b200s_rectify(b200s_handle* h, int src_id, int dst_id, int interpolation)
{
    if (!h) return B200S_EINVAL;
    DeviceGuard g(h->device);
    if (!h->model_ok) return fail(h, B200S_ENOTINIT, "stereo model not initialised");
    if (interpolation != B200S_INTER_LINEAR && interpolation != B200S_INTER_NEAREST)
        return fail(h, B200S_EUNSUPPORTED, "only INTER_LINEAR and INTER_NEAREST rectification are implemented");
    Mat* src = find_mat(h, src_id);
    if (!src) return fail(h, B200S_ENOBUF, "source buffer is empty");
    int ch = src->type == B200S_8UC1 ? 1 : (src->type == B200S_8UC3 ? 3 : (src->type == B200S_8UC4 ? 4 : 0));
    if (!ch) return fail(h, B200S_EUNSUPPORTED, "rectify expects 8-bit 1/3/4-channel data");
    int side = (src_id & B200S_SIDE_R) && !(src_id & B200S_SIDE_L) ? 1 : 0;
    const Camera& cam = h->cam[side];
    int W = cam.info.width, H = cam.info.height;
    cudaStream_t st = stream_of(h, src_id);
    int sW = src->cols, sH = src->rows, type = src->type;
    std::string enc = src->enc;
    Mat* dst;
    int rc = alloc_mat(h, dst_id, H, W, type, enc.c_str(), &dst);
    if (rc) return rc;
    src = find_mat(h, src_id);
    CUDA_OK(h, order_after_left_readers(h, st));
    if (interpolation == B200S_INTER_LINEAR) {
        MapMode mode = MAP_FLY;
        const void* map = nullptr;
        rc = map_for(h, side, st, &mode, &map);
        if (rc) return rc;
        h->launches += launch_remap((const uint8_t*)src->buf.p, sW, sH, ch, map, mode, cam.cm, (uint8_t*)dst->buf.p, W, H, st);
    } else {
        h->launches += launch_remap_nearest((const uint8_t*)src->buf.p, sW, sH, ch, cam.cm, (uint8_t*)dst->buf.p, W, H, st);
    }
    return check_kernels(h, "rectify");
}

// ---- disparity -------------------------------------------------------------------------------------------
int b200s_compute_disparity(b200s_handle* h, int left_id, int right_id, int disp_id)
{
    if (!h) return B200S_EINVAL;
    DeviceGuard g(h->device);
    Mat* L = find_mat(h, left_id);
    Mat* R = find_mat(h, right_id);
    if (!L || !R) return fail(h, B200S_ENOBUF, "left/right buffer is empty");
    if (L->type != B200S_8UC1 || R->type != B200S_8UC1 || L->rows != R->rows || L->cols != R->cols)
        return fail(h, B200S_EINVAL, "both input images must have CV_8UC1 format and equal size");
    int rows = L->rows, cols = L->cols;
    // the right image may still be in flight on r_strm (reference bug B6: no such dependency there)
    cudaStream_t st = h->l_strm;
    CUDA_OK(h, cudaEventRecord(h->ev_r, h->r_strm));
    CUDA_OK(h, cudaStreamWaitEvent(st, h->ev_r, 0));
    Mat* D;
    int rc = alloc_mat(h, disp_id, rows, cols, B200S_16SC1, "", &D);
    if (rc) return rc;
    L = find_mat(h, left_id);
    R = find_mat(h, right_id);
    rc = run_disparity(h, h->w0, (const uint8_t*)L->buf.p, (const uint8_t*)R->buf.p, false, rows, cols, (int16_t*)D->buf.p, st);
    if (rc) return rc;
    // float plane of the same side (what computeDisparity(cv::Mat...) returns, src/GPUStereoProcessor.cpp:320)
    int side = disp_id & B200S_SIDE_MASK;
    Mat* F;
    rc = alloc_mat(h, B200S_SRC_DISPARITY_32F | side, rows, cols, B200S_32FC1, "32FC1", &F);
    if (rc) return rc;
    D = find_mat(h, disp_id);
    rc = ensure_misc(h, h->w0);
    if (rc) return rc;
    h->launches += launch_disparity_to_float((const int16_t*)D->buf.p, (float*)F->buf.p, rows * cols,
                                             h->model_ok ? h->cxd : 0.0, (int*)h->w0.misc.p, st);
    CUDA_OK(h, cudaEventRecord(h->ev_l, st));    // the right-side stream must not overwrite its planes before this point
    return check_kernels(h, "compute_disparity");
}

// cv::cuda::StereoBM compatibility mode: what block_matcher_gpu_->compute produces at src/GPUStereoProcessor.cpp:283
int b200s_compute_disparity_cuda_compat(b200s_handle* h, int left_id, int right_id, int disp_id)
{
    if (!h) return B200S_EINVAL;
    DeviceGuard g(h->device);
    Mat* L = find_mat(h, left_id);
    Mat* R = find_mat(h, right_id);
    if (!L || !R) return fail(h, B200S_ENOBUF, "left/right buffer is empty");
    if (L->type != B200S_8UC1 || R->type != B200S_8UC1 || L->rows != R->rows || L->cols != R->cols)
        return fail(h, B200S_EINVAL, "both input images must have CV_8UC1 format and equal size");
    const b200s_params& p = h->prm;
    // cv::cuda::StereoBM limits (opencv_contrib cudastereo): ndisp a multiple of 8 up to 256, window radius 1..25
    if (p.num_disparities <= 0 || p.num_disparities > 256 || (p.num_disparities & 7))
        return fail(h, B200S_EINVAL, "cuda-compat: numDisparities must be a positive multiple of 8, at most 256");
    if ((p.block_size >> 1) < 1 || (p.block_size >> 1) > 25) return fail(h, B200S_EINVAL, "cuda-compat: unsupported window size (3..51)");
    if (p.pre_filter_cap < 1 || p.pre_filter_cap > 63 || p.texture_threshold < 0) return fail(h, B200S_EINVAL, "cuda-compat: bad prefilter cap / texture threshold");
    const int rows = L->rows, cols = L->cols;
    cudaStream_t st = h->l_strm;
    CUDA_OK(h, cudaEventRecord(h->ev_r, h->r_strm));
    CUDA_OK(h, cudaStreamWaitEvent(st, h->ev_r, 0));
    Mat* D;
    int rc = alloc_mat(h, disp_id, rows, cols, B200S_8UC1, "mono8", &D);
    if (rc) return rc;
    L = find_mat(h, left_id);
    R = find_mat(h, right_id);
    const size_t n = (size_t)rows * cols;
    const bool xs = p.pre_filter_type == 1;
    if (xs && (h->w0.preL.ensure(plane_bytes(cols, rows)) || h->w0.preR.ensure(plane_bytes(cols, rows))))
        return fail(h, B200S_ENOMEM, "cudaMalloc failed (prefilter planes)");
    h->launches += launch_cuda_compat_bm((const uint8_t*)L->buf.p, (const uint8_t*)R->buf.p, (uint8_t*)h->w0.preL.p, (uint8_t*)h->w0.preR.p,
                                         cols, rows, p.num_disparities, p.block_size, xs, p.pre_filter_cap, p.texture_threshold,
                                         (uint8_t*)D->buf.p, st);
    (void)n;
    CUDA_OK(h, cudaEventRecord(h->ev_l, st));
    return check_kernels(h, "compute_disparity_cuda_compat");
}

int b200s_filter_speckles(b200s_handle* h, int disp_id)
{
    if (!h) return B200S_EINVAL;
    DeviceGuard g(h->device);
    Mat* D = find_mat(h, disp_id);
    if (!D || (D->type != B200S_16SC1 && D->type != B200S_8UC1)) return fail(h, B200S_ENOBUF, "disparity buffer is empty or neither CV_16SC1 nor CV_8UC1");
    const b200s_params& p = h->prm;
    if (p.speckle_window_size <= 0 || p.speckle_range < 0) return B200S_OK;
    size_t n = (size_t)D->rows * D->cols;
    if (h->w0.ccl.ensure(3 * n * sizeof(int))) return fail(h, B200S_ENOMEM, "cudaMalloc failed (speckle scratch)");
    cudaStream_t st = stream_of(h, disp_id);
    if (D->type == B200S_8UC1) {
        // the reference's own flow on the cuda matcher's u8 plane (src/GPUStereoProcessor.cpp:367-385):
        // convertTo(CV_16S), cv::filterSpeckles(newVal 0, maxSpeckleSize, maxSpeckleDiff in integer disparities), convertTo(CV_8U)
        if (h->w0.disp.ensure(n * sizeof(int16_t))) return fail(h, B200S_ENOMEM, "cudaMalloc failed (speckle plane)");
        h->launches += launch_u8_to_s16((const uint8_t*)D->buf.p, (int16_t*)h->w0.disp.p, n, st);
        h->launches += launch_filter_speckles((int16_t*)h->w0.disp.p, D->cols, D->rows, 0, p.speckle_window_size,
                                              (p.speckle_range + 8) / 16, (int*)h->w0.ccl.p, st);
        h->launches += launch_s16_to_u8((const int16_t*)h->w0.disp.p, (uint8_t*)D->buf.p, n, st);
        return check_kernels(h, "filter_speckles (u8)");
    }
    h->launches += launch_filter_speckles((int16_t*)D->buf.p, D->cols, D->rows, (p.min_disparity - 1) * 16,
                                          p.speckle_window_size, p.speckle_range, (int*)h->w0.ccl.p, st);
    return check_kernels(h, "filter_speckles");
}

int b200s_filter_speckles_host(b200s_handle* h, int16_t* img, int rows, int cols, size_t step, int new_val, int max_size, int max_diff)
{
    if (!h || !img || rows <= 0 || cols <= 0) return B200S_EINVAL;
    DeviceGuard g(h->device);
    if (step == 0) step = (size_t)cols * 2;
    size_t n = (size_t)rows * cols;
    DevBuf tmp;
    if (tmp.ensure(n * 2) || h->w0.ccl.ensure(3 * n * sizeof(int))) { tmp.release(); return fail(h, B200S_ENOMEM, "cudaMalloc failed"); }
    cudaStream_t st = h->l_strm;
    cudaError_t e = cudaMemcpy2DAsync(tmp.p, (size_t)cols * 2, img, step, (size_t)cols * 2, rows, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        h->launches += launch_filter_speckles((int16_t*)tmp.p, cols, rows, new_val, max_size, max_diff, (int*)h->w0.ccl.p, st);
        e = cudaMemcpy2DAsync(img, step, tmp.p, (size_t)cols * 2, (size_t)cols * 2, rows, cudaMemcpyDeviceToHost, st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    tmp.release();
    if (e != cudaSuccess) return fail(h, B200S_ECUDA, cudaGetErrorString(e));
    return check_kernels(h, "filter_speckles_host");
}

int b200s_compute_disparity_image(b200s_handle* h, int disp_id, int img_id)
{
    if (!h) return B200S_EINVAL;
    DeviceGuard g(h->device);
    Mat* D = find_mat(h, disp_id);
    if (!D || D->type != B200S_16SC1) return fail(h, B200S_ENOBUF, "disparity buffer is empty or not CV_16SC1");
    int rows = D->rows, cols = D->cols;
    Mat* I;
    int rc = alloc_mat(h, img_id, rows, cols, B200S_8UC4, "bgra8", &I);
    if (rc) return rc;
    D = find_mat(h, disp_id);
    h->launches += launch_disparity_color((const int16_t*)D->buf.p, (uint8_t*)I->buf.p, rows * cols, h->prm.num_disparities, stream_of(h, disp_id));
    return check_kernels(h, "disparity_image");
}

int b200s_project_to_3d(b200s_handle* h, int disp_id, int points_id)
{
    if (!h) return B200S_EINVAL;
    DeviceGuard g(h->device);
    if (!h->model_ok) return fail(h, B200S_ENOTINIT, "stereo model not initialised");
    disp_id = fixed_point_plane_of(disp_id);
    Mat* D = find_mat(h, disp_id);
    if (!D || D->type != B200S_16SC1) return fail(h, B200S_ENOBUF, "disparity buffer is empty or not CV_16SC1");
    int rows = D->rows, cols = D->cols;
    Mat* P;
    int rc = alloc_mat(h, points_id, rows, cols, B200S_32FC3, "32FC3", &P);
    if (rc) return rc;
    D = find_mat(h, disp_id);
    rc = ensure_misc(h, h->w0);
    if (rc) return rc;
    cudaStream_t st = stream_of(h, disp_id);
    h->launches += launch_disparity_to_float((const int16_t*)D->buf.p, nullptr, rows * cols, h->cxd, (int*)h->w0.misc.p, st);
    h->launches += launch_reproject_pack((const int16_t*)D->buf.p, cols, rows, h->cxd, (const double*)h->Qdev.p, h->qmask,
                                         (const int*)h->w0.misc.p, nullptr, 1, (float*)P->buf.p, nullptr, st);
    return check_kernels(h, "project_to_3d");
}

int b200s_wait(b200s_handle* h, int side)
{
    if (!h) return B200S_EINVAL;
    DeviceGuard g(h->device);
    if (side == 0 || (side & B200S_SIDE_L)) CUDA_OK(h, cudaStreamSynchronize(h->l_strm));
    if (side == 0 || (side & B200S_SIDE_R)) CUDA_OK(h, cudaStreamSynchronize(h->r_strm));
    return B200S_OK;
}

// ---- packing ---------------------------------------------------------------------------------------------
// The reference's senders publish from a stream callback once the device->host copy has finished
// (src/GpuSenderIfc.cpp:13-26).  The *_async entry points do the same: kernels (+ copy) are enqueued on the side's
// stream, then `done(user, status)` is called by the CUDA runtime's callback thread when the payload is complete in
// dst (no CUDA calls inside `done`).  done == NULL makes the call synchronous (returns with the payload in dst).
namespace {

struct DoneCtx {
    b200s_done_fn fn;
    void* user;
};

void CUDART_CB done_trampoline(void* p)
{
    DoneCtx* c = (DoneCtx*)p;
    c->fn(c->user, B200S_OK);
    delete c;
}

// src == nullptr: a kernel has already written the payload straight into (pinned) dst
int finish_pack(b200s_handle* h, cudaStream_t st, void* dst, const void* src, size_t bytes, b200s_done_fn done, void* user)
{
    if (!done) {
        if (src) return d2h_sync(h, dst, src, bytes, st);
        CUDA_OK(h, cudaStreamSynchronize(st));
        return B200S_OK;
    }
    if (src) CUDA_OK(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    DoneCtx* c = new DoneCtx{done, user};
    cudaError_t e = cudaLaunchHostFunc(st, done_trampoline, c);
    if (e != cudaSuccess) {
        delete c;
        return fail(h, B200S_ECUDA, std::string("cudaLaunchHostFunc: ") + cudaGetErrorString(e));
    }
    return B200S_OK;
}

}  // namespace

int b200s_pack_image_async(b200s_handle* h, int mat_id, void* dst, size_t cap_bytes, int* rows, int* cols, int* step,
                           b200s_done_fn done, void* user)
{
    if (!h || !dst) return B200S_EINVAL;
    DeviceGuard g(h->device);
    Mat* m = find_mat(h, mat_id);
    if (!m) return fail(h, B200S_ENOBUF, "named buffer is empty");
    if (cap_bytes < m->bytes()) return fail(h, B200S_EINVAL, "destination too small");
    if (rows) *rows = m->rows;
    if (cols) *cols = m->cols;
    if (step) *step = m->cols * elem_size(m->type);   // GpuSenderImage.cpp:20: width * bitdepth * channels / 8
    return finish_pack(h, stream_of(h, mat_id), dst, m->buf.p, m->bytes(), done, user);
}

int b200s_pack_image(b200s_handle* h, int mat_id, void* dst, size_t cap_bytes, int* rows, int* cols, int* step)
{
    return b200s_pack_image_async(h, mat_id, dst, cap_bytes, rows, cols, step, nullptr, nullptr);
}

int b200s_pack_disparity_async(b200s_handle* h, int disp_id, void* dst, size_t cap_bytes, b200s_disparity_meta* meta,
                               b200s_done_fn done, void* user)
{
    if (!h || !dst) return B200S_EINVAL;
    DeviceGuard g(h->device);
    disp_id = fixed_point_plane_of(disp_id);
    Mat* D = find_mat(h, disp_id);
    if (!D || D->type != B200S_16SC1) return fail(h, B200S_ENOBUF, "disparity buffer is empty or not CV_16SC1");
    size_t n = (size_t)D->rows * D->cols;
    if (cap_bytes < n * 4) return fail(h, B200S_EINVAL, "destination too small");
    int rc = ensure_misc(h, h->w0);
    if (rc) return rc;
    cudaStream_t st = stream_of(h, disp_id);
    void* direct = h->pack_direct ? mapped_alias(dst) : nullptr;
    if (!direct && h->w0.df.ensure(n * 4)) return fail(h, B200S_ENOMEM, "cudaMalloc failed");
    h->launches += launch_disparity_to_float((const int16_t*)D->buf.p, direct ? (float*)direct : (float*)h->w0.df.p, (int)n,
                                             h->model_ok ? h->cxd : 0.0, (int*)h->w0.misc.p, st);
    if (meta) fill_disparity_meta(h, D->rows, D->cols, meta);
    rc = check_kernels(h, "pack_disparity");
    if (rc) return rc;
    return finish_pack(h, st, dst, direct ? nullptr : h->w0.df.p, n * 4, done, user);
}

int b200s_pack_disparity(b200s_handle* h, int disp_id, void* dst, size_t cap_bytes, b200s_disparity_meta* meta)
{
    return b200s_pack_disparity_async(h, disp_id, dst, cap_bytes, meta, nullptr, nullptr);
}

int b200s_pack_pointcloud2_async(b200s_handle* h, int disp_id, int color_id, void* dst, size_t cap_bytes, b200s_pc2_meta* meta,
                                 b200s_done_fn done, void* user)
{
    if (!h || !dst) return B200S_EINVAL;
    DeviceGuard g(h->device);
    if (!h->model_ok) return fail(h, B200S_ENOTINIT, "stereo model not initialised");
    disp_id = fixed_point_plane_of(disp_id);     // the reference passes its POINTS2 buffer (src/StereoProcessor.cpp:281)
    Mat* D = find_mat(h, disp_id);
    Mat* Cc = find_mat(h, color_id);
    if (!D || D->type != B200S_16SC1) return fail(h, B200S_ENOBUF, "disparity buffer is empty or not CV_16SC1");
    if (!Cc || (Cc->type != B200S_8UC1 && Cc->type != B200S_8UC3) || Cc->rows != D->rows || Cc->cols != D->cols)
        return fail(h, B200S_ENOBUF, "colour buffer is empty, not 8UC1/8UC3, or of a different size");
    size_t n = (size_t)D->rows * D->cols;
    if (cap_bytes < n * 32) return fail(h, B200S_EINVAL, "destination too small");
    int rc = ensure_misc(h, h->w0);
    if (rc) return rc;
    void* direct = h->pack_direct ? mapped_alias(dst) : nullptr;
    if (!direct && h->w0.pc2.ensure(n * 32)) return fail(h, B200S_ENOMEM, "cudaMalloc failed");
    cudaStream_t st = h->l_strm;
    CUDA_OK(h, cudaEventRecord(h->ev_r, h->r_strm));
    CUDA_OK(h, cudaStreamWaitEvent(st, h->ev_r, 0));
    h->launches += launch_disparity_to_float((const int16_t*)D->buf.p, nullptr, (int)n, h->cxd, (int*)h->w0.misc.p, st);
    h->launches += launch_reproject_pack((const int16_t*)D->buf.p, D->cols, D->rows, h->cxd, (const double*)h->Qdev.p, h->qmask,
                                         (const int*)h->w0.misc.p, (const uint8_t*)Cc->buf.p, Cc->type == B200S_8UC3 ? 3 : 1,
                                         nullptr, direct ? (uint8_t*)direct : (uint8_t*)h->w0.pc2.p, st);
    CUDA_OK(h, cudaEventRecord(h->ev_l, st));
    if (meta) fill_pc2_meta(D->rows, D->cols, meta);
    rc = check_kernels(h, "pack_pointcloud2");
    if (rc) return rc;
    return finish_pack(h, st, dst, direct ? nullptr : h->w0.pc2.p, n * 32, done, user);
}

int b200s_pack_pointcloud2(b200s_handle* h, int disp_id, int color_id, void* dst, size_t cap_bytes, b200s_pc2_meta* meta)
{
    return b200s_pack_pointcloud2_async(h, disp_id, color_id, dst, cap_bytes, meta, nullptr, nullptr);
}

// convertColor(src, dst, src_encoding, dst_encoding) (src/GPUStereoProcessor.cpp:119-172) for the encodings on the hot
// path: mono8 <-> bgr8 / rgb8 / bgra8 source, mono8 or bgr8 destination
int b200s_convert_color(b200s_handle* h, int src_id, int dst_id, const char* src_encoding, const char* dst_encoding)
{
    if (!h || !src_encoding || !dst_encoding) return B200S_EINVAL;
    DeviceGuard g(h->device);
    Mat* src = find_mat(h, src_id);
    if (!src) return fail(h, B200S_ENOBUF, "source buffer is empty");
    const std::string se = src_encoding, de = dst_encoding;
    const bool s_mono = se == "mono8", s_bgr = se == "bgr8", s_rgb = se == "rgb8";
    const bool d_mono = de == "mono8", d_bgr = de == "bgr8";
    if (!(s_mono || s_bgr || s_rgb) || !(d_mono || d_bgr))
        return fail(h, B200S_EUNSUPPORTED, "convertColor: '" + se + "' -> '" + de + "' is outside the hot path (mono8, bgr8, rgb8 -> mono8, bgr8)");
    if ((s_mono && src->type != B200S_8UC1) || (!s_mono && src->type != B200S_8UC3))
        return fail(h, B200S_EINVAL, "convertColor: buffer type does not match the source encoding");
    cudaStream_t st = stream_of(h, src_id);
    CUDA_OK(h, order_after_left_readers(h, st));
    const int rows = src->rows, cols = src->cols, n = rows * cols;
    Mat* dst;
    int rc = alloc_mat(h, dst_id, rows, cols, d_mono ? B200S_8UC1 : B200S_8UC3, de.c_str(), &dst);
    if (rc) return rc;
    src = find_mat(h, src_id);
    if (s_mono && d_mono) CUDA_OK(h, cudaMemcpyAsync(dst->buf.p, src->buf.p, src->bytes(), cudaMemcpyDeviceToDevice, st));
    else if (s_mono) h->launches += launch_gray_to_bgr((const uint8_t*)src->buf.p, (uint8_t*)dst->buf.p, n, st);
    else if (d_mono) h->launches += launch_bgr_to_gray((const uint8_t*)src->buf.p, (uint8_t*)dst->buf.p, n, s_rgb ? 1 : 0, st);
    else if (s_rgb) h->launches += launch_swap_rb((const uint8_t*)src->buf.p, (uint8_t*)dst->buf.p, n, st);
    else CUDA_OK(h, cudaMemcpyAsync(dst->buf.p, src->buf.p, src->bytes(), cudaMemcpyDeviceToDevice, st));
    return check_kernels(h, "convert_color");
}

// printStats (src/GPUStereoProcessor.cpp:421-435) on a named buffer
int b200s_mat_stats(b200s_handle* h, int mat_id, double* mn, double* mx, double* mean, int* channels)
{
    if (!h) return B200S_EINVAL;
    DeviceGuard g(h->device);
    Mat* m = find_mat(h, mat_id);
    if (!m) return fail(h, B200S_ENOBUF, "named buffer is empty");
    int ch = 1, kind = 0;
    switch (m->type) {
        case B200S_8UC1: ch = 1; kind = 0; break;
        case B200S_8UC3: ch = 3; kind = 0; break;
        case B200S_8UC4: ch = 4; kind = 0; break;
        case B200S_16SC1: ch = 1; kind = 1; break;
        case B200S_32FC1: ch = 1; kind = 2; break;
        case B200S_32FC3: ch = 3; kind = 2; break;
        default: return fail(h, B200S_EUNSUPPORTED, "unsupported element type");
    }
    const int nb = 128;
    DevBuf part;
    if (part.ensure((size_t)nb * ch * 3 * sizeof(double))) return fail(h, B200S_ENOMEM, "cudaMalloc failed");
    cudaStream_t st = stream_of(h, mat_id);
    const size_t npix = (size_t)m->rows * m->cols;
    h->launches += launch_mat_stats(m->buf.p, kind, npix, ch, (double*)part.p, nb, st);
    std::vector<double> hp((size_t)nb * ch * 3);
    cudaError_t e = cudaMemcpyAsync(hp.data(), part.p, hp.size() * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    part.release();
    if (e != cudaSuccess) return fail(h, B200S_ECUDA, cudaGetErrorString(e));
    for (int c = 0; c < ch; ++c) {
        double a = 1e300, b = -1e300, s = 0;
        for (int k = 0; k < nb; ++k) {
            const double* o = &hp[((size_t)k * ch + c) * 3];
            a = std::min(a, o[0]); b = std::max(b, o[1]); s += o[2];
        }
        if (mn) mn[c] = a;
        if (mx) mx[c] = b;
        if (mean) mean[c] = s / (double)npix;
    }
    if (channels) *channels = ch;
    return check_kernels(h, "mat_stats");
}

// ---- instrumentation -------------------------------------------------------------------------------------
uint64_t b200s_kernel_launches(const b200s_handle* h) { return h ? h->launches : 0; }

int b200s_last_bm_time(b200s_handle* h, int slot, float* ms, double* evals)
{
    if (!h || !ms) return B200S_EINVAL;
    Work* w = slot < 0 ? &h->w0 : (slot < (int)h->slots.size() ? &h->slots[slot] : nullptr);
    if (!w || !w->timed) return fail(h, B200S_ENOBUF, "no timed matcher run on this slot (b200s_enable_timing)");
    DeviceGuard g(h->device);
    CUDA_OK(h, cudaEventSynchronize(w->ev_bm1));
    CUDA_OK(h, cudaEventElapsedTime(ms, w->ev_bm0, w->ev_bm1));
    if (evals) *evals = w->last_evals;
    return B200S_OK;
}

int b200s_host_alloc(void** p, size_t bytes)
{
    if (!p) return B200S_EINVAL;
    return cudaHostAlloc(p, bytes, cudaHostAllocPortable) == cudaSuccess ? B200S_OK : B200S_ENOMEM;   // pinned for every GPU of the process
}

int b200s_host_alloc_mode(void** p, size_t bytes, int mode)
{
    if (!p) return B200S_EINVAL;
    return cudaHostAlloc(p, bytes, cudaHostAllocPortable | (mode == 1 ? cudaHostAllocWriteCombined : 0)) == cudaSuccess ? B200S_OK : B200S_ENOMEM;
}

int b200s_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? B200S_OK : B200S_ECUDA; }

int b200s_int_peak(b200s_handle* h, int which, double* lane_ops_per_s, double* sm_clock_mhz)
{
    if (!h || !lane_ops_per_s) return B200S_EINVAL;
    DeviceGuard g(h->device);
    int rc = run_int_peak(which, lane_ops_per_s, sm_clock_mhz, h->l_strm);
    if (rc) return fail(h, B200S_ECUDA, "int peak micro-benchmark failed");
    return B200S_OK;
}

}  // extern "C"
