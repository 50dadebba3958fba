// Host pipeline + extern "C" layer of libb200stereo.so (declared in include/b200_stereo.h).
// Mirrors the state and method set of gpuimageproc::GpuStereoProcessor
// (reference: include/gpuimageproc/GPUStereoProcessor.h:63-126, src/GPUStereoProcessor.cpp) with the
// OpenCV / image_geometry calls replaced by the sm_100a kernels in this directory.
#include "../../include/b200_stereo.h"
#include "kernels.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

using namespace b200s;

namespace {

int elem_size(int type)
{
    switch (type) {
        case B200S_8UC1: return 1;
        case B200S_16SC1: return 2;
        case B200S_32FC1: return 4;
        case B200S_8UC3: return 3;
        case B200S_32FC3: return 12;
        case B200S_8UC4: return 4;
        default: return 0;
    }
}

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        if (cudaMalloc(&p, bytes) != cudaSuccess) return -1;
        cap = bytes;
        return 0;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct Mat {
    DevBuf buf;
    int rows = 0, cols = 0, type = -1;
    std::string enc;
    size_t bytes() const { return (size_t)rows * cols * elem_size(type); }
    bool empty() const { return type < 0 || rows == 0 || cols == 0; }
};

// all device scratch one pipeline instance needs
struct Work {
    cudaStream_t st = nullptr;
    bool own_stream = false;
    DevBuf rawL, rawR, rectL, rectR, preL, preR, disp, cost, df, xyz, pc2, vol, ccl, normtmp, misc;
    cudaEvent_t ev_bm0 = nullptr, ev_bm1 = nullptr, ev_done = nullptr;
    double last_evals = 0;
    bool timed = false;
    // CUDA graph of the frame chain of this slot: the first frame with a given key runs eagerly (allocations, map
    // build), the second is captured, the following ones are replayed with one cudaGraphLaunch
    std::string gkey, warm_key;
    cudaGraphExec_t gexec = nullptr;
    uint64_t glaunches = 0;
    double gevals = 0;
    void drop_graph()
    {
        if (gexec) cudaGraphExecDestroy(gexec);
        gexec = nullptr;
        gkey.clear();
        warm_key.clear();
    }
    void release()
    {
        drop_graph();
        DevBuf* all[] = {&rawL, &rawR, &rectL, &rectR, &preL, &preR, &disp, &cost, &df, &xyz, &pc2, &vol, &ccl, &normtmp, &misc};
        for (DevBuf* b : all) b->release();
        if (ev_bm0) cudaEventDestroy(ev_bm0);
        if (ev_bm1) cudaEventDestroy(ev_bm1);
        if (ev_done) cudaEventDestroy(ev_done);
        if (own_stream && st) cudaStreamDestroy(st);
        ev_bm0 = ev_bm1 = ev_done = nullptr;
        st = nullptr;
    }
};

struct Camera {
    b200s_caminfo info;
    CamModel cm;
    DevBuf map;   // int2 per pixel
    bool map_valid = false;
};

}  // namespace

struct b200s_handle {
    int device = 0;
    std::string err;
    b200s_params prm;
    bool model_ok = false;
    bool rect_fly = false;
    bool timing = false;
    Camera cam[2];
    double Q[16];
    double baseline = 0, fx_right = 0, cxd = 0;
    unsigned qmask = 0xFFFFu;
    DevBuf Qdev;
    std::unordered_map<int, Mat> mats;
    cudaStream_t l_strm = nullptr, r_strm = nullptr;
    cudaEvent_t ev_r = nullptr;
    Work w0;                  // scratch of the named-buffer API (runs on l_strm)
    std::vector<Work> slots;
    std::vector<cudaEvent_t> batch_end;
    cudaEvent_t batch_start = nullptr;
    int slot_rows = 0, slot_cols = 0;
    uint64_t launches = 0;
    void* stage_host[2] = {nullptr, nullptr};   // pinned staging for device->host copies into pageable user memory
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    uint64_t model_version = 0;   // bumped by every calibration change (part of the graph key)
    int use_graphs = 1;           // B200S_GRAPH=0 or b200s_set_graph_mode(h, 0) turns the replay off
    uint64_t graph_replays = 0;
    int pack_direct = 0;          // 1: the pack kernels store straight into pinned (mapped) host destinations
};

namespace {

int fail(b200s_handle* h, int code, const std::string& msg)
{
    if (h) h->err = msg;
    return code;
}

#define CUDA_OK(h, call)                                                                        \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return fail(h, B200S_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__));  \
    } while (0)

struct DeviceGuard {
    explicit DeviceGuard(int d) { cudaSetDevice(d); }
};

int check_kernels(b200s_handle* h, const char* what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, B200S_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return B200S_OK;
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

int ensure_map(b200s_handle* h, int side /*0 L, 1 R*/, cudaStream_t st)
{
    Camera& c = h->cam[side];
    if (c.map_valid) return B200S_OK;
    size_t n = (size_t)c.info.width * c.info.height;
    if (c.map.ensure(n * sizeof(int2)) != 0) return fail(h, B200S_ENOMEM, "cudaMalloc failed for the rectification map");
    h->launches += launch_build_map(c.cm, c.info.width, c.info.height, (int2*)c.map.p, st);
    // the map may be consumed on another stream: make it visible everywhere once
    CUDA_OK(h, cudaStreamSynchronize(st));
    c.map_valid = true;
    return check_kernels(h, "build_map");
}

// pitched prefiltered planes with readable slack on both sides (zeroed once so that stray reads are defined)
int ensure_pre_planes(b200s_handle* h, Work& w, int rows, int cols)
{
    size_t need = plane_bytes(cols, rows);
    for (DevBuf* b : {&w.preL, &w.preR}) {
        if (b->cap < need) {
            if (b->ensure(need)) return fail(h, B200S_ENOMEM, "cudaMalloc failed (prefilter planes)");
            if (cudaMemset(b->p, 0, need) != cudaSuccess) return fail(h, B200S_ECUDA, "cudaMemset failed (prefilter planes)");
        }
    }
    return B200S_OK;
}

// prefilter + match + post-filters on rectified device planes; disp must hold rows*cols int16
int run_disparity(b200s_handle* h, Work& w, const uint8_t* L, const uint8_t* R, bool already_prefiltered, int rows,
                  int cols, int16_t* disp, cudaStream_t st)
{
    const b200s_params& p = h->prm;
    int rc = validate_params(h, p);
    if (rc) return rc;
    if (p.block_size >= (rows < cols ? rows : cols))
        return fail(h, B200S_EINVAL, "SADWindowSize must be odd, be within 5..255 and be not larger than image width or height");
    size_t n = (size_t)rows * cols;
    BMConfig cfg = bm_config(p);
    const size_t pitch = plane_pitch(cols);
    if (!already_prefiltered) {
        int rc2 = ensure_pre_planes(h, w, rows, cols);
        if (rc2) return rc2;
        uint8_t* pl = (uint8_t*)w.preL.p + PLANE_LEAD;
        uint8_t* pr = (uint8_t*)w.preR.p + PLANE_LEAD;
        if (p.pre_filter_type == 1) {
            h->launches += launch_prefilter_xsobel(L, pl, pitch, cols, rows, p.pre_filter_cap, st);
            h->launches += launch_prefilter_xsobel(R, pr, pitch, cols, rows, p.pre_filter_cap, st);
        } else {
            const int one = launch_norm_prefilter_pair(L, R, cols, rows, false, nullptr, nullptr, h->cam[0].cm, h->cam[1].cm, nullptr, nullptr,
                                                       pl, pr, pitch, cols, rows, p.pre_filter_size, p.pre_filter_cap, st);
            h->launches += one;
            if (!one) {     // preFilterSize > 21: two passes through a scratch plane
                if (w.normtmp.ensure(n * sizeof(int))) return fail(h, B200S_ENOMEM, "cudaMalloc failed (prefilter scratch)");
                h->launches += launch_prefilter_norm(L, pl, pitch, cols, rows, p.pre_filter_size, p.pre_filter_cap, (int*)w.normtmp.p, st);
                h->launches += launch_prefilter_norm(R, pr, pitch, cols, rows, p.pre_filter_size, p.pre_filter_cap, (int*)w.normtmp.p, st);
            }
        }
    }
    // the matcher always reads the pitched, slack-padded prefiltered planes of this Work
    const uint8_t* Lp = (const uint8_t*)w.preL.p + PLANE_LEAD;
    const uint8_t* Rp = (const uint8_t*)w.preR.p + PLANE_LEAD;
    (void)L; (void)R;
    int16_t* cost = nullptr;
    if (cfg.disp12MaxDiff >= 0) {
        if (w.cost.ensure(n * sizeof(int16_t))) return fail(h, B200S_ENOMEM, "cudaMalloc failed (cost plane)");
        cost = (int16_t*)w.cost.p;
    }
    if (w.vol.ensure(bm_scratch_bytes(cols, rows, cfg))) return fail(h, B200S_ENOMEM, "cudaMalloc failed (matcher scratch)");
    BMScratch sc{(int*)w.vol.p, w.vol.cap};
    if (h->timing) {
        if (!w.ev_bm0) { cudaEventCreate(&w.ev_bm0); cudaEventCreate(&w.ev_bm1); }
        cudaEventRecord(w.ev_bm0, st);
    }
    int l = launch_block_match(Lp, Rp, pitch, cols, rows, cfg, disp, cost, &sc, st, &w.last_evals);
    if (h->timing) { cudaEventRecord(w.ev_bm1, st); w.timed = true; }
    if (l < 0) return fail(h, B200S_ECUDA, "block matcher launch failed (code " + std::to_string(l) + "): " + cudaGetErrorString(cudaGetLastError()));
    h->launches += l;
    if (cfg.disp12MaxDiff >= 0) {
        int v = launch_validate_disp12(disp, cost, cols, rows, cfg, st);
        if (v < 0) return fail(h, B200S_EUNSUPPORTED, "disp12MaxDiff needs image width <= 65535");
        h->launches += v;     // the valid-ROI mask of the rows is applied by the same kernel
    }
    if (p.speckle_window_size > 0 && p.speckle_range >= 0) {
        if (w.ccl.ensure(3 * n * sizeof(int))) return fail(h, B200S_ENOMEM, "cudaMalloc failed (speckle scratch)");
        h->launches += launch_filter_speckles(disp, cols, rows, (p.min_disparity - 1) * 16, p.speckle_window_size,
                                              p.speckle_range, (int*)w.ccl.p, st);
    }
    return check_kernels(h, "disparity chain");
}

int ensure_misc(b200s_handle* h, Work& w)
{
    if (w.misc.ensure(256)) return fail(h, B200S_ENOMEM, "cudaMalloc failed (misc)");
    return B200S_OK;
}

// Device-side alias of a pinned (page-locked, mapped) host buffer, or nullptr when `p` is pageable / not host memory.
// Under unified addressing every cudaHostAlloc / cudaHostRegister range is directly writable by kernels.
void* mapped_alias(const void* p)
{
    if (!p) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

int copy_out(b200s_handle* h, void* dst, const void* src, size_t bytes, bool dst_on_device, cudaStream_t st)
{
    if (!dst) return B200S_OK;
    CUDA_OK(h, cudaMemcpyAsync(dst, src, bytes, dst_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    return B200S_OK;
}

}  // namespace

// ==========================================================================================================
extern "C" {

const char* b200s_version(void) { return "b200-stereo 0.1 (sm_100a)"; }

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
    if (cudaSetDevice(device) != cudaSuccess) return B200S_ECUDA;
    b200s_handle* h = new (std::nothrow) b200s_handle();
    if (!h) return B200S_ENOMEM;
    h->device = device;
    // state of the reference's CPU matcher after its constructor (src/GPUStereoProcessor.cpp:18-38):
    // createStereoBM(48, 19) mirrored into cv::StereoBM, preFilterSize 5
    if (const char* e = getenv("B200S_GRAPH")) h->use_graphs = atoi(e) != 0;
    if (const char* e = getenv("B200S_PACK_DIRECT")) h->pack_direct = atoi(e);
    b200s_default_params(&h->prm);
    h->prm.pre_filter_type = 0; h->prm.pre_filter_size = 5; h->prm.num_disparities = 48; h->prm.block_size = 19;
    h->prm.texture_threshold = 3; h->prm.uniqueness_ratio = 0; h->prm.disp12_max_diff = 0;
    if (cudaStreamCreateWithFlags(&h->l_strm, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->r_strm, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_r, cudaEventDisableTiming) != cudaSuccess) {
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
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (auto& kv : h->mats) kv.second.buf.release();
    for (Work& w : h->slots) w.release();
    h->w0.release();
    for (int s = 0; s < 2; ++s) h->cam[s].map.release();
    h->Qdev.release();
    if (h->ev_r) cudaEventDestroy(h->ev_r);
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
    const int2* map = nullptr;
    if (!h->rect_fly && interpolation == B200S_INTER_LINEAR) {
        rc = ensure_map(h, side, st);
        if (rc) return rc;
        map = (const int2*)h->cam[side].map.p;
    }
    if (interpolation == B200S_INTER_LINEAR)
        h->launches += launch_remap((const uint8_t*)src->buf.p, sW, sH, ch, map, cam.cm, (uint8_t*)dst->buf.p, W, H, st);
    else
        h->launches += launch_remap_nearest((const uint8_t*)src->buf.p, sW, sH, ch, map, cam.cm, (uint8_t*)dst->buf.p, W, H, st);
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
int b200s_pack_image(b200s_handle* h, int mat_id, void* dst, size_t cap_bytes, int* rows, int* cols, int* step)
{
    if (!h || !dst) return B200S_EINVAL;
    DeviceGuard g(h->device);
    Mat* m = find_mat(h, mat_id);
    if (!m) return fail(h, B200S_ENOBUF, "named buffer is empty");
    if (cap_bytes < m->bytes()) return fail(h, B200S_EINVAL, "destination too small");
    cudaStream_t st = stream_of(h, mat_id);
    { int rcd = d2h_sync(h, dst, m->buf.p, m->bytes(), st); if (rcd) return rcd; }
    if (rows) *rows = m->rows;
    if (cols) *cols = m->cols;
    if (step) *step = m->cols * elem_size(m->type);   // GpuSenderImage.cpp:20: width * bitdepth * channels / 8
    return B200S_OK;
}

static void fill_disparity_meta(const b200s_handle* h, int rows, int cols, b200s_disparity_meta* m)
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

int b200s_pack_disparity(b200s_handle* h, int disp_id, void* dst, size_t cap_bytes, b200s_disparity_meta* meta)
{
    if (!h || !dst) return B200S_EINVAL;
    DeviceGuard g(h->device);
    Mat* D = find_mat(h, disp_id);
    if (!D || D->type != B200S_16SC1) return fail(h, B200S_ENOBUF, "disparity buffer is empty or not CV_16SC1");
    size_t n = (size_t)D->rows * D->cols;
    if (cap_bytes < n * 4) return fail(h, B200S_EINVAL, "destination too small");
    int rc = ensure_misc(h, h->w0);
    if (rc) return rc;
    if (h->w0.df.ensure(n * 4)) return fail(h, B200S_ENOMEM, "cudaMalloc failed");
    cudaStream_t st = stream_of(h, disp_id);
    h->launches += launch_disparity_to_float((const int16_t*)D->buf.p, (float*)h->w0.df.p, (int)n, h->model_ok ? h->cxd : 0.0, (int*)h->w0.misc.p, st);
    { int rcd = d2h_sync(h, dst, h->w0.df.p, n * 4, st); if (rcd) return rcd; }
    if (meta) fill_disparity_meta(h, D->rows, D->cols, meta);
    return check_kernels(h, "pack_disparity");
}

static void fill_pc2_meta(int rows, int cols, b200s_pc2_meta* m)
{
    m->width = cols; m->height = rows; m->point_step = 32; m->row_step = 32 * cols;
    m->is_bigendian = 0; m->is_dense = 0;
    m->off_x = 0; m->off_y = 4; m->off_z = 8; m->off_rgb = 16;
}

int b200s_pack_pointcloud2(b200s_handle* h, int disp_id, int color_id, void* dst, size_t cap_bytes, b200s_pc2_meta* meta)
{
    if (!h || !dst) return B200S_EINVAL;
    DeviceGuard g(h->device);
    if (!h->model_ok) return fail(h, B200S_ENOTINIT, "stereo model not initialised");
    Mat* D = find_mat(h, disp_id);
    Mat* Cc = find_mat(h, color_id);
    if (!D || D->type != B200S_16SC1) return fail(h, B200S_ENOBUF, "disparity buffer is empty or not CV_16SC1");
    if (!Cc || (Cc->type != B200S_8UC1 && Cc->type != B200S_8UC3) || Cc->rows != D->rows || Cc->cols != D->cols)
        return fail(h, B200S_ENOBUF, "colour buffer is empty, not 8UC1/8UC3, or of a different size");
    size_t n = (size_t)D->rows * D->cols;
    if (cap_bytes < n * 32) return fail(h, B200S_EINVAL, "destination too small");
    int rc = ensure_misc(h, h->w0);
    if (rc) return rc;
    if (h->w0.pc2.ensure(n * 32)) return fail(h, B200S_ENOMEM, "cudaMalloc failed");
    cudaStream_t st = h->l_strm;
    CUDA_OK(h, cudaEventRecord(h->ev_r, h->r_strm));
    CUDA_OK(h, cudaStreamWaitEvent(st, h->ev_r, 0));
    h->launches += launch_disparity_to_float((const int16_t*)D->buf.p, nullptr, (int)n, h->cxd, (int*)h->w0.misc.p, st);
    h->launches += launch_reproject_pack((const int16_t*)D->buf.p, D->cols, D->rows, h->cxd, (const double*)h->Qdev.p, h->qmask,
                                         (const int*)h->w0.misc.p, (const uint8_t*)Cc->buf.p, Cc->type == B200S_8UC3 ? 3 : 1,
                                         nullptr, (uint8_t*)h->w0.pc2.p, st);
    { int rcd = d2h_sync(h, dst, h->w0.pc2.p, n * 32, st); if (rcd) return rcd; }
    if (meta) fill_pc2_meta(D->rows, D->cols, meta);
    return check_kernels(h, "pack_pointcloud2");
}

// ---- fused frame path ------------------------------------------------------------------------------------
int b200s_configure_slots(b200s_handle* h, int n_slots, int rows, int cols)
{
    if (!h || n_slots < 1 || n_slots > 64 || rows <= 0 || cols <= 0) return B200S_EINVAL;
    DeviceGuard g(h->device);
    cudaDeviceSynchronize();
    for (Work& w : h->slots) w.release();
    h->slots.clear();
    h->slots.resize(n_slots);
    for (Work& w : h->slots) {
        if (cudaStreamCreateWithFlags(&w.st, cudaStreamNonBlocking) != cudaSuccess) return fail(h, B200S_ECUDA, "cudaStreamCreate failed");
        w.own_stream = true;
        cudaEventCreateWithFlags(&w.ev_done, cudaEventDisableTiming);
        cudaEventCreate(&w.ev_bm0);
        cudaEventCreate(&w.ev_bm1);
    }
    h->slot_rows = rows;
    h->slot_cols = cols;
    return B200S_OK;
}

namespace {

// everything of a frame after the inputs are on the device: rectify -> disparity -> float / reproject+pack -> outputs
int run_frame_chain(b200s_handle* h, Work& w, const b200s_frame_io* io, const uint8_t* L, const uint8_t* R, cudaStream_t st)
{
    const int rows = h->slot_rows, cols = h->slot_cols;
    const size_t n = (size_t)rows * cols;
    bool prefiltered = false;
    const uint8_t *rl = L, *rr = R;
    if (io->rectify) {
        if (w.rectL.ensure(n + 64) || w.rectR.ensure(n + 64)) return fail(h, B200S_ENOMEM, "cudaMalloc failed (rectified planes)");
        const int2 *mapL = nullptr, *mapR = nullptr;
        if (!h->rect_fly) {
            int rc = ensure_map(h, 0, st);
            if (rc) return rc;
            rc = ensure_map(h, 1, st);
            if (rc) return rc;
            mapL = (const int2*)h->cam[0].map.p;
            mapR = (const int2*)h->cam[1].map.p;
        }
        if (h->prm.pre_filter_type == 1) {
            int rc2 = ensure_pre_planes(h, w, rows, cols);
            if (rc2) return rc2;
            const size_t pitch = plane_pitch(cols);
            h->launches += launch_rectify_xsobel_pair(L, R, cols, rows, mapL, mapR, h->cam[0].cm, h->cam[1].cm, (uint8_t*)w.rectL.p,
                                                      (uint8_t*)w.rectR.p, (uint8_t*)w.preL.p + PLANE_LEAD,
                                                      (uint8_t*)w.preR.p + PLANE_LEAD, pitch, cols, rows, h->prm.pre_filter_cap, st);
            prefiltered = true;
        } else {
            int one = 0;
            if (h->prm.pre_filter_size <= 21) {
                int rc2 = ensure_pre_planes(h, w, rows, cols);
                if (rc2) return rc2;
                one = launch_norm_prefilter_pair(L, R, cols, rows, true, mapL, mapR, h->cam[0].cm, h->cam[1].cm, (uint8_t*)w.rectL.p,
                                                 (uint8_t*)w.rectR.p, (uint8_t*)w.preL.p + PLANE_LEAD, (uint8_t*)w.preR.p + PLANE_LEAD,
                                                 plane_pitch(cols), cols, rows, h->prm.pre_filter_size, h->prm.pre_filter_cap, st);
            }
            if (one) {
                h->launches += one;
                prefiltered = true;
            } else {
                h->launches += launch_remap(L, cols, rows, 1, mapL, h->cam[0].cm, (uint8_t*)w.rectL.p, cols, rows, st);
                h->launches += launch_remap(R, cols, rows, 1, mapR, h->cam[1].cm, (uint8_t*)w.rectR.p, cols, rows, st);
            }
        }
        rl = (const uint8_t*)w.rectL.p;
        rr = (const uint8_t*)w.rectR.p;
    }
    if (w.disp.ensure(n * 2 + 64)) return fail(h, B200S_ENOMEM, "cudaMalloc failed (disparity plane)");
    int rc = run_disparity(h, w, rl, rr, prefiltered, rows, cols, (int16_t*)w.disp.p, st);
    if (rc) return rc;
    const bool want_pc = io->want & B200S_OUT_POINTCLOUD2, want_xyz = io->want & B200S_OUT_POINTS_XYZ;
    const bool want_df = io->want & B200S_OUT_DISPARITY32F;
    if (want_df || want_pc || want_xyz) {
        rc = ensure_misc(h, w);
        if (rc) return rc;
        if (want_df && w.df.ensure(n * 4)) return fail(h, B200S_ENOMEM, "cudaMalloc failed (float disparity)");
        h->launches += launch_disparity_to_float((const int16_t*)w.disp.p, want_df ? (float*)w.df.p : nullptr, (int)n,
                                                 h->model_ok ? h->cxd : 0.0, (int*)w.misc.p, st);
    }
    // pack mode "direct" (the north star's wording): the PointCloud2 records are stored by the kernel straight into the
    // caller's pinned host buffer (PCIe posted writes), no HBM copy of the cloud and no copy-engine transfer afterwards
    void* pc_direct = (h->pack_direct && want_pc && !io->outputs_on_device) ? mapped_alias(io->pointcloud2) : nullptr;
    if (want_pc || want_xyz) {
        if (want_pc && !pc_direct && w.pc2.ensure(n * 32)) return fail(h, B200S_ENOMEM, "cudaMalloc failed (point cloud)");
        if (want_xyz && w.xyz.ensure(n * 12)) return fail(h, B200S_ENOMEM, "cudaMalloc failed (points)");
        h->launches += launch_reproject_pack((const int16_t*)w.disp.p, cols, rows, h->cxd, (const double*)h->Qdev.p, h->qmask,
                                             (const int*)w.misc.p, rl, 1, want_xyz ? (float*)w.xyz.p : nullptr,
                                             want_pc ? (pc_direct ? (uint8_t*)pc_direct : (uint8_t*)w.pc2.p) : nullptr, st);
    }
    rc = check_kernels(h, "process_pair");
    if (rc) return rc;
    const bool od = io->outputs_on_device != 0;
    if (io->rectify) {
        if ((io->want & B200S_OUT_RECT_L) && (rc = copy_out(h, io->rect_left, w.rectL.p, n, od, st))) return rc;
        if ((io->want & B200S_OUT_RECT_R) && (rc = copy_out(h, io->rect_right, w.rectR.p, n, od, st))) return rc;
    }
    if ((io->want & B200S_OUT_DISPARITY16) && (rc = copy_out(h, io->disparity16, w.disp.p, n * 2, od, st))) return rc;
    if (want_df && (rc = copy_out(h, io->disparity32f, w.df.p, n * 4, od, st))) return rc;
    if (want_pc && !pc_direct && (rc = copy_out(h, io->pointcloud2, w.pc2.p, n * 32, od, st))) return rc;
    if (want_xyz && (rc = copy_out(h, io->points_xyz, w.xyz.p, n * 12, od, st))) return rc;
    return B200S_OK;
}

// everything a captured chain depends on besides the (fixed) slot buffers
std::string frame_graph_key(const b200s_handle* h, const b200s_frame_io* io)
{
    std::string k;
    auto add = [&k](const void* p, size_t n) { k.append((const char*)p, n); };
    add(&h->prm, sizeof h->prm);
    add(io, sizeof *io);
    add(&h->model_version, sizeof h->model_version);
    add(&h->slot_rows, sizeof h->slot_rows);
    add(&h->slot_cols, sizeof h->slot_cols);
    return k;
}

}  // namespace

int b200s_process_pair_async(b200s_handle* h, int slot, const void* left, const void* right, const b200s_frame_io* io)
{
    if (!h || !left || !right || !io) return B200S_EINVAL;
    DeviceGuard g(h->device);
    if (slot < 0 || slot >= (int)h->slots.size()) return fail(h, B200S_EINVAL, "slot out of range (call b200s_configure_slots)");
    Work& w = h->slots[slot];
    const int rows = h->slot_rows, cols = h->slot_cols;
    const size_t n = (size_t)rows * cols;
    cudaStream_t st = w.st;
    const bool need_model = io->rectify || (io->want & (B200S_OUT_POINTCLOUD2 | B200S_OUT_POINTS_XYZ));
    if (need_model && !h->model_ok) return fail(h, B200S_ENOTINIT, "stereo model not initialised");
    if (io->rectify && (h->cam[0].info.width != cols || h->cam[0].info.height != rows))
        return fail(h, B200S_EINVAL, "slot size differs from the calibration size");
    const bool graphs = h->use_graphs && !h->timing;
    // inputs: host frames always go through the slot's raw planes; with graph replay device frames do too, so that
    // the captured kernels see fixed addresses
    const uint8_t *L = (const uint8_t*)left, *R = (const uint8_t*)right;
    if (!io->inputs_on_device || graphs) {
        if (w.rawL.ensure(n + 64) || w.rawR.ensure(n + 64)) return fail(h, B200S_ENOMEM, "cudaMalloc failed (input planes)");
        const cudaMemcpyKind kind = io->inputs_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        CUDA_OK(h, cudaMemcpyAsync(w.rawL.p, left, n, kind, st));
        CUDA_OK(h, cudaMemcpyAsync(w.rawR.p, right, n, kind, st));
        L = (const uint8_t*)w.rawL.p;
        R = (const uint8_t*)w.rawR.p;
    }
    int rc = B200S_OK;
    if (!graphs) {
        rc = run_frame_chain(h, w, io, L, R, st);
    } else {
        const std::string key = frame_graph_key(h, io);
        if (w.gexec && key == w.gkey) {
            CUDA_OK(h, cudaGraphLaunch(w.gexec, st));
            h->launches += w.glaunches;
            w.last_evals = w.gevals;
            ++h->graph_replays;
        } else if (key == w.warm_key) {
            // second frame with this key: every buffer exists, the maps are built -> capture, instantiate, launch
            if (w.gexec) { cudaGraphExecDestroy(w.gexec); w.gexec = nullptr; w.gkey.clear(); }
            const uint64_t l0 = h->launches;
            cudaGraph_t graph = nullptr;
            cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed);
            if (e == cudaSuccess) {
                rc = run_frame_chain(h, w, io, L, R, st);
                e = cudaStreamEndCapture(st, &graph);
                if (rc == B200S_OK && e == cudaSuccess && graph) e = cudaGraphInstantiate(&w.gexec, graph, 0);
                if (graph) cudaGraphDestroy(graph);
            }
            if (rc != B200S_OK || e != cudaSuccess || !w.gexec) {
                // capture is an optimisation only: fall back to eager launches for this handle
                cudaGetLastError();
                w.drop_graph();
                h->use_graphs = 0;
                h->launches = l0;
                rc = run_frame_chain(h, w, io, L, R, st);
            } else {
                w.glaunches = h->launches - l0;
                w.gevals = w.last_evals;
                w.gkey = key;
                h->launches = l0 + w.glaunches;
                CUDA_OK(h, cudaGraphLaunch(w.gexec, st));
            }
        } else {
            rc = run_frame_chain(h, w, io, L, R, st);
            w.warm_key = rc == B200S_OK ? key : std::string();
        }
    }
    if (rc) return rc;
    CUDA_OK(h, cudaEventRecord(w.ev_done, st));
    return B200S_OK;
}

int b200s_set_graph_mode(b200s_handle* h, int on)
{
    if (!h) return B200S_EINVAL;
    h->use_graphs = on ? 1 : 0;
    for (Work& w : h->slots) w.drop_graph();
    return B200S_OK;
}

uint64_t b200s_graph_replays(const b200s_handle* h) { return h ? h->graph_replays : 0; }

int b200s_wait_slot(b200s_handle* h, int slot)
{
    if (!h) return B200S_EINVAL;
    if (slot < 0 || slot >= (int)h->slots.size()) return fail(h, B200S_EINVAL, "slot out of range");
    DeviceGuard g(h->device);
    CUDA_OK(h, cudaStreamSynchronize(h->slots[slot].st));
    return B200S_OK;
}

// non-blocking completion test of a slot's last frame (the reference publishes from a stream callback, GpuSenderIfc.cpp:13-26)
int b200s_poll_slot(b200s_handle* h, int slot, int* done)
{
    if (!h || !done) return B200S_EINVAL;
    if (slot < 0 || slot >= (int)h->slots.size()) return fail(h, B200S_EINVAL, "slot out of range");
    DeviceGuard g(h->device);
    cudaError_t e = cudaEventQuery(h->slots[slot].ev_done);
    if (e != cudaSuccess && e != cudaErrorNotReady) return fail(h, B200S_ECUDA, cudaGetErrorString(e));
    *done = e == cudaSuccess;
    return B200S_OK;
}

int b200s_slot_device_ptr(b200s_handle* h, int slot, uint32_t which, void** dptr, size_t* bytes)
{
    if (!h || !dptr) return B200S_EINVAL;
    if (slot < 0 || slot >= (int)h->slots.size()) return fail(h, B200S_EINVAL, "slot out of range");
    Work& w = h->slots[slot];
    size_t n = (size_t)h->slot_rows * h->slot_cols;
    DevBuf* b = nullptr;
    size_t sz = 0;
    switch (which) {
        case B200S_OUT_RECT_L: b = &w.rectL; sz = n; break;
        case B200S_OUT_RECT_R: b = &w.rectR; sz = n; break;
        case B200S_OUT_DISPARITY16: b = &w.disp; sz = n * 2; break;
        case B200S_OUT_DISPARITY32F: b = &w.df; sz = n * 4; break;
        case B200S_OUT_POINTCLOUD2: b = &w.pc2; sz = n * 32; break;
        case B200S_OUT_POINTS_XYZ: b = &w.xyz; sz = n * 12; break;
        default: return fail(h, B200S_EINVAL, "unknown product");
    }
    if (!b->p) return fail(h, B200S_ENOBUF, "product has not been computed on this slot yet");
    *dptr = b->p;
    if (bytes) *bytes = sz;
    return B200S_OK;
}

int b200s_process_pair(b200s_handle* h, const void* left, const void* right, const b200s_frame_io* io)
{
    if (!h) return B200S_EINVAL;
    if (h->slots.empty()) return fail(h, B200S_EINVAL, "call b200s_configure_slots first");
    int rc = b200s_process_pair_async(h, 0, left, right, io);
    if (rc) return rc;
    return b200s_wait_slot(h, 0);
}

// ---- multi-GPU pool: one handle (stream set, slots, calibration, parameters) per GPU inside one process ----------
// Independent stereo frames shard over the GPUs with no exchange step: frame k -> GPU k mod N, slot
// (k div N) mod S.  All calls only enqueue work, so one host thread drives every GPU; with graph replay a submit is
// two copies and one graph launch.
struct b200s_pool {
    std::vector<b200s_handle*> h;
    int slots = 0;
    std::string err;
};

int b200s_pool_create(int n_gpus, const int* devices, int slots_per_gpu, int rows, int cols, b200s_pool** out)
{
    if (!out || n_gpus < 1 || n_gpus > 64 || slots_per_gpu < 1) return B200S_EINVAL;
    b200s_pool* p = new b200s_pool;
    p->slots = slots_per_gpu;
    for (int i = 0; i < n_gpus; ++i) {
        b200s_handle* h = nullptr;
        int rc = b200s_create(devices ? devices[i] : i, &h);
        if (rc == B200S_OK) {
            p->h.push_back(h);
            rc = b200s_configure_slots(h, slots_per_gpu, rows, cols);
        }
        if (rc != B200S_OK) {
            for (b200s_handle* q : p->h) b200s_destroy(q);
            delete p;
            return rc;
        }
    }
    *out = p;
    return B200S_OK;
}

int b200s_pool_destroy(b200s_pool* p)
{
    if (!p) return B200S_OK;
    for (b200s_handle* h : p->h) b200s_destroy(h);
    delete p;
    return B200S_OK;
}

int b200s_pool_size(const b200s_pool* p) { return p ? (int)p->h.size() : 0; }
b200s_handle* b200s_pool_handle(b200s_pool* p, int gpu) { return (p && gpu >= 0 && gpu < (int)p->h.size()) ? p->h[gpu] : nullptr; }
const char* b200s_pool_last_error_string(const b200s_pool* p) { return p ? p->err.c_str() : "null pool"; }

extern "C++" {
namespace {
template <class F>
int pool_each(b200s_pool* p, F f)
{
    if (!p) return B200S_EINVAL;
    for (b200s_handle* h : p->h) {
        int rc = f(h);
        if (rc != B200S_OK) { p->err = h->err; return rc; }
    }
    return B200S_OK;
}
}  // namespace
}

int b200s_pool_set_calibration(b200s_pool* p, const b200s_caminfo* l, const b200s_caminfo* r)
{
    return pool_each(p, [&](b200s_handle* h) { return b200s_set_calibration(h, l, r); });
}

int b200s_pool_set_params(b200s_pool* p, const b200s_params* prm)
{
    return pool_each(p, [&](b200s_handle* h) { return b200s_set_params(h, prm); });
}

int b200s_pool_submit(b200s_pool* p, uint64_t frame_index, const void* left, const void* right, const b200s_frame_io* io, int* gpu, int* slot)
{
    if (!p || p->h.empty()) return B200S_EINVAL;
    const int n = (int)p->h.size();
    const int g = (int)(frame_index % (uint64_t)n), s = (int)((frame_index / (uint64_t)n) % (uint64_t)p->slots);
    if (gpu) *gpu = g;
    if (slot) *slot = s;
    b200s_handle* h = p->h[g];
    // the slot's previous frame (and the caller's output buffers for it) must be complete before it is reused
    int rc = b200s_wait_slot(h, s);
    if (rc == B200S_OK) rc = b200s_process_pair_async(h, s, left, right, io);
    if (rc != B200S_OK) p->err = h->err;
    return rc;
}

int b200s_pool_wait(b200s_pool* p, int gpu, int slot)
{
    b200s_handle* h = b200s_pool_handle(p, gpu);
    if (!h) return B200S_EINVAL;
    int rc = b200s_wait_slot(h, slot);
    if (rc != B200S_OK) p->err = h->err;
    return rc;
}

int b200s_pool_wait_all(b200s_pool* p)
{
    return pool_each(p, [&](b200s_handle* h) {
        for (int s = 0; s < (int)h->slots.size(); ++s) {
            int rc = b200s_wait_slot(h, s);
            if (rc != B200S_OK) return rc;
        }
        return (int)B200S_OK;
    });
}

// ---- batch timing: one start event all slot streams wait on, one end event per slot stream ----------------
int b200s_batch_begin(b200s_handle* h)
{
    if (!h || h->slots.empty()) return B200S_EINVAL;
    DeviceGuard g(h->device);
    if (!h->batch_start) CUDA_OK(h, cudaEventCreate(&h->batch_start));
    while (h->batch_end.size() < h->slots.size()) {
        cudaEvent_t e;
        CUDA_OK(h, cudaEventCreate(&e));
        h->batch_end.push_back(e);
    }
    CUDA_OK(h, cudaDeviceSynchronize());
    CUDA_OK(h, cudaEventRecord(h->batch_start, h->slots[0].st));
    for (size_t i = 1; i < h->slots.size(); ++i) CUDA_OK(h, cudaStreamWaitEvent(h->slots[i].st, h->batch_start, 0));
    return B200S_OK;
}

int b200s_batch_end(b200s_handle* h, float* ms)
{
    if (!h || !ms || !h->batch_start || h->batch_end.size() < h->slots.size()) return B200S_EINVAL;
    DeviceGuard g(h->device);
    for (size_t i = 0; i < h->slots.size(); ++i) CUDA_OK(h, cudaEventRecord(h->batch_end[i], h->slots[i].st));
    float best = 0;
    for (size_t i = 0; i < h->slots.size(); ++i) {
        CUDA_OK(h, cudaEventSynchronize(h->batch_end[i]));
        float t = 0;
        CUDA_OK(h, cudaEventElapsedTime(&t, h->batch_start, h->batch_end[i]));
        if (t > best) best = t;
    }
    *ms = best;
    return B200S_OK;
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
    return cudaHostAlloc(p, bytes, cudaHostAllocDefault) == cudaSuccess ? B200S_OK : B200S_ENOMEM;
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
