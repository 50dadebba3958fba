// State behind the opaque b200s_handle and the helpers shared by the translation units of the host layer:
//   api.cu    named-buffer API = the method set of gpuimageproc::GpuStereoProcessor (src/GPUStereoProcessor.cpp)
//   slots.cu  fused frame path on slots / batches (StereoProcessor::imageCb, src/StereoProcessor.cpp:157-298), CUDA-graph
//             replay, multi-GPU pool, timing
#pragma once
#include "../../include/b200_stereo.h"
#include "kernels.h"

#include <string>
#include <unordered_map>
#include <vector>

namespace b200s {

inline int elem_size(int type)
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

// stages of the frame chain that b200s_last_stage_times reports (CUDA events on the slot's stream)
enum { ST_RECTIFY = 0, ST_MATCH = 1, ST_POST = 2, ST_TOFLOAT = 3, ST_PACK = 4, ST_COUNT = 5 };

struct GraphEntry {            // one captured frame chain of a slot
    std::string key;
    cudaGraphExec_t exec = nullptr;
    uint64_t launches = 0, last_use = 0;
    double evals = 0;
};

// all device scratch one pipeline instance needs; a slot holds `depth` frames (one batch), frame f of every plane at
// base + f * stride
struct Work {
    cudaStream_t st = nullptr;
    bool own_stream = false;
    int depth = 1;
    DevBuf rawL, rawR, rawC, rectL, rectR, rectC, preL, preR, disp, cost, df, xyz, pc2, vol, ccl, normtmp, misc, lut;
    cudaEvent_t ev_bm0 = nullptr, ev_bm1 = nullptr, ev_done = nullptr;
    cudaEvent_t ev_stage[ST_COUNT + 1] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double last_evals = 0;
    bool timed = false, stages_timed = false;
    // CUDA graphs of the frame chain of this slot: the first frame with a given allocation signature runs eagerly
    // (allocations, map build), later ones are captured once per key and replayed with one cudaGraphLaunch
    std::vector<std::string> warm_keys;   // allocation signatures this slot has already run eagerly
    std::vector<GraphEntry> graphs;
    uint64_t use_counter = 0;
    std::string tab_cache;        // last input-address table written to the device (slots.cu)
    std::string border_key;       // geometry + FILTERED value the border of the slot's disparity planes was last filled for
    std::string lut_key;          // Q, cx difference and disparity range the pack kernel's per-disparity table was built for
    int lut_n = 0;                // entries of that table (0: none, the pack kernel computes per pixel)
    const void* in_tab[3 * MAX_BATCH] = {nullptr};   // host copy of that table: L[32], R[32], colour[32]
    void drop_graphs()
    {
        for (GraphEntry& g : graphs)
            if (g.exec) cudaGraphExecDestroy(g.exec);
        graphs.clear();
        warm_keys.clear();
    }
    void release()
    {
        drop_graphs();
        DevBuf* all[] = {&rawL, &rawR, &rawC, &rectL, &rectR, &rectC, &preL, &preR, &disp, &cost, &df, &xyz, &pc2, &vol, &ccl, &normtmp, &misc, &lut};
        lut_key.clear();
        lut_n = 0;
        for (DevBuf* b : all) b->release();
        if (ev_bm0) cudaEventDestroy(ev_bm0);
        if (ev_bm1) cudaEventDestroy(ev_bm1);
        if (ev_done) cudaEventDestroy(ev_done);
        for (cudaEvent_t& e : ev_stage) {
            if (e) cudaEventDestroy(e);
            e = nullptr;
        }
        if (own_stream && st) cudaStreamDestroy(st);
        ev_bm0 = ev_bm1 = ev_done = nullptr;
        st = nullptr;
    }
};

struct Camera {
    b200s_caminfo info;
    CamModel cm;
    DevBuf map;               // MAP_DELTA16: short2 per pixel, MAP_ABS32: int2 per pixel
    MapMode map_mode = MAP_DELTA16;
    bool map_valid = false;
};

// byte distance between consecutive frames of a slot's planes (multiples of 256)
struct SlotLayout {
    size_t raw = 0, rawc = 0, pre = 0, disp = 0, df = 0, xyz = 0, pc2 = 0, ccl = 0;
};

}  // namespace b200s

struct b200s_handle {
    int device = 0;
    std::string err;
    b200s_params prm;
    bool model_ok = false;
    bool rect_fly = false;
    bool timing = false;
    b200s::Camera cam[2];
    double Q[16];
    double baseline = 0, fx_right = 0, cxd = 0;
    unsigned qmask = 0xFFFFu;
    b200s::DevBuf Qdev, flagdev;
    std::unordered_map<int, b200s::Mat> mats;
    cudaStream_t l_strm = nullptr, r_strm = nullptr;
    cudaEvent_t ev_r = nullptr, ev_l = nullptr;
    b200s::Work w0;                  // scratch of the named-buffer API (runs on l_strm)
    std::vector<b200s::Work> slots;
    std::vector<cudaEvent_t> batch_end;
    cudaEvent_t batch_start = nullptr;
    int slot_rows = 0, slot_cols = 0, slot_depth = 1;
    b200s::SlotLayout lay;
    uint64_t launches = 0;
    void* stage_host[2] = {nullptr, nullptr};   // pinned staging for device->host copies into pageable user memory
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    uint64_t model_version = 0;   // bumped by every calibration change (part of the graph key)
    int use_graphs = 1;           // B200S_GRAPH=0 or b200s_set_graph_mode(h, 0) turns the replay off
    uint64_t graph_replays = 0;
    int pack_direct = 0;          // 1: pack kernels store straight into pinned (mapped) host destinations (measured slower than the copy engine)
    uint64_t stats_frames = 0;    // printStats: frames through the fused path / named disparity calls
};

namespace b200s {

int fail(b200s_handle* h, int code, const std::string& msg);

#define CUDA_OK(h, call)                                                                        \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return b200s::fail(h, B200S_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

// makes the handle's GPU current for the duration of an API call and restores the caller's device afterwards
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int d)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != d) cudaSetDevice(d);
        else prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int check_kernels(b200s_handle* h, const char* what);
int validate_params(b200s_handle* h, const b200s_params& p);
BMConfig bm_config(const b200s_params& p);
int ensure_map(b200s_handle* h, int side /*0 L, 1 R*/, cudaStream_t st);
// rectification table of a side as the kernels want it (mode MAP_FLY / nullptr when maps are evaluated on the fly)
int map_for(b200s_handle* h, int side, cudaStream_t st, MapMode* mode, const void** map);
int ensure_pre_planes(b200s_handle* h, Work& w, int rows, int cols, int nf);
int ensure_misc(b200s_handle* h, Work& w);
// prefilter + match + post-filters on rectified device planes (a batch of nf frames: sources src_stride bytes apart,
// prefiltered planes plane_stride(cols, rows) apart, disparity planes disp_stride apart)
int run_disparity(b200s_handle* h, Work& w, const uint8_t* L, const uint8_t* R, bool already_prefiltered, int rows, int cols,
                  int16_t* disp, cudaStream_t st, int nf = 1, size_t src_stride = 0, size_t disp_stride = 0,
                  const uint8_t* const* tabL = nullptr, const uint8_t* const* tabR = nullptr, bool keep_border = false);
constexpr size_t MISC_BYTES = 2048;   // per-frame minima (ints) + the per-frame input address table of the slot
inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
inline size_t plane_stride(int cols, int rows) { return align256(plane_bytes(cols, rows)); }
// Device-side alias of a pinned (page-locked, mapped) host buffer, or nullptr when `p` is pageable / not host memory
void* mapped_alias(const void* p);
int d2h_sync(b200s_handle* h, void* dst, const void* src, size_t bytes, cudaStream_t st);
void fill_disparity_meta(const b200s_handle* h, int rows, int cols, b200s_disparity_meta* m);
void fill_pc2_meta(int rows, int cols, b200s_pc2_meta* m);

}  // namespace b200s
