// Fused frame path of libb200stereo.so: the whole StereoProcessor::imageCb chain (reference src/StereoProcessor.cpp:157-298)
// on frame slots.  A slot = one stream + device buffers for a BATCH of up to `frames_per_slot` frames; every kernel of the
// chain takes the batch in one launch (frame index = a grid dimension), which keeps all SMs busy on the small BASELINE
// shapes and lets the matcher use tall row bands.  The chain of a slot is captured into CUDA graphs (one per parameter /
// product / destination set, a few cached per slot) and replayed with one cudaGraphLaunch.
#include "handle.h"

#include <cstdlib>
#include <cstring>

using namespace b200s;

namespace {

int copy_out(b200s_handle* h, void* dst, const void* src, size_t bytes, bool dst_on_device, cudaStream_t st)
{
    if (!dst) return B200S_OK;
    CUDA_OK(h, cudaMemcpyAsync(dst, src, bytes, dst_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    return B200S_OK;
}

constexpr int MAX_GRAPHS_PER_SLOT = 4;

void stage_mark(b200s_handle* h, Work& w, int i, cudaStream_t st)
{
    if (!h->timing) return;
    if (!w.ev_stage[i]) cudaEventCreate(&w.ev_stage[i]);
    cudaEventRecord(w.ev_stage[i], st);
}

// Per-frame input addresses live in a small device table inside w.misc (after the per-frame minima): the kernels of the
// chain read tab[frame], so the caller's own device images are used in place and a captured graph does not depend on
// where the inputs are -- only the table is rewritten (one small H2D copy) when the addresses change.
constexpr size_t MISC_TAB_OFFSET = 256;      // after the per-frame minima; ensure_misc() sizes the buffer (MISC_BYTES, handle.h)
static_assert(MISC_TAB_OFFSET + 3 * MAX_BATCH * sizeof(void*) <= MISC_BYTES, "misc buffer too small for the address table");
const uint8_t* const* tab_of(const Work& w, int which /*0 L, 1 R, 2 colour*/)
{
    return (const uint8_t* const*)((const uint8_t*)w.misc.p + MISC_TAB_OFFSET) + which * MAX_BATCH;
}

// everything of a batch after the input table is set: rectify -> disparity -> float / reproject+pack -> outputs
int run_frame_chain(b200s_handle* h, Work& w, int nf, const b200s_frame_io* ios, bool have_color, cudaStream_t st)
{
    const b200s_frame_io* io = &ios[0];
    const int rows = h->slot_rows, cols = h->slot_cols;
    const size_t n = (size_t)rows * cols;
    const SlotLayout& lay = h->lay;
    const size_t pstride = plane_stride(cols, rows);
    bool prefiltered = false;
    const uint8_t* const* tabL = tab_of(w, 0);
    const uint8_t* const* tabR = tab_of(w, 1);
    const uint8_t* const* tabC = have_color ? tab_of(w, 2) : nullptr;
    const uint8_t *rl = nullptr, *rr = nullptr, *rc_color = nullptr;
    size_t rect_stride = 0, rcol_stride = 0;
    stage_mark(h, w, 0, st);
    if (io->rectify) {
        if (w.rectL.ensure(lay.raw * w.depth) || w.rectR.ensure(lay.raw * w.depth)) return fail(h, B200S_ENOMEM, "cudaMalloc failed (rectified planes)");
        MapMode mode = MAP_FLY, modeR = MAP_FLY;
        const void *mapL = nullptr, *mapR = nullptr;
        int rc = map_for(h, 0, st, &mode, &mapL);
        if (rc) return rc;
        rc = map_for(h, 1, st, &modeR, &mapR);
        if (rc) return rc;
        const bool same_mode = mode == modeR;
        int rc2 = ensure_pre_planes(h, w, rows, cols, w.depth);
        if (rc2) return rc2;
        uint8_t* pl = (uint8_t*)w.preL.p + PLANE_LEAD;
        uint8_t* pr = (uint8_t*)w.preR.p + PLANE_LEAD;
        if (same_mode && h->prm.pre_filter_type == 1) {
            h->launches += launch_rectify_xsobel_pair(nullptr, nullptr, cols, rows, mode, mapL, mapR, h->cam[0].cm, h->cam[1].cm,
                                                      (uint8_t*)w.rectL.p, (uint8_t*)w.rectR.p, pl, pr, plane_pitch(cols), cols, rows,
                                                      h->prm.pre_filter_cap, st, nf, 0, lay.raw, pstride, tabL, tabR);
            prefiltered = true;
        } else {
            int one = 0;
            if (same_mode && h->prm.pre_filter_type == 0 && h->prm.pre_filter_size <= 21)
                one = launch_norm_prefilter_pair(nullptr, nullptr, cols, rows, mode, mapL, mapR, h->cam[0].cm, h->cam[1].cm,
                                                 (uint8_t*)w.rectL.p, (uint8_t*)w.rectR.p, pl, pr, plane_pitch(cols), cols, rows,
                                                 h->prm.pre_filter_size, h->prm.pre_filter_cap, st, nf, 0, lay.raw, pstride, tabL, tabR);
            if (one) {
                h->launches += one;
                prefiltered = true;
            } else {
                h->launches += launch_remap(nullptr, cols, rows, 1, mapL, mode, h->cam[0].cm, (uint8_t*)w.rectL.p, cols, rows, st, nf, 0, lay.raw, tabL);
                h->launches += launch_remap(nullptr, cols, rows, 1, mapR, modeR, h->cam[1].cm, (uint8_t*)w.rectR.p, cols, rows, st, nf, 0, lay.raw, tabR);
            }
        }
        rl = (const uint8_t*)w.rectL.p;
        rr = (const uint8_t*)w.rectR.p;
        rect_stride = lay.raw;
        if (have_color) {
            // the colour image feeds the point cloud (src/StereoProcessor.cpp:201-217: L_RECT_COLOR -> enqueueSendPoints)
            if (w.rectC.ensure(lay.rawc * w.depth)) return fail(h, B200S_ENOMEM, "cudaMalloc failed (rectified colour plane)");
            h->launches += launch_remap(nullptr, cols, rows, 3, mapL, mode, h->cam[0].cm, (uint8_t*)w.rectC.p, cols, rows, st, nf, 0, lay.rawc, tabC);
            rc_color = (const uint8_t*)w.rectC.p;
            rcol_stride = lay.rawc;
        }
    }
    if (w.disp.ensure(lay.disp * w.depth)) return fail(h, B200S_ENOMEM, "cudaMalloc failed (disparity plane)");
    int rc = run_disparity(h, w, rl, rr, prefiltered, rows, cols, (int16_t*)w.disp.p, st, nf, rect_stride, lay.disp, rl ? nullptr : tabL,
                           rr ? nullptr : tabR, true);
    if (rc) return rc;
    h->stats_frames += nf - 1;
    stage_mark(h, w, 3, st);
    const bool want_pc = io->want & B200S_OUT_POINTCLOUD2, want_xyz = io->want & B200S_OUT_POINTS_XYZ;
    const bool want_df = io->want & B200S_OUT_DISPARITY32F;
    const bool od = io->outputs_on_device != 0;
    // "direct" pack mode (the north star's wording): float disparity and PointCloud2 records are stored by the kernels
    // straight into the caller's pinned host buffers (posted PCIe writes) -- no HBM copy of the payload, no copy-engine
    // transfer afterwards.  Needs every frame's destination to be page-locked host memory.
    PtrList df_list, pc_list;
    bool df_direct = h->pack_direct && want_df && !od, pc_direct = h->pack_direct && want_pc && !od;
    for (int f = 0; f < nf; ++f) {
        df_list.p[f] = df_direct ? mapped_alias(ios[f].disparity32f) : nullptr;
        pc_list.p[f] = pc_direct ? mapped_alias(ios[f].pointcloud2) : nullptr;
        df_direct = df_direct && df_list.p[f];
        pc_direct = pc_direct && pc_list.p[f];
    }
    for (int f = nf; f < MAX_BATCH; ++f) df_list.p[f] = pc_list.p[f] = nullptr;
    // The float disparity plane (DisparityImage payload) comes out of the reproject + pack pass over d16; the missing value
    // of cv::reprojectImageTo3D (= the minimum of the plane) is FILTERED for every plane this chain produces, because the
    // border columns always hold it -- no reduction pass.  Only a lone float plane still uses the conversion kernel.
    const int filtered = (h->prm.min_disparity - 1) * 16;
    if (want_df && !df_direct && w.df.ensure(lay.df * w.depth)) return fail(h, B200S_ENOMEM, "cudaMalloc failed (float disparity)");
    if (want_df && !(want_pc || want_xyz))
        h->launches += launch_disparity_to_float((const int16_t*)w.disp.p, !df_direct ? (float*)w.df.p : nullptr, (int)n,
                                                 h->model_ok ? h->cxd : 0.0, (int*)w.misc.p, st, nf, lay.disp, lay.df,
                                                 df_direct ? &df_list : nullptr);
    stage_mark(h, w, 4, st);
    if (want_pc || want_xyz) {
        if (want_pc && !pc_direct && w.pc2.ensure(lay.pc2 * w.depth)) return fail(h, B200S_ENOMEM, "cudaMalloc failed (point cloud)");
        if (want_xyz && w.xyz.ensure(lay.xyz * w.depth)) return fail(h, B200S_ENOMEM, "cudaMalloc failed (points)");
        ReprojectExtras ex;
        ex.dmin_const = filtered;
        if (w.lut_n > 0 && !want_xyz) {
            ex.lut = w.lut.p;
            ex.lut_n = w.lut_n;
        }
        if (want_df) {
            ex.df = df_direct ? nullptr : (float*)w.df.p;
            ex.df_stride = lay.df;
            ex.df_list = df_direct ? &df_list : nullptr;
        }
        // colour of the records: the rectified colour image, else the rectified grey image, else (rectify = 0) the inputs
        const uint8_t* col = have_color ? rc_color : rl;
        size_t col_stride = have_color ? rcol_stride : rect_stride;
        if (!col) ex.color_tab = have_color ? tabC : tabL;
        h->launches += launch_reproject_pack((const int16_t*)w.disp.p, cols, rows, h->cxd, (const double*)h->Qdev.p, h->qmask,
                                             nullptr, col, have_color ? 3 : 1, want_xyz ? (float*)w.xyz.p : nullptr,
                                             want_pc && !pc_direct ? (uint8_t*)w.pc2.p : nullptr, st, nf, lay.disp, col_stride, lay.xyz,
                                             lay.pc2, pc_direct ? &pc_list : nullptr, &ex);
    }
    stage_mark(h, w, 5, st);
    rc = ensure_misc(h, w);
    if (rc) return rc;
    rc = check_kernels(h, "process_pair");
    if (rc) return rc;
    for (int f = 0; f < nf; ++f) {
        const b200s_frame_io& o = ios[f];
        if (io->rectify) {
            if ((io->want & B200S_OUT_RECT_L) && (rc = copy_out(h, o.rect_left, (uint8_t*)w.rectL.p + f * lay.raw, n, od, st))) return rc;
            if ((io->want & B200S_OUT_RECT_R) && (rc = copy_out(h, o.rect_right, (uint8_t*)w.rectR.p + f * lay.raw, n, od, st))) return rc;
            if ((io->want & B200S_OUT_RECT_COLOR_L) && have_color &&
                (rc = copy_out(h, o.rect_color_left, (uint8_t*)w.rectC.p + f * lay.rawc, n * 3, od, st))) return rc;
        }
        if ((io->want & B200S_OUT_DISPARITY16) && (rc = copy_out(h, o.disparity16, (uint8_t*)w.disp.p + f * lay.disp, n * 2, od, st))) return rc;
        if (want_df && !df_direct && (rc = copy_out(h, o.disparity32f, (uint8_t*)w.df.p + f * lay.df, n * 4, od, st))) return rc;
        if (want_pc && !pc_direct && (rc = copy_out(h, o.pointcloud2, (uint8_t*)w.pc2.p + f * lay.pc2, n * 32, od, st))) return rc;
        if (want_xyz && (rc = copy_out(h, o.points_xyz, (uint8_t*)w.xyz.p + f * lay.xyz, n * 12, od, st))) return rc;
    }
    return B200S_OK;
}

// The pack kernel's per-disparity table (kernels.h launch_reproject_lut) of this slot: rebuilt on the slot's stream, outside
// any graph, when the calibration or the disparity range changed.  The buffer keeps its address (graphs hold it).
static int ensure_reproject_lut(b200s_handle* h, Work& w, cudaStream_t st)
{
    static const int use_lut = getenv("B200S_PACK_LUT") ? atoi(getenv("B200S_PACK_LUT")) : 1;
    const int dmin = (h->prm.min_disparity - 1) * 16;
    int n = (h->prm.num_disparities + 1) * 16 + 1;             // FILTERED .. the largest value the sub-pixel fit can round to
    // test hook: a table that is too short, so that most disparities take the kernel's arithmetic for values outside it
    static const int max_entries = getenv("B200S_PACK_LUT_ENTRIES") ? atoi(getenv("B200S_PACK_LUT_ENTRIES")) : 0;
    if (max_entries > 0 && n > max_entries) n = max_entries;
    if (!use_lut || !h->model_ok || n > REPROJECT_LUT_MAX) {
        w.lut_n = 0;
        w.lut_key.clear();
        return B200S_OK;
    }
    std::string key((const char*)h->Q, sizeof h->Q);
    key.append((const char*)&h->cxd, sizeof h->cxd);
    key.append((const char*)&dmin, sizeof dmin);
    key.append((const char*)&n, sizeof n);
    if (key == w.lut_key) return B200S_OK;
    if (w.lut.ensure((size_t)REPROJECT_LUT_MAX * 16)) return fail(h, B200S_ENOMEM, "cudaMalloc failed (reprojection table)");
    const int built = launch_reproject_lut(w.lut.p, n, dmin, h->cxd, (const double*)h->Qdev.p, h->qmask, st);
    h->launches += built;
    w.lut_n = built ? n : 0;
    w.lut_key = key;
    return B200S_OK;
}

// everything a captured chain depends on besides the (fixed) slot buffers; with_ptrs = false gives the allocation
// signature (what decides which buffers exist), used to tell whether the slot has been warmed up for this kind of frame
std::string frame_graph_key(const b200s_handle* h, int nf, const b200s_frame_io* ios, bool with_color, bool with_ptrs)
{
    std::string k;
    auto add = [&k](const void* p, size_t n) { k.append((const char*)p, n); };
    add(&h->prm, sizeof h->prm);
    add(&h->model_version, sizeof h->model_version);
    add(&h->slot_rows, sizeof h->slot_rows);
    add(&h->slot_cols, sizeof h->slot_cols);
    add(&h->pack_direct, sizeof h->pack_direct);
    add(&nf, sizeof nf);
    add(&with_color, sizeof with_color);
    for (int f = 0; f < nf; ++f) {
        b200s_frame_io o = ios[f];
        o.color_left = nullptr;                      // inputs are copied into the slot outside the graph
        if (!with_ptrs) {
            const bool host = !o.outputs_on_device;
            // pinned-ness of the destinations decides between direct stores and staging copies
            const int pin = host && h->pack_direct ? ((mapped_alias(o.disparity32f) ? 1 : 0) | (mapped_alias(o.pointcloud2) ? 2 : 0)) : 0;
            o.rect_left = o.rect_right = o.disparity16 = o.disparity32f = o.pointcloud2 = o.points_xyz = o.rect_color_left = nullptr;
            add(&pin, sizeof pin);
        }
        add(&o, sizeof o);
    }
    return k;
}

}  // namespace

// ==========================================================================================================
extern "C" {

int b200s_configure_slots_batched(b200s_handle* h, int n_slots, int rows, int cols, int frames_per_slot)
{
    if (!h || n_slots < 1 || n_slots > 64 || rows <= 0 || cols <= 0 || frames_per_slot < 1 || frames_per_slot > MAX_BATCH) return B200S_EINVAL;
    DeviceGuard g(h->device);
    cudaDeviceSynchronize();
    for (Work& w : h->slots) w.release();
    h->slots.clear();
    h->slots.resize(n_slots);
    for (Work& w : h->slots) {
        if (cudaStreamCreateWithFlags(&w.st, cudaStreamNonBlocking) != cudaSuccess) return fail(h, B200S_ECUDA, "cudaStreamCreate failed");
        w.own_stream = true;
        w.depth = frames_per_slot;
        cudaEventCreateWithFlags(&w.ev_done, cudaEventDisableTiming);
        cudaEventCreate(&w.ev_bm0);
        cudaEventCreate(&w.ev_bm1);
    }
    h->slot_rows = rows;
    h->slot_cols = cols;
    h->slot_depth = frames_per_slot;
    const size_t n = (size_t)rows * cols;
    h->lay.raw = align256(n + 64);
    h->lay.rawc = align256(3 * n + 64);
    h->lay.pre = plane_stride(cols, rows);
    h->lay.disp = align256(2 * n + 64);
    h->lay.df = align256(4 * n);
    h->lay.xyz = align256(12 * n);
    h->lay.pc2 = align256(32 * n);
    return B200S_OK;
}

int b200s_configure_slots(b200s_handle* h, int n_slots, int rows, int cols)
{
    return b200s_configure_slots_batched(h, n_slots, rows, cols, 1);
}

int b200s_process_batch_async(b200s_handle* h, int slot, int n_frames, const void* const* left, const void* const* right,
                              const b200s_frame_io* ios)
{
    if (!h || !left || !right || !ios || n_frames < 1) return B200S_EINVAL;
    DeviceGuard g(h->device);
    if (slot < 0 || slot >= (int)h->slots.size()) return fail(h, B200S_EINVAL, "slot out of range (call b200s_configure_slots)");
    Work& w = h->slots[slot];
    if (n_frames > w.depth) return fail(h, B200S_EINVAL, "batch larger than the slot's frames_per_slot (b200s_configure_slots_batched)");
    const int nf = n_frames;
    const int rows = h->slot_rows, cols = h->slot_cols;
    const size_t n = (size_t)rows * cols;
    const b200s_frame_io* io = &ios[0];
    bool with_color = false;
    for (int f = 0; f < nf; ++f) {
        const b200s_frame_io& o = ios[f];
        if (o.want != io->want || o.rectify != io->rectify || o.inputs_on_device != io->inputs_on_device ||
            o.outputs_on_device != io->outputs_on_device || o.color_encoding != io->color_encoding || !o.color_left != !io->color_left)
            return fail(h, B200S_EINVAL, "all frames of a batch must ask for the same products, flags and colour encoding");
        if ((o.rows && o.rows != rows) || (o.cols && o.cols != cols))
            return fail(h, B200S_EINVAL, "frame size differs from the configured slot size (call b200s_configure_slots again)");
        if (!right[f] || (!left[f] && !o.color_left)) return B200S_EINVAL;
        if (o.color_left && o.color_encoding != B200S_COLOR_BGR8 && o.color_encoding != B200S_COLOR_RGB8)
            return fail(h, B200S_EUNSUPPORTED, "colour input must be bgr8 or rgb8");
    }
    with_color = io->color_left != nullptr;
    cudaStream_t st = w.st;
    const bool need_model = io->rectify || (io->want & (B200S_OUT_POINTCLOUD2 | B200S_OUT_POINTS_XYZ));
    if (need_model && !h->model_ok) return fail(h, B200S_ENOTINIT, "stereo model not initialised");
    if (io->rectify && (h->cam[0].info.width != cols || h->cam[0].info.height != rows))
        return fail(h, B200S_EINVAL, "slot size differs from the calibration size");
    const bool graphs = h->use_graphs && !h->timing;
    // Inputs.  Host frames are copied into the slot's raw planes; device frames are read where they are.  Either way the
    // kernels get the addresses from the slot's device table, which is rewritten only when it changes.
    const SlotLayout& lay = h->lay;
    {
        int rcm = ensure_misc(h, w);      // one fixed size for the life of the slot: the table must never move
        if (rcm) return rcm;
    }
    const void* tab[3 * MAX_BATCH];
    memset(tab, 0, sizeof tab);
    const bool host_in = !io->inputs_on_device;
    if (host_in || with_color) {
        if ((host_in || !left[0]) && (w.rawL.ensure(lay.raw * w.depth))) return fail(h, B200S_ENOMEM, "cudaMalloc failed (input planes)");
        if (host_in && w.rawR.ensure(lay.raw * w.depth)) return fail(h, B200S_ENOMEM, "cudaMalloc failed (input planes)");
        if (with_color && (host_in || io->color_encoding == B200S_COLOR_RGB8) && w.rawC.ensure(lay.rawc * w.depth))
            return fail(h, B200S_ENOMEM, "cudaMalloc failed (colour input planes)");
    }
    for (int f = 0; f < nf; ++f) {
        const void *l = left[f], *r = right[f], *c = with_color ? ios[f].color_left : nullptr;
        if (host_in) {
            if (l) CUDA_OK(h, cudaMemcpyAsync((uint8_t*)w.rawL.p + f * lay.raw, l, n, cudaMemcpyHostToDevice, st));
            CUDA_OK(h, cudaMemcpyAsync((uint8_t*)w.rawR.p + f * lay.raw, r, n, cudaMemcpyHostToDevice, st));
            if (c) CUDA_OK(h, cudaMemcpyAsync((uint8_t*)w.rawC.p + f * lay.rawc, c, 3 * n, cudaMemcpyHostToDevice, st));
            if (l) l = (uint8_t*)w.rawL.p + f * lay.raw;
            r = (uint8_t*)w.rawR.p + f * lay.raw;
            if (c) c = (uint8_t*)w.rawC.p + f * lay.rawc;
        }
        if (c && io->color_encoding == B200S_COLOR_RGB8) {
            // convertRawToColor of a colour camera (src/GPUStereoProcessor.cpp:65-88): the slot keeps BGR
            uint8_t* cf = (uint8_t*)w.rawC.p + f * lay.rawc;
            h->launches += launch_swap_rb((const uint8_t*)c, cf, (int)n, st);
            c = cf;
        }
        if (c && !l) {
            // convertRawToMono: the matcher's grey image comes from the colour image when no mono image was given
            uint8_t* lf = (uint8_t*)w.rawL.p + f * lay.raw;
            h->launches += launch_bgr_to_gray((const uint8_t*)c, lf, (int)n, 0, st);
            l = lf;
        }
        tab[f] = l;
        tab[MAX_BATCH + f] = r;
        tab[2 * MAX_BATCH + f] = c;
    }
    if (w.tab_cache.size() != sizeof tab || memcmp(w.tab_cache.data(), tab, sizeof tab) != 0) {
        // pageable source: the driver stages small host->device copies before cudaMemcpyAsync returns
        CUDA_OK(h, cudaMemcpyAsync((uint8_t*)w.misc.p + MISC_TAB_OFFSET, tab, sizeof tab, cudaMemcpyHostToDevice, st));
        w.tab_cache.assign((const char*)tab, sizeof tab);
    }
    memcpy(w.in_tab, tab, sizeof tab);
    int rc = ensure_reproject_lut(h, w, st);
    if (rc) return rc;
    if (!graphs) {
        rc = run_frame_chain(h, w, nf, ios, with_color, st);
    } else {
        const std::string key = frame_graph_key(h, nf, ios, with_color, true);
        GraphEntry* hit = nullptr;
        for (GraphEntry& ge : w.graphs)
            if (ge.key == key) hit = &ge;
        if (hit) {
            CUDA_OK(h, cudaGraphLaunch(hit->exec, st));
            hit->last_use = ++w.use_counter;
            h->launches += hit->launches;
            h->stats_frames += nf;
            w.last_evals = hit->evals;
            ++h->graph_replays;
        } else {
            const std::string akey = frame_graph_key(h, nf, ios, with_color, false);
            bool warm = false;
            for (const std::string& k2 : w.warm_keys) warm = warm || k2 == akey;
            if (!warm) {
                // first frame of this kind on the slot: run eagerly (allocations, map build).  Slot buffers have fixed
                // sizes and are never reallocated, so graphs captured earlier stay valid.
                rc = run_frame_chain(h, w, nf, ios, with_color, st);
                if (rc == B200S_OK) {
                    if (w.warm_keys.size() >= 8) w.warm_keys.erase(w.warm_keys.begin());
                    w.warm_keys.push_back(akey);
                }
            } else {
                // every buffer exists, the maps are built -> capture, instantiate, launch
                const uint64_t l0 = h->launches, f0 = h->stats_frames;
                cudaGraph_t graph = nullptr;
                cudaGraphExec_t exec = nullptr;
                cudaError_t e = cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed);
                if (e == cudaSuccess) {
                    rc = run_frame_chain(h, w, nf, ios, with_color, st);
                    e = cudaStreamEndCapture(st, &graph);
                    if (rc == B200S_OK && e == cudaSuccess && graph) e = cudaGraphInstantiate(&exec, graph, 0);
                    if (graph) cudaGraphDestroy(graph);
                }
                if (rc != B200S_OK || e != cudaSuccess || !exec) {
                    // capture is an optimisation only: fall back to eager launches for this handle
                    cudaGetLastError();
                    if (exec) cudaGraphExecDestroy(exec);
                    w.drop_graphs();
                    h->use_graphs = 0;
                    h->launches = l0;
                    h->stats_frames = f0;
                    rc = run_frame_chain(h, w, nf, ios, with_color, st);
                } else {
                    if ((int)w.graphs.size() >= MAX_GRAPHS_PER_SLOT) {       // evict the least recently used graph
                        size_t lru = 0;
                        for (size_t i = 1; i < w.graphs.size(); ++i)
                            if (w.graphs[i].last_use < w.graphs[lru].last_use) lru = i;
                        cudaGraphExecDestroy(w.graphs[lru].exec);
                        w.graphs.erase(w.graphs.begin() + lru);
                    }
                    GraphEntry ge;
                    ge.key = key;
                    ge.exec = exec;
                    ge.launches = h->launches - l0;
                    ge.evals = w.last_evals;
                    ge.last_use = ++w.use_counter;
                    w.graphs.push_back(ge);
                    CUDA_OK(h, cudaGraphLaunch(exec, st));
                }
            }
        }
    }
    if (rc) return rc;
    CUDA_OK(h, cudaEventRecord(w.ev_done, st));
    return B200S_OK;
}

int b200s_process_pair_async(b200s_handle* h, int slot, const void* left, const void* right, const b200s_frame_io* io)
{
    if (!io) return B200S_EINVAL;
    const void* l[1] = {left};
    const void* r[1] = {right};
    return b200s_process_batch_async(h, slot, 1, l, r, io);
}

int b200s_set_graph_mode(b200s_handle* h, int on)
{
    if (!h) return B200S_EINVAL;
    h->use_graphs = on ? 1 : 0;
    for (Work& w : h->slots) w.drop_graphs();
    return B200S_OK;
}

uint64_t b200s_graph_replays(const b200s_handle* h) { return h ? h->graph_replays : 0; }

int b200s_set_pack_mode(b200s_handle* h, int direct)
{
    if (!h) return B200S_EINVAL;
    h->pack_direct = direct ? 1 : 0;
    return B200S_OK;
}

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

int b200s_slot_frame_device_ptr(b200s_handle* h, int slot, int frame, uint32_t which, void** dptr, size_t* bytes)
{
    if (!h || !dptr) return B200S_EINVAL;
    if (slot < 0 || slot >= (int)h->slots.size()) return fail(h, B200S_EINVAL, "slot out of range");
    Work& w = h->slots[slot];
    if (frame < 0 || frame >= w.depth) return fail(h, B200S_EINVAL, "frame out of range");
    size_t n = (size_t)h->slot_rows * h->slot_cols;
    const SlotLayout& lay = h->lay;
    DevBuf* b = nullptr;
    size_t sz = 0, stride = 0;
    switch (which) {
        case B200S_OUT_RECT_L: b = &w.rectL; sz = n; stride = lay.raw; break;
        case B200S_OUT_RECT_R: b = &w.rectR; sz = n; stride = lay.raw; break;
        case B200S_OUT_RECT_COLOR_L: b = &w.rectC; sz = n * 3; stride = lay.rawc; break;
        case B200S_OUT_DISPARITY16: b = &w.disp; sz = n * 2; stride = lay.disp; break;
        case B200S_OUT_DISPARITY32F: b = &w.df; sz = n * 4; stride = lay.df; break;
        case B200S_OUT_POINTCLOUD2: b = &w.pc2; sz = n * 32; stride = lay.pc2; break;
        case B200S_OUT_POINTS_XYZ: b = &w.xyz; sz = n * 12; stride = lay.xyz; break;
        default: return fail(h, B200S_EINVAL, "unknown product");
    }
    if (!b->p || b->cap < stride * (size_t)frame + sz) return fail(h, B200S_ENOBUF, "product has not been computed on this slot yet");
    *dptr = (uint8_t*)b->p + stride * (size_t)frame;
    if (bytes) *bytes = sz;
    return B200S_OK;
}

int b200s_slot_device_ptr(b200s_handle* h, int slot, uint32_t which, void** dptr, size_t* bytes)
{
    return b200s_slot_frame_device_ptr(h, slot, 0, which, dptr, bytes);
}

int b200s_process_pair(b200s_handle* h, const void* left, const void* right, const b200s_frame_io* io)
{
    if (!h) return B200S_EINVAL;
    if (h->slots.empty()) return fail(h, B200S_EINVAL, "call b200s_configure_slots first");
    int rc = b200s_process_pair_async(h, 0, left, right, io);
    if (rc) return rc;
    return b200s_wait_slot(h, 0);
}

// device time of the stages of the slot's last frame chain (b200s_enable_timing must be on; graphs are bypassed then)
int b200s_last_stage_times(b200s_handle* h, int slot, float* ms /* [B200S_STAGE_COUNT] */)
{
    if (!h || !ms) return B200S_EINVAL;
    if (slot < 0 || slot >= (int)h->slots.size()) return fail(h, B200S_EINVAL, "slot out of range");
    Work& w = h->slots[slot];
    DeviceGuard g(h->device);
    if (!w.ev_stage[0] || !w.ev_stage[3] || !w.ev_stage[4] || !w.ev_stage[5] || !w.timed)
        return fail(h, B200S_ENOBUF, "no timed frame on this slot (b200s_enable_timing)");
    CUDA_OK(h, cudaEventSynchronize(w.ev_stage[5]));
    for (int i = 0; i < ST_COUNT; ++i) ms[i] = 0.f;
    CUDA_OK(h, cudaEventElapsedTime(&ms[ST_RECTIFY], w.ev_stage[0], w.ev_bm0));
    CUDA_OK(h, cudaEventElapsedTime(&ms[ST_MATCH], w.ev_bm0, w.ev_bm1));
    CUDA_OK(h, cudaEventElapsedTime(&ms[ST_POST], w.ev_bm1, w.ev_stage[3]));
    CUDA_OK(h, cudaEventElapsedTime(&ms[ST_TOFLOAT], w.ev_stage[3], w.ev_stage[4]));
    CUDA_OK(h, cudaEventElapsedTime(&ms[ST_PACK], w.ev_stage[4], w.ev_stage[5]));
    return B200S_OK;
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

// ---- bare copy probe: what the host side of PCIe absorbs with no kernels at all -----------------------------------
int b200s_copy_probe_ex(int device, size_t bytes, double seconds, int host_mode, int with_h2d, int n_streams, int n_buffers,
                        double* d2h_gbs)
{
    if (!d2h_gbs || bytes == 0 || seconds <= 0 || n_streams < 1 || n_streams > 8 || n_buffers < 1 || n_buffers > 16) return B200S_EINVAL;
    DeviceGuard g(device);
    void* dev = nullptr;
    void* din = nullptr;
    void* hin = nullptr;
    void* host[16] = {nullptr};
    cudaStream_t st[9] = {nullptr};        // n_streams copy streams + one for the host->device side
    cudaEvent_t e0 = nullptr, e1[8] = {nullptr};
    const size_t in_bytes = 4u << 20;
    int rc = B200S_OK;
    auto cleanup = [&]() {
        for (void* p : host)
            if (p) {
                if (host_mode == 2) { cudaHostUnregister(p); free(p); }
                else cudaFreeHost(p);
            }
        if (hin) cudaFreeHost(hin);
        if (dev) cudaFree(dev);
        if (din) cudaFree(din);
        for (cudaStream_t s : st) if (s) cudaStreamDestroy(s);
        if (e0) cudaEventDestroy(e0);
        for (cudaEvent_t e : e1) if (e) cudaEventDestroy(e);
    };
    bool ok = cudaMalloc(&dev, bytes) == cudaSuccess && cudaMalloc(&din, in_bytes) == cudaSuccess &&
              cudaHostAlloc(&hin, in_bytes, cudaHostAllocDefault) == cudaSuccess && cudaMemset(dev, 7, bytes) == cudaSuccess;
    for (int i = 0; ok && i < n_buffers; ++i) {
        if (host_mode == 2) {
            host[i] = aligned_alloc(4096, (bytes + 4095) & ~(size_t)4095);
            ok = host[i] && cudaHostRegister(host[i], bytes, cudaHostRegisterDefault) == cudaSuccess;
        } else {
            ok = cudaHostAlloc(&host[i], bytes, host_mode == 1 ? cudaHostAllocWriteCombined : cudaHostAllocDefault) == cudaSuccess;
        }
        if (ok) memset(host[i], 0, bytes);      // touch the pages (placement follows the caller's memory policy)
    }
    for (int i = 0; ok && i <= n_streams; ++i) ok = cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreate(&e0) == cudaSuccess;
    for (int i = 0; ok && i < n_streams; ++i) ok = cudaEventCreate(&e1[i]) == cudaSuccess;
    if (!ok) { cleanup(); cudaGetLastError(); return B200S_ECUDA; }
    for (int i = 0; i < n_buffers; ++i) cudaMemcpyAsync(host[i], dev, bytes, cudaMemcpyDeviceToHost, st[i % n_streams]);
    cudaDeviceSynchronize();
    // number of rounds from a calibration pass, so that the timed pass is pure copies of about the requested duration
    const int per_round = 2 * n_streams;         // two copies queued per stream
    long rounds = (long)(0.2 * seconds * 20e9 / ((double)per_round * bytes)) + 1;
    long copies = 0;
    float ms = 0;
    for (int pass = 0; pass < 2; ++pass) {
        cudaDeviceSynchronize();
        cudaEventRecord(e0, st[0]);
        for (int i = 1; i < n_streams; ++i) cudaStreamWaitEvent(st[i], e0, 0);
        copies = 0;
        for (long r = 0; r < rounds; ++r) {
            for (int i = 0; i < per_round; ++i) {
                cudaMemcpyAsync(host[copies % n_buffers], dev, bytes, cudaMemcpyDeviceToHost, st[i % n_streams]);
                if (with_h2d) cudaMemcpyAsync(din, hin, in_bytes, cudaMemcpyHostToDevice, st[n_streams]);
                ++copies;
            }
            cudaStreamSynchronize(st[0]);       // at most two rounds queued ahead
        }
        float worst = 0;
        for (int i = 0; i < n_streams; ++i) cudaEventRecord(e1[i], st[i]);
        cudaDeviceSynchronize();
        for (int i = 0; i < n_streams; ++i) {
            float a = 0;
            cudaEventElapsedTime(&a, e0, e1[i]);
            worst = a > worst ? a : worst;
        }
        ms = worst;
        if (pass == 0) {        // calibration pass: rescale the number of rounds to the requested duration
            const double gbs = copies * (double)bytes / (ms * 1e-3) / 1e9;
            rounds = (long)(seconds * gbs * 1e9 / ((double)per_round * bytes)) + 1;
        }
    }
    if (cudaGetLastError() != cudaSuccess || ms <= 0) rc = B200S_ECUDA;
    else *d2h_gbs = copies * (double)bytes / (ms * 1e-3) / 1e9;
    if (((volatile unsigned char*)host[0])[bytes / 2] != 7) rc = B200S_ECUDA;     // the copies really arrived
    cleanup();
    return rc;
}

int b200s_copy_probe(int device, size_t bytes, double seconds, int host_mode, int with_h2d, double* d2h_gbs)
{
    return b200s_copy_probe_ex(device, bytes, seconds, host_mode, with_h2d, 2, 4, d2h_gbs);
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

}  // extern "C"
