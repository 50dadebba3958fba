// Rectification for sm_100a: cv::initUndistortRectifyMap-equivalent (FP64, no FMA contraction) and
// cv::remap's 5-bit fixed-point bilinear interpolation, optionally fused with the x-Sobel prefilter.
// Replaces PinholeCameraModel::rectifyImageGPU (reference call sites src/GPUStereoProcessor.cpp:244,248)
// with the CPU semantics of rectifyImageLeft/Right (:252-262), which is what the reference's goldens pin
// (test/UTest.cpp:247-256).  Algorithm: SURVEY.md A.1.
#include "kernels.h"

#include <climits>
#include <cstdlib>
#include <cstdint>

namespace b200s {

// (sx, sy) = (rint(float(u)*32), rint(float(v)*32)); every double op is an explicit round-to-nearest
// intrinsic so that ptxas can never contract a multiply-add (bit-exactness against the CPU evaluation).
__device__ __forceinline__ void map_uv(const CamModel& c, int j, int i, float& uf, float& vf)
{
    const double dj = (double)j, di = (double)i;
    double _x = __dadd_rn(__dmul_rn(dj, c.ir[0]), __dadd_rn(__dmul_rn(di, c.ir[1]), c.ir[2]));
    double _y = __dadd_rn(__dmul_rn(dj, c.ir[3]), __dadd_rn(__dmul_rn(di, c.ir[4]), c.ir[5]));
    double _w = __dadd_rn(__dmul_rn(dj, c.ir[6]), __dadd_rn(__dmul_rn(di, c.ir[7]), c.ir[8]));
    double w = __ddiv_rn(1.0, _w);
    double x = __dmul_rn(_x, w), y = __dmul_rn(_y, w);
    double x2 = __dmul_rn(x, x), y2 = __dmul_rn(y, y);
    double r2 = __dadd_rn(x2, y2);
    double _2xy = __dmul_rn(__dmul_rn(2.0, x), y);
    double num = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(c.k3, r2), c.k2), r2), c.k1), r2));
    double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(c.k6, r2), c.k5), r2), c.k4), r2));
    double kr = __ddiv_rn(num, den);
    double xd = __dadd_rn(__dadd_rn(__dmul_rn(x, kr), __dmul_rn(c.p1, _2xy)),
                          __dmul_rn(c.p2, __dadd_rn(r2, __dmul_rn(2.0, x2))));
    double yd = __dadd_rn(__dadd_rn(__dmul_rn(y, kr), __dmul_rn(c.p1, __dadd_rn(r2, __dmul_rn(2.0, y2)))),
                          __dmul_rn(c.p2, _2xy));
    double u = __dadd_rn(__dmul_rn(c.fx, xd), c.cx);
    double v = __dadd_rn(__dmul_rn(c.fy, yd), c.cy);
    uf = __double2float_rn(u);
    vf = __double2float_rn(v);
}

__device__ __forceinline__ int sat16(int v) { return max(-32768, min(32767, v)); }

// The integer part saturates to int16 like cv::convertMaps' CV_16SC2 coordinates; it is applied here once, so that every
// map entry (table or on the fly) already holds sat16(sx >> 5) * 32 + (sx & 31) and consumers may use entry >> 5 as it is.
__device__ __forceinline__ int2 map_point(const CamModel& c, int j, int i)
{
    float uf, vf;
    map_uv(c, j, i, uf, vf);
    const int sx = __float2int_rn(__fmul_rn(uf, 32.0f)), sy = __float2int_rn(__fmul_rn(vf, 32.0f));
    return make_int2(sat16(sx >> 5) * 32 + (sx & 31), sat16(sy >> 5) * 32 + (sy & 31));
}

__device__ __forceinline__ int fetch1(const uint8_t* __restrict__ s, int sW, int sH, int x, int y, int ch, int c)
{
    return ((unsigned)x < (unsigned)sW && (unsigned)y < (unsigned)sH) ? (int)__ldg(s + ((size_t)y * sW + x) * ch + c) : 0;
}

// one bilinear sample; weights (32-a)(32-b) etc. are the 15-bit table of cv::remap divided by 32
__device__ __forceinline__ int sample_linear(const uint8_t* __restrict__ s, int sW, int sH, int ch, int c, int2 m)
{
    int a = m.x & 31, b = m.y & 31;
    int X0 = sat16(m.x >> 5), Y0 = sat16(m.y >> 5);
    int s00 = fetch1(s, sW, sH, X0, Y0, ch, c), s01 = fetch1(s, sW, sH, X0 + 1, Y0, ch, c);
    int s10 = fetch1(s, sW, sH, X0, Y0 + 1, ch, c), s11 = fetch1(s, sW, sH, X0 + 1, Y0 + 1, ch, c);
    int acc = (32 - a) * (32 - b) * s00 + a * (32 - b) * s01 + (32 - a) * b * s10 + a * b * s11;
    return (acc + 512) >> 10;
}

// DELTA: 4 B/px table of (sx - 32 x, sy - 32 y) as int16 pairs; *overflow is raised when a delta does not fit
template <bool DELTA>
__global__ void __launch_bounds__(256) build_map_kernel(CamModel cm, int W, int H, void* __restrict__ map, int* __restrict__ overflow)
{
    int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const int2 m = map_point(cm, x, y);
    if (!DELTA) {
        ((int2*)map)[(size_t)y * W + x] = m;
        return;
    }
    const int dx = m.x - 32 * x, dy = m.y - 32 * y;
    if (dx < -32768 || dx > 32767 || dy < -32768 || dy > 32767) atomicOr(overflow, 1);
    ((uint32_t*)map)[(size_t)y * W + x] = ((uint32_t)dx & 0xffffu) | ((uint32_t)dy << 16);
}

template <int MODE>
__device__ __forceinline__ int2 map_at(const void* __restrict__ map, const CamModel& cm, int x, int y, int W)
{
    if (MODE == MAP_FLY) return map_point(cm, x, y);
    if (MODE == MAP_DELTA16) {
        const uint32_t v = __ldg((const uint32_t*)map + (size_t)y * W + x);
        return make_int2(32 * x + (int)(int16_t)(v & 0xffffu), 32 * y + ((int)v >> 16));
    }
    if (MODE == MAP_ABS32) return __ldg((const int2*)map + (size_t)y * W + x);
    return make_int2(32 * x, 32 * y);
}

template <int MODE, int CH>
__global__ void __launch_bounds__(256) remap_kernel(const uint8_t* __restrict__ src, int sW, int sH,
                                                    const void* __restrict__ map, CamModel cm,
                                                    uint8_t* __restrict__ dst, int W, int H, size_t src_stride, size_t dst_stride,
                                                    const uint8_t* const* __restrict__ src_tab)
{
    int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    src = src_tab ? src_tab[blockIdx.z] : src + blockIdx.z * src_stride;
    dst += blockIdx.z * dst_stride;
    const int2 m = map_at<MODE>(map, cm, x, y, W);
#pragma unroll
    for (int c = 0; c < CH; ++c) dst[((size_t)y * W + x) * CH + c] = (uint8_t)sample_linear(src, sW, sH, CH, c, m);
}

// cv::remap INTER_NEAREST on float maps: source pixel (cvRound(u), cvRound(v)); always evaluates the map
// (the cached x32 fixed-point map cannot reproduce the rounding of u itself).
__global__ void __launch_bounds__(256) remap_nearest_kernel(const uint8_t* __restrict__ src, int sW, int sH, int ch,
                                                            CamModel cm, uint8_t* __restrict__ dst, int W, int H)
{
    int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    float uf, vf;
    map_uv(cm, x, y, uf, vf);
    int X = __float2int_rn(uf), Y = __float2int_rn(vf);
    for (int c = 0; c < ch; ++c) dst[((size_t)y * W + x) * ch + c] = (uint8_t)fetch1(src, sW, sH, X, Y, ch, c);
}

// ---- fused rectify + x-Sobel ---------------------------------------------------------------------------
// Tile of TX x TY output pixels; the block rectifies the tile plus a 1-pixel halo into shared memory,
// writes the interior to `rect`, then applies OpenCV's prefilterXSobel border rules:
// 3x3 Sobel-x with rows mirrored (reflect-101), columns 0 and W-1 = cap, odd-height last row = cap.
#ifndef B200S_RECT_FTY
#define B200S_RECT_FTY 16
#endif
constexpr int FTX = 64, FTY = B200S_RECT_FTY;        // 16 tile rows = 256 threads; 8 rows = 128 threads (small enough in
                                                     // registers to share an SM with a resident matcher block)
constexpr int FT_THREADS = 16 * FTY;                 // the Sobel phase gives each thread 4 adjacent pixels of a row
constexpr int FT_N = (FTX + 2) * (FTY + 2);          // tile + halo pixels
constexpr int FT_PER = (FT_N + FT_THREADS - 1) / FT_THREADS;   // pixels per thread
constexpr int RECT_WIN_BYTES = 10240;                // shared-memory window of the source image per tile (e.g. 128 x 80)

struct RectSide {
    const uint8_t* src;
    const void* map;
    uint8_t* rect;
    uint8_t* pre;
    CamModel cm;
};
struct BatchStrides {      // bytes between consecutive frames of a batch (blockIdx.z = 2 * frame + side)
    size_t src, rect, pre;
    // optional device-resident tables of per-frame source addresses (frames that are not one strided buffer, e.g. the
    // caller's own device images): tabL[f] / tabR[f] replace src + f * stride
    const uint8_t* const* tabL;
    const uint8_t* const* tabR;
};
__device__ __forceinline__ const uint8_t* frame_src(const RectSide& S, const BatchStrides& bs, int side, int frame)
{
    const uint8_t* const* tab = side ? bs.tabR : bs.tabL;
    return tab ? tab[frame] : S.src + frame * bs.src;
}

// blockIdx.z selects frame and side, so that one launch rectifies and prefilters the left and the right images of a
// whole batch.  Each thread first fetches (or evaluates) the map entries of its FT_PER tile pixels, then issues all
// 4*FT_PER gathers, then blends: the dependent global loads of different pixels overlap.
// MODE = MAP_NONE: the source is already rectified (tile = source pixels, no rectified plane is written).
template <int MODE>
__global__ void __launch_bounds__(FT_THREADS) rectify_xsobel_kernel(RectSide sl, RectSide sr, BatchStrides bs, int sW, int sH,
                                                             size_t ppitch, int W, int H, int cap)
{
    __shared__ __align__(16) uint8_t tile[FTY + 2][FTX + 2 + 2];
    __shared__ __align__(16) uint8_t win[MODE != MAP_NONE ? RECT_WIN_BYTES : 16];
    __shared__ int bbox[4];
    const RectSide& S = (blockIdx.z & 1) ? sr : sl;
    const int frame = blockIdx.z >> 1;
    const uint8_t* __restrict__ src = frame_src(S, bs, blockIdx.z & 1, frame);
    const int x0 = blockIdx.x * FTX, y0 = blockIdx.y * FTY;
    int2 m[FT_PER];
    bool ok[FT_PER];
#pragma unroll
    for (int k = 0; k < FT_PER; ++k) {
        const int i = threadIdx.x + FT_THREADS * k;
        const int ty = i / (FTX + 2), tx = i - ty * (FTX + 2);
        const int x = x0 + tx - 1, y = y0 + ty - 1;
        // rows are mirrored for the Sobel taps: the halo row above row 0 is row 1, below row H-1 is row H-2
        const int ys = y < 0 ? 1 : (y >= H ? H - 2 : y);
        ok[k] = i < FT_N && x >= 0 && x < W && ys >= 0 && ys < H;
        m[k] = make_int2(0, 0);
        if (ok[k]) m[k] = map_at<MODE>(S.map, S.cm, x, ys, W);
    }
    // Source window of the tile: the 2x2 footprints of all tile pixels lie in a small bounding box of the source image
    // (a smooth map moves a 66x18 tile to roughly 70x22 source pixels).  The block finds that box (min / max of the
    // integer map parts), copies it into shared memory with coalesced 16-byte loads -- zeros outside the image, which is
    // cv::remap's BORDER_CONSTANT -- and gathers the taps from there: 4 byte-LDS per pixel instead of 4 scattered global
    // loads whose latency the kernel could not hide.  Boxes that do not fit (extreme maps) keep the global gathers.
    bool staged = false;
    int wx0 = 0, wy0 = 0, wpitch = 0;
    if (MODE != MAP_NONE) {
        int xmn = INT_MAX, xmx = INT_MIN, ymn = INT_MAX, ymx = INT_MIN;
#pragma unroll
        for (int k = 0; k < FT_PER; ++k)
            if (ok[k]) {
                const int X0 = sat16(m[k].x >> 5), Y0 = sat16(m[k].y >> 5);
                xmn = min(xmn, X0); xmx = max(xmx, X0); ymn = min(ymn, Y0); ymx = max(ymx, Y0);
            }
        xmn = __reduce_min_sync(0xffffffffu, xmn); xmx = __reduce_max_sync(0xffffffffu, xmx);
        ymn = __reduce_min_sync(0xffffffffu, ymn); ymx = __reduce_max_sync(0xffffffffu, ymx);
        if (threadIdx.x == 0) { bbox[0] = INT_MAX; bbox[1] = INT_MIN; bbox[2] = INT_MAX; bbox[3] = INT_MIN; }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&bbox[0], xmn); atomicMax(&bbox[1], xmx); atomicMin(&bbox[2], ymn); atomicMax(&bbox[3], ymx);
        }
        __syncthreads();
        xmn = bbox[0]; xmx = bbox[1]; ymn = bbox[2]; ymx = bbox[3];
        if (xmn <= xmx) {
            wx0 = xmn & ~15;                                   // floor to a multiple of 16 (also for negative x)
            wy0 = ymn;
            wpitch = (xmx + 2 - wx0 + 15) & ~15;
            const int wh = ymx + 2 - ymn;
            staged = (long long)wpitch * wh <= RECT_WIN_BYTES;
            if (staged) {
                const int upr = wpitch >> 4, nunits = upr * wh;
                const bool vec = ((sW & 15) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
                for (int u = threadIdx.x; u < nunits; u += FT_THREADS) {
                    const int row = u / upr, cx = u - row * upr;
                    const int y = wy0 + row, x = wx0 + 16 * cx;
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if ((unsigned)y < (unsigned)sH) {
                        const uint8_t* p = src + (size_t)y * sW + x;
                        if (vec && x >= 0 && x + 16 <= sW) {
                            v = __ldg((const uint4*)p);
                        } else {
                            uint32_t w4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if ((unsigned)(x + j) < (unsigned)sW) w4[j >> 2] |= (uint32_t)__ldg(p + j) << (8 * (j & 3));
                            v = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                        }
                    }
                    *(uint4*)(win + (size_t)row * wpitch + 16 * cx) = v;
                }
            }
        }
        __syncthreads();
    }
    int s00[FT_PER], s01[FT_PER], s10[FT_PER], s11[FT_PER];
#pragma unroll
    for (int k = 0; k < FT_PER; ++k) {
        const int X0 = sat16(m[k].x >> 5), Y0 = sat16(m[k].y >> 5);
        s00[k] = s01[k] = s10[k] = s11[k] = 0;
        if (MODE == MAP_NONE) {
            if (ok[k]) s00[k] = __ldg(src + (size_t)Y0 * sW + X0);      // identity map: a = b = 0, only s00 counts
        } else if (ok[k]) {
            if (staged) {
                const uint8_t* p = win + (Y0 - wy0) * wpitch + (X0 - wx0);
                s00[k] = p[0]; s01[k] = p[1]; s10[k] = p[wpitch]; s11[k] = p[wpitch + 1];
            } else if ((unsigned)X0 < (unsigned)(sW - 1) && (unsigned)Y0 < (unsigned)(sH - 1)) {
                // the whole 2x2 footprint is inside the source image (the common case): no per-tap border test
                const uint8_t* p = src + (size_t)Y0 * sW + X0;
                s00[k] = __ldg(p); s01[k] = __ldg(p + 1); s10[k] = __ldg(p + sW); s11[k] = __ldg(p + sW + 1);
            } else {
                s00[k] = fetch1(src, sW, sH, X0, Y0, 1, 0);
                s01[k] = fetch1(src, sW, sH, X0 + 1, Y0, 1, 0);
                s10[k] = fetch1(src, sW, sH, X0, Y0 + 1, 1, 0);
                s11[k] = fetch1(src, sW, sH, X0 + 1, Y0 + 1, 1, 0);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < FT_PER; ++k) {
        const int i = threadIdx.x + FT_THREADS * k;
        if (i >= FT_N) continue;
        const int ty = i / (FTX + 2), tx = i - ty * (FTX + 2);
        const int a = m[k].x & 31, b = m[k].y & 31;
        // (32-a)(32-b) s00 + a(32-b) s01 + (32-a) b s10 + a b s11, factored (exact in integers)
        const int top = 32 * s00[k] + a * (s01[k] - s00[k]), bot = 32 * s10[k] + a * (s11[k] - s10[k]);
        const int acc = 32 * top + b * (bot - top);
        tile[ty][tx] = (uint8_t)(ok[k] ? (acc + 512) >> 10 : 0);
    }
    __syncthreads();
    // each thread finishes 4 adjacent pixels with packed 16-bit arithmetic: per tile row two aligned words give the six
    // bytes x-1 .. x+4; one 32-bit store to the (pitched) prefiltered plane and one to the rectified plane
    {
        const int ty = threadIdx.x >> 4, tx = (threadIdx.x & 15) * 4;
        const int x = x0 + tx, y = y0 + ty;
        if (x < W && y < H) {
            const bool last_odd = (H & 1) && (y == H - 1);
            uint32_t de = 0, dO = 0, rect4 = 0;     // Sobel sums of pixels (0, 2) and (1, 3) as s16x2
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) {
                const uint32_t w0 = *(const uint32_t*)&tile[ty + rr][tx], w1 = *(const uint32_t*)&tile[ty + rr][tx + 4];
                const uint32_t hi = __funnelshift_r(w0, w1, 16);                      // bytes x+1 .. x+4
                const uint32_t le = w0 & 0x00ff00ffu, lo = (w0 >> 8) & 0x00ff00ffu;    // x-1, x+1 | x, x+2
                const uint32_t he = hi & 0x00ff00ffu, ho = (hi >> 8) & 0x00ff00ffu;    // x+1, x+3 | x+2, x+4
                uint32_t d_e = __vsub2(he, le), d_o = __vsub2(ho, lo);                 // pixels 0, 2 | pixels 1, 3
                if (rr == 1) {
                    d_e = __vadd2(d_e, d_e);
                    d_o = __vadd2(d_o, d_o);
                    rect4 = __funnelshift_r(w0, w1, 8);                                // bytes x .. x+3 of the centre row
                }
                de = __vadd2(de, d_e);
                dO = __vadd2(dO, d_o);
            }
            const uint32_t capw = (uint32_t)cap * 0x00010001u, ncapw = (uint32_t)(-cap & 0xffff) * 0x00010001u;
            de = __vadd2(__vmins2(__vmaxs2(de, ncapw), capw), capw);
            dO = __vadd2(__vmins2(__vmaxs2(dO, ncapw), capw), capw);
            uint32_t pre4 = __byte_perm(de, dO, 0x6240);                               // p0 p1 p2 p3
            if (last_odd || H <= 1) pre4 = (uint32_t)cap * 0x01010101u;
            else if (x == 0 || x + 3 >= W - 1) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (x + i == 0 || x + i >= W - 1) pre4 = (pre4 & ~(0xffu << (8 * i))) | ((uint32_t)cap << (8 * i));
            }
            uint8_t* pp = S.pre + frame * bs.pre + (size_t)y * ppitch + x;
            if (x + 3 < W) *(uint32_t*)pp = pre4;                    // ppitch % 16 == 0 and x % 4 == 0
            else for (int i = 0; i < 4 && x + i < W; ++i) pp[i] = (uint8_t)(pre4 >> (8 * i));
            if (MODE != MAP_NONE) {
                uint8_t* rp = S.rect + frame * bs.rect + (size_t)y * W + x;
                if (x + 3 < W && (W & 3) == 0 && (bs.rect & 3) == 0) *(uint32_t*)rp = rect4;
                else for (int i = 0; i < 4 && x + i < W; ++i) rp[i] = (uint8_t)(rect4 >> (8 * i));
            }
        }
    }
}

// ---- the same, four adjacent pixels per thread ("quad" form, the one launched for rectifying maps) --------------------
// The tile-with-halo is 18 rows of 17 quads (columns x0 - 2 .. x0 + 65), one quad per thread: the row terms (mirroring,
// validity, addresses) are formed once per four pixels, the four map entries arrive with one or two vector loads, and the
// four rectified bytes leave with one 32-bit shared-memory store.  Source window, blend and Sobel as above; same bytes.
constexpr int QT_QPR = (FTX + 4) / 4;                    // quads per tile row
constexpr int QT_ROWS = FTY + 2;
constexpr int QT_NQ = QT_QPR * QT_ROWS;
constexpr int QT_THREADS = ((QT_NQ > 16 * FTY ? QT_NQ : 16 * FTY) + 31) / 32 * 32;
constexpr int QT_PITCH = 4 * QT_QPR + 4;                 // tile row stride in bytes (the Sobel phase reads two whole words past a quad)

template <int MODE>
__global__ void __launch_bounds__(QT_THREADS, MODE == MAP_FLY ? 4 : 6) rectify_xsobel_quad_kernel(RectSide sl, RectSide sr, BatchStrides bs, int sW, int sH,
                                                                         size_t ppitch, int W, int H, int cap)
{
    static_assert(MODE != MAP_NONE, "the identity map keeps the per-pixel kernel");
    __shared__ __align__(16) uint8_t tile[QT_ROWS][QT_PITCH];
    __shared__ __align__(16) uint8_t win[RECT_WIN_BYTES];
    __shared__ int bbox[4];
    const RectSide& S = (blockIdx.z & 1) ? sr : sl;
    const int frame = blockIdx.z >> 1;
    const uint8_t* __restrict__ src = frame_src(S, bs, blockIdx.z & 1, frame);
    const int x0 = blockIdx.x * FTX, y0 = blockIdx.y * FTY;
    const int q = threadIdx.x;
    const bool qact = q < QT_NQ;
    const int ty = q / QT_QPR, tq = q - ty * QT_QPR;
    const int xq = x0 - 2 + 4 * tq, y = y0 + ty - 1;
    // rows are mirrored for the Sobel taps: the halo row above row 0 is row 1, below row H-1 is row H-2
    const int ys = y < 0 ? 1 : (y >= H ? H - 2 : y);
    const bool rowok = qact && ys >= 0 && ys < H;
    int mx[4], my[4];
    // bit k: pixel xq + k is inside the image
    const int klo = max(-xq, 0), khi = min(W - xq, 4);
    const unsigned okm = rowok && khi > klo ? ((1u << khi) - 1u) & ~((1u << klo) - 1u) : 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) mx[k] = my[k] = 0;
    if (okm == 0xFu && MODE == MAP_DELTA16 && !(W & 1)) {
        // xq is even and so is ys * W: the four 4-byte entries are two aligned 8-byte loads
        const uint2* mp = (const uint2*)((const uint32_t*)S.map + (size_t)ys * W + xq);
        const uint2 e0 = __ldg(mp), e1 = __ldg(mp + 1);
        const uint32_t v[4] = {e0.x, e0.y, e1.x, e1.y};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            mx[k] = 32 * (xq + k) + (int)(int16_t)(v[k] & 0xffffu);
            my[k] = 32 * ys + ((int)v[k] >> 16);
        }
    } else if (okm == 0xFu && MODE == MAP_ABS32 && !(W & 1)) {
        const int4* mp = (const int4*)((const int2*)S.map + (size_t)ys * W + xq);
        const int4 e0 = __ldg(mp), e1 = __ldg(mp + 1);
        mx[0] = e0.x; my[0] = e0.y; mx[1] = e0.z; my[1] = e0.w;
        mx[2] = e1.x; my[2] = e1.y; mx[3] = e1.z; my[3] = e1.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if ((okm >> k) & 1u) {
                const int2 m = map_at<MODE>(S.map, S.cm, xq + k, ys, W);
                mx[k] = m.x; my[k] = m.y;
            }
    }
    // source window of the tile (see rectify_xsobel_kernel); integer source coordinates, saturated like cv::remap's maps
    int X0[4], Y0[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { X0[k] = mx[k] >> 5; Y0[k] = my[k] >> 5; }      // map entries are saturated when they are made
    int xmn = INT_MAX, xmx = INT_MIN, ymn = INT_MAX, ymx = INT_MIN;
    if (okm == 0xFu) {
        xmn = min(min(X0[0], X0[1]), min(X0[2], X0[3])); xmx = max(max(X0[0], X0[1]), max(X0[2], X0[3]));
        ymn = min(min(Y0[0], Y0[1]), min(Y0[2], Y0[3])); ymx = max(max(Y0[0], Y0[1]), max(Y0[2], Y0[3]));
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if ((okm >> k) & 1u) { xmn = min(xmn, X0[k]); xmx = max(xmx, X0[k]); ymn = min(ymn, Y0[k]); ymx = max(ymx, Y0[k]); }
    }
    xmn = __reduce_min_sync(0xffffffffu, xmn); xmx = __reduce_max_sync(0xffffffffu, xmx);
    ymn = __reduce_min_sync(0xffffffffu, ymn); ymx = __reduce_max_sync(0xffffffffu, ymx);
    if (threadIdx.x == 0) { bbox[0] = INT_MAX; bbox[1] = INT_MIN; bbox[2] = INT_MAX; bbox[3] = INT_MIN; }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&bbox[0], xmn); atomicMax(&bbox[1], xmx); atomicMin(&bbox[2], ymn); atomicMax(&bbox[3], ymx);
    }
    __syncthreads();
    xmn = bbox[0]; xmx = bbox[1]; ymn = bbox[2]; ymx = bbox[3];
    bool staged = false;
    int wx0 = 0, wy0 = 0, wpitch = 0;
    if (xmn <= xmx) {
        wx0 = xmn & ~15;                                   // floor to a multiple of 16 (also for negative x)
        wy0 = ymn;
        wpitch = (xmx + 2 - wx0 + 15) & ~15;
        const int wh = ymx + 2 - ymn;
        staged = (long long)wpitch * wh <= RECT_WIN_BYTES;
        if (staged) {
            // 8 threads walk the 16-byte units of a window row, QT_THREADS / 8 rows at a time
            const int upr = wpitch >> 4;
            const bool vec = ((sW & 15) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
            for (int row = threadIdx.x >> 3; row < wh; row += QT_THREADS / 8) {
                const int yy = wy0 + row;
                for (int cx = threadIdx.x & 7; cx < upr; cx += 8) {
                    const int x = wx0 + 16 * cx;
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if ((unsigned)yy < (unsigned)sH) {
                        const uint8_t* p = src + (size_t)yy * sW + x;
                        if (vec && x >= 0 && x + 16 <= sW) {
                            v = __ldg((const uint4*)p);
                        } else {
                            uint32_t w4[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if ((unsigned)(x + j) < (unsigned)sW) w4[j >> 2] |= (uint32_t)__ldg(p + j) << (8 * (j & 3));
                            v = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                        }
                    }
                    *(uint4*)(win + (size_t)row * wpitch + 16 * cx) = v;
                }
            }
        }
    }
    __syncthreads();
    if (qact) {
        uint32_t out4 = 0;
        // (32-a)(32-b) s00 + a(32-b) s01 + (32-a) b s10 + a b s11, factored (exact in integers)
        auto blend = [](int mxk, int myk, int s00, int s01, int s10, int s11) {
            const int a = mxk & 31, b = myk & 31;
            const int top = 32 * s00 + a * (s01 - s00), bot = 32 * s10 + a * (s11 - s10);
            const int acc = 32 * top + b * (bot - top);
            return (uint32_t)((acc + 512) >> 10);
        };
        if (staged && okm == 0xFu) {
            const uint8_t* wbase = win - wy0 * wpitch - wx0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint8_t* p = wbase + Y0[k] * wpitch + X0[k];
                out4 |= blend(mx[k], my[k], p[0], p[1], p[wpitch], p[wpitch + 1]) << (8 * k);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!((okm >> k) & 1u)) continue;
                int s00, s01, s10, s11;
                if (staged) {
                    const uint8_t* p = win + (Y0[k] - wy0) * wpitch + (X0[k] - wx0);
                    s00 = p[0]; s01 = p[1]; s10 = p[wpitch]; s11 = p[wpitch + 1];
                } else if ((unsigned)X0[k] < (unsigned)(sW - 1) && (unsigned)Y0[k] < (unsigned)(sH - 1)) {
                    const uint8_t* p = src + (size_t)Y0[k] * sW + X0[k];
                    s00 = __ldg(p); s01 = __ldg(p + 1); s10 = __ldg(p + sW); s11 = __ldg(p + sW + 1);
                } else {
                    s00 = fetch1(src, sW, sH, X0[k], Y0[k], 1, 0);
                    s01 = fetch1(src, sW, sH, X0[k] + 1, Y0[k], 1, 0);
                    s10 = fetch1(src, sW, sH, X0[k], Y0[k] + 1, 1, 0);
                    s11 = fetch1(src, sW, sH, X0[k] + 1, Y0[k] + 1, 1, 0);
                }
                out4 |= blend(mx[k], my[k], s00, s01, s10, s11) << (8 * k);
            }
        }
        *(uint32_t*)&tile[ty][4 * tq] = out4;
    }
    __syncthreads();
    // Sobel: thread = 4 adjacent pixels x .. x + 3 (tile columns 4 t + 2 .. 4 t + 5, column 0 is x0 - 2) of two rows: four
    // tile rows give both, every row's horizontal differences are formed once
    if (threadIdx.x < 8 * FTY) {
        const int sy = (threadIdx.x >> 4) * 2, tx = (threadIdx.x & 15) * 4;
        const int x = x0 + tx, yy = y0 + sy;
        if (x < W && yy < H) {
            uint32_t dE[4], dOd[4], rc[4];           // per tile row: differences of pixels (0, 2) and (1, 3) as s16x2, bytes x .. x+3
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const uint32_t w0 = *(const uint32_t*)&tile[sy + rr][tx], w1 = *(const uint32_t*)&tile[sy + rr][tx + 4];
                const uint32_t lw = __funnelshift_r(w0, w1, 8);                       // bytes x-1 .. x+2
                const uint32_t hi = __funnelshift_r(w0, w1, 24);                      // bytes x+1 .. x+4
                const uint32_t le = lw & 0x00ff00ffu, lo = (lw >> 8) & 0x00ff00ffu;    // x-1, x+1 | x, x+2
                const uint32_t he = hi & 0x00ff00ffu, ho = (hi >> 8) & 0x00ff00ffu;    // x+1, x+3 | x+2, x+4
                dE[rr] = __vsub2(he, le);
                dOd[rr] = __vsub2(ho, lo);
                rc[rr] = __funnelshift_r(w0, w1, 16);
            }
            const uint32_t capw = (uint32_t)cap * 0x00010001u, ncapw = (uint32_t)(-cap & 0xffff) * 0x00010001u;
            const bool edge_cols = x == 0 || x + 3 >= W - 1;
            const bool vec_pre = x + 3 < W;                          // ppitch % 16 == 0 and x % 4 == 0
            const bool vec_rect = x + 3 < W && (W & 3) == 0 && (bs.rect & 3) == 0;
            uint8_t* pp = S.pre + frame * bs.pre + (size_t)yy * ppitch + x;
            uint8_t* rp = S.rect + frame * bs.rect + (size_t)yy * W + x;
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                if (yy + o >= H) break;
                uint32_t de = __vadd2(__vadd2(dE[o], dE[o + 2]), __vadd2(dE[o + 1], dE[o + 1]));
                uint32_t dO = __vadd2(__vadd2(dOd[o], dOd[o + 2]), __vadd2(dOd[o + 1], dOd[o + 1]));
                de = __vadd2(__vmins2(__vmaxs2(de, ncapw), capw), capw);
                dO = __vadd2(__vmins2(__vmaxs2(dO, ncapw), capw), capw);
                uint32_t pre4 = __byte_perm(de, dO, 0x6240);                           // p0 p1 p2 p3
                const bool last_odd = (H & 1) && (yy + o == H - 1);
                if (last_odd || H <= 1) pre4 = (uint32_t)cap * 0x01010101u;
                else if (edge_cols) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (x + i == 0 || x + i >= W - 1) pre4 = (pre4 & ~(0xffu << (8 * i))) | ((uint32_t)cap << (8 * i));
                }
                const uint32_t rect4 = rc[o + 1];
                if (vec_pre) *(uint32_t*)pp = pre4;
                else for (int i = 0; i < 4 && x + i < W; ++i) pp[i] = (uint8_t)(pre4 >> (8 * i));
                if (vec_rect) *(uint32_t*)rp = rect4;
                else for (int i = 0; i < 4 && x + i < W; ++i) rp[i] = (uint8_t)(rect4 >> (8 * i));
                pp += ppitch;
                rp += W;
            }
        }
    }
}

// ---- fused (rectify +) normalised-response prefilter -------------------------------------------------------
// cv::StereoBM's NORMALIZED_RESPONSE prefilter: out = clip((c * scale_g - boxsum * scale_s) >> 10) + cap
// with c = 4 I(x,y) + its 4 neighbours and boxsum over preFilterSize^2, replicate borders.  One block rectifies (or just
// loads) a 64x16 tile plus a halo of preFilterSize/2 at clamped coordinates into shared memory, forms the vertical box
// sums of every tile column with a sliding sum, then every thread finishes 4 adjacent pixels with a sliding horizontal
// sum.  Replaces remap x2 + two prefilter passes through a global scratch plane; both sides in one launch.
constexpr int NTX = 64, NTY = 16, NP2MAX = 10;                 // preFilterSize <= 21
constexpr int NTW = NTX + 2 * NP2MAX + 4;                      // tile row stride (bytes / u16 entries)

template <int MODE>     // MapMode: MAP_NONE = source is already rectified
__global__ void __launch_bounds__(256) norm_prefilter_kernel(RectSide sl, RectSide sr, BatchStrides bs, int sW, int sH,
                                                             size_t ppitch, int W, int H, int p2, int scale_g, int scale_s, int cap)
{
    __shared__ __align__(16) uint8_t tile[NTY + 2 * NP2MAX][NTW];
    __shared__ __align__(16) uint16_t vs[NTY][NTW];
    const RectSide& S = (blockIdx.z & 1) ? sr : sl;
    const int frame = blockIdx.z >> 1;
    const uint8_t* __restrict__ src = frame_src(S, bs, blockIdx.z & 1, frame);
    const int x0 = blockIdx.x * NTX, y0 = blockIdx.y * NTY;
    const int tw = NTX + 2 * p2, th = NTY + 2 * p2, tn = tw * th;
    for (int i0 = threadIdx.x; i0 < tn; i0 += 4 * 256) {
        int2 m[4];
        int xs[4], ys[4];
        bool in[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = i0 + 256 * k;
            in[k] = i < tn;
            const int ty = i / tw, tx = i - ty * tw;
            xs[k] = min(max(x0 + tx - p2, 0), W - 1);           // replicate border of the (rectified) image
            ys[k] = min(max(y0 + ty - p2, 0), H - 1);
            m[k] = make_int2(0, 0);
            if (MODE != MAP_NONE && in[k]) m[k] = map_at<MODE>(S.map, S.cm, xs[k], ys[k], W);
        }
        int v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] = 0;
            if (!in[k]) continue;
            if (MODE == MAP_NONE) v[k] = __ldg(src + (size_t)ys[k] * sW + xs[k]);
            else v[k] = sample_linear(src, sW, sH, 1, 0, m[k]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = i0 + 256 * k;
            if (!in[k]) continue;
            const int ty = i / tw, tx = i - ty * tw;
            tile[ty][tx] = (uint8_t)v[k];
        }
    }
    __syncthreads();
    // vertical box sums: thread = tile column, sliding down the 16 output rows (255 * 21 fits 16 bits)
    if (threadIdx.x < tw) {
        const int c = threadIdx.x;
        int s = 0;
        for (int j = 0; j <= 2 * p2; ++j) s += tile[j][c];
        vs[0][c] = (uint16_t)s;
        for (int y = 1; y < NTY; ++y) {
            s += (int)tile[y + 2 * p2][c] - (int)tile[y - 1][c];
            vs[y][c] = (uint16_t)s;
        }
    }
    __syncthreads();
    {
        const int ty = threadIdx.x >> 4, tx = (threadIdx.x & 15) * 4;
        const int x = x0 + tx, y = y0 + ty;
        if (x < W && y < H) {
            int sum = 0;
            for (int j = 0; j <= 2 * p2; ++j) sum += vs[ty][tx + j];
            uint32_t pre4 = 0, rect4 = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint8_t* cr = &tile[ty + p2][tx + i + p2];
                const int c = 4 * (int)cr[0] + (int)cr[-1] + (int)cr[1] + (int)cr[-NTW] + (int)cr[NTW];
                const int val = (c * scale_g - sum * scale_s) >> 10;
                pre4 |= (uint32_t)(min(max(val, -cap), cap) + cap) << (8 * i);
                rect4 |= (uint32_t)cr[0] << (8 * i);
                if (i < 3) sum += (int)vs[ty][tx + i + 2 * p2 + 1] - (int)vs[ty][tx + i];
            }
            uint8_t* pp = S.pre + frame * bs.pre + (size_t)y * ppitch + x;
            if (x + 3 < W) *(uint32_t*)pp = pre4;                // ppitch % 16 == 0 and x % 4 == 0
            else for (int i = 0; i < 4 && x + i < W; ++i) pp[i] = (uint8_t)(pre4 >> (8 * i));
            if (MODE != MAP_NONE && S.rect) {
                uint8_t* rp = S.rect + frame * bs.rect + (size_t)y * W + x;
                if (x + 3 < W && (W & 3) == 0 && (bs.rect & 3) == 0) *(uint32_t*)rp = rect4;
                else for (int i = 0; i < 4 && x + i < W; ++i) rp[i] = (uint8_t)(rect4 >> (8 * i));
            }
        }
    }
}

static inline dim3 grid2d(int W, int H, int nf = 1) { return dim3((W + 31) / 32, (H + 7) / 8, nf); }

// returns 0 when preFilterSize is too large for the tile kernel (the caller uses the two-pass kernels then)
int launch_norm_prefilter_pair(const uint8_t* srcL, const uint8_t* srcR, int sW, int sH, MapMode mode, const void* mapL,
                               const void* mapR, const CamModel& cmL, const CamModel& cmR, uint8_t* rectL, uint8_t* rectR,
                               uint8_t* preL, uint8_t* preR, size_t pre_pitch, int W, int H, int ps, int cap, cudaStream_t st,
                               int nf, size_t src_stride, size_t rect_stride, size_t pre_stride, const uint8_t* const* tabL,
                               const uint8_t* const* tabR)
{
    const int p2 = ps / 2;
    if (p2 > NP2MAX) return 0;
    int scale_g = ps * ps / 8, scale_s = (1024 + scale_g) / (scale_g * 2);
    scale_g *= scale_s;
    dim3 g((W + NTX - 1) / NTX, (H + NTY - 1) / NTY, 2 * nf);
    RectSide l{srcL, mapL, rectL, preL, cmL}, r{srcR, mapR, rectR, preR, cmR};
    BatchStrides bs{src_stride, rect_stride, pre_stride, tabL, tabR};
    switch (mode) {
    case MAP_NONE: norm_prefilter_kernel<MAP_NONE><<<g, 256, 0, st>>>(l, r, bs, sW, sH, pre_pitch, W, H, p2, scale_g, scale_s, cap); break;
    case MAP_ABS32: norm_prefilter_kernel<MAP_ABS32><<<g, 256, 0, st>>>(l, r, bs, sW, sH, pre_pitch, W, H, p2, scale_g, scale_s, cap); break;
    case MAP_DELTA16: norm_prefilter_kernel<MAP_DELTA16><<<g, 256, 0, st>>>(l, r, bs, sW, sH, pre_pitch, W, H, p2, scale_g, scale_s, cap); break;
    default: norm_prefilter_kernel<MAP_FLY><<<g, 256, 0, st>>>(l, r, bs, sW, sH, pre_pitch, W, H, p2, scale_g, scale_s, cap); break;
    }
    return 1;
}

int launch_build_map(const CamModel& cm, int W, int H, void* map, MapMode mode, int* overflow, cudaStream_t st)
{
    if (mode == MAP_DELTA16) {
        cudaMemsetAsync(overflow, 0, sizeof(int), st);
        build_map_kernel<true><<<grid2d(W, H), 256, 0, st>>>(cm, W, H, map, overflow);
    } else {
        build_map_kernel<false><<<grid2d(W, H), 256, 0, st>>>(cm, W, H, map, overflow);
    }
    return 1;
}

template <int CH>
static void launch_remap_ch(const uint8_t* src, int sW, int sH, const void* map, MapMode mode, const CamModel& cm, uint8_t* dst,
                            int W, int H, cudaStream_t st, int nf, size_t ss, size_t ds, const uint8_t* const* tab)
{
    const dim3 g = grid2d(W, H, nf);
    switch (mode) {
    case MAP_ABS32: remap_kernel<MAP_ABS32, CH><<<g, 256, 0, st>>>(src, sW, sH, map, cm, dst, W, H, ss, ds, tab); break;
    case MAP_DELTA16: remap_kernel<MAP_DELTA16, CH><<<g, 256, 0, st>>>(src, sW, sH, map, cm, dst, W, H, ss, ds, tab); break;
    default: remap_kernel<MAP_FLY, CH><<<g, 256, 0, st>>>(src, sW, sH, map, cm, dst, W, H, ss, ds, tab); break;
    }
}

int launch_remap(const uint8_t* src, int sW, int sH, int ch, const void* map, MapMode mode, const CamModel& cm, uint8_t* dst,
                 int W, int H, cudaStream_t st, int nf, size_t src_stride, size_t dst_stride, const uint8_t* const* src_tab)
{
    if (!map) mode = MAP_FLY;
    if (ch == 1) launch_remap_ch<1>(src, sW, sH, map, mode, cm, dst, W, H, st, nf, src_stride, dst_stride, src_tab);
    else if (ch == 3) launch_remap_ch<3>(src, sW, sH, map, mode, cm, dst, W, H, st, nf, src_stride, dst_stride, src_tab);
    else if (ch == 4) launch_remap_ch<4>(src, sW, sH, map, mode, cm, dst, W, H, st, nf, src_stride, dst_stride, src_tab);
    else return -1;
    return 1;
}

int launch_remap_nearest(const uint8_t* src, int sW, int sH, int ch, const CamModel& cm, uint8_t* dst, int W, int H,
                         cudaStream_t st)
{
    remap_nearest_kernel<<<grid2d(W, H), 256, 0, st>>>(src, sW, sH, ch, cm, dst, W, H);
    return 1;
}

int launch_rectify_xsobel_pair(const uint8_t* srcL, const uint8_t* srcR, int sW, int sH, MapMode mode, const void* mapL,
                               const void* mapR, const CamModel& cmL, const CamModel& cmR, uint8_t* rectL, uint8_t* rectR,
                               uint8_t* preL, uint8_t* preR, size_t pre_pitch, int W, int H, int cap, cudaStream_t st,
                               int nf, size_t src_stride, size_t rect_stride, size_t pre_stride, const uint8_t* const* tabL,
                               const uint8_t* const* tabR)
{
    dim3 g((W + FTX - 1) / FTX, (H + FTY - 1) / FTY, 2 * nf);
    RectSide l{srcL, mapL, rectL, preL, cmL}, r{srcR, mapR, rectR, preR, cmR};
    BatchStrides bs{src_stride, rect_stride, pre_stride, tabL, tabR};
    static const int use_quad = getenv("B200S_RECT_QUAD") ? atoi(getenv("B200S_RECT_QUAD")) : 1;
    if (use_quad && mode != MAP_NONE) {
        switch (mode) {
        case MAP_ABS32: rectify_xsobel_quad_kernel<MAP_ABS32><<<g, QT_THREADS, 0, st>>>(l, r, bs, sW, sH, pre_pitch, W, H, cap); break;
        case MAP_DELTA16: rectify_xsobel_quad_kernel<MAP_DELTA16><<<g, QT_THREADS, 0, st>>>(l, r, bs, sW, sH, pre_pitch, W, H, cap); break;
        default: rectify_xsobel_quad_kernel<MAP_FLY><<<g, QT_THREADS, 0, st>>>(l, r, bs, sW, sH, pre_pitch, W, H, cap); break;
        }
        return 1;
    }
    switch (mode) {
    case MAP_NONE: rectify_xsobel_kernel<MAP_NONE><<<g, FT_THREADS, 0, st>>>(l, r, bs, sW, sH, pre_pitch, W, H, cap); break;
    case MAP_ABS32: rectify_xsobel_kernel<MAP_ABS32><<<g, FT_THREADS, 0, st>>>(l, r, bs, sW, sH, pre_pitch, W, H, cap); break;
    case MAP_DELTA16: rectify_xsobel_kernel<MAP_DELTA16><<<g, FT_THREADS, 0, st>>>(l, r, bs, sW, sH, pre_pitch, W, H, cap); break;
    default: rectify_xsobel_kernel<MAP_FLY><<<g, FT_THREADS, 0, st>>>(l, r, bs, sW, sH, pre_pitch, W, H, cap); break;
    }
    return 1;
}

}  // namespace b200s
