// Disparity -> float plane, cv::reprojectImageTo3D(handleMissingValues = true) and PointCloud2 / DisparityImage
// payload packing in one pass.  Replaces projectDisparityImageTo3dGPU
// (src/GPUStereoProcessor.cpp:332-346) and the host loops of GPUSenderPc2::fillInData (src/GpuSenderPc2.cpp:15-72)
// and GPUSenderDisparity::fillInData (src/GpuSenderDisparity.cpp:18-48).
// The 4-term products are evaluated in FP64 without FMA contraction, like the CPU code, so the points are
// bit-identical to cv2's (the 1e-5 bar of the north star needs FP64: pure f32 reaches 2e-5).
#include "kernels.h"

#include <cstdlib>

#include <climits>

namespace b200s {

__device__ __forceinline__ float disp_to_float(int d16, double cxd)
{
    // cv::Mat::convertTo(CV_32F, 1/16., -(cx_l - cx_r)): saturate_cast<float>(d * alpha + beta) in double
    return __double2float_rn(__dadd_rn(__dmul_rn((double)d16, 1.0 / 16.0), -cxd));
}

__global__ void __launch_bounds__(256) set_int_kernel(int* p, int v) { *p = v; }

struct FrameDst {          // per-frame destinations of a batch: strided planes, or one caller-owned (mapped pinned) buffer per frame
    size_t stride;
    int use_list;
    PtrList list;
};

// 8 disparities per thread as two groups of 4, 128 floats apart inside the warp's 256-element chunk: every store
// instruction of a warp writes 512 contiguous bytes (full lines for HBM and for posted PCIe writes into pinned host
// memory alike); block-level min, one atomic per block; blockIdx.y = frame of the batch
__global__ void __launch_bounds__(256) disparity_to_float_kernel(const int16_t* __restrict__ d16, float* __restrict__ df,
                                                                 int n, double cxd, int* __restrict__ min_d16, size_t d_stride,
                                                                 const FrameDst fd)
{
    __shared__ int wmin[8];
    d16 = (const int16_t*)((const uint8_t*)d16 + blockIdx.y * d_stride);
    if (fd.use_list) df = (float*)fd.list.p[blockIdx.y];
    else if (df) df = (float*)((uint8_t*)df + blockIdx.y * fd.stride);
    min_d16 += blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int base = blockIdx.x * 2048 + warp * 256;
    int v = INT_MAX;
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        const int i0 = base + g * 128 + 4 * lane;
        if (i0 + 4 <= n) {
            const uint2 q = *(const uint2*)(d16 + i0);
            const int a = (int)(int16_t)(q.x & 0xffffu), b = (int)(int16_t)(q.x >> 16);
            const int c = (int)(int16_t)(q.y & 0xffffu), d = (int)(int16_t)(q.y >> 16);
            v = min(v, min(min(a, b), min(c, d)));
            if (df) *(float4*)(df + i0) = make_float4(disp_to_float(a, cxd), disp_to_float(b, cxd), disp_to_float(c, cxd), disp_to_float(d, cxd));
        } else {
            for (int i = i0; i < n; ++i) {
                const int a = d16[i];
                v = min(v, a);
                if (df) df[i] = disp_to_float(a, cxd);
            }
        }
    }
    v = __reduce_min_sync(0xffffffffu, v);
    if (lane == 0) wmin[warp] = v;
    __syncthreads();
    if (threadIdx.x < 8) {
        v = wmin[threadIdx.x];
        v = __reduce_min_sync(0xffu, v);
        if (threadIdx.x == 0 && v != INT_MAX) atomicMin(min_d16, v);
    }
}

#ifndef B200S_PACK_ROWS
#define B200S_PACK_ROWS 8
#endif
constexpr int RP_ROWS = B200S_PACK_ROWS;      // block = 32 columns x RP_ROWS warps
constexpr int RP_RPW = 4;                     // consecutive image rows per warp (column-only terms are computed once)
#ifndef B200S_PACK_PREFETCH
#define B200S_PACK_PREFETCH 0                 // 1: issue the loads of all rows of a warp before the arithmetic of the first
#endif                                        // (measured: 27.6 us at C4 against 25.8 us with the loads inside the row loop)

// bit (4 r + c) set = Q[r][c] != 0.  image_geometry's Q (StereoCameraModel::updateQ) has exactly these entries:
//   X = Q00 x + Q03,  Y = Q11 y + Q13,  Z = Q23,  W = Q32 d + Q33      (Q33 = fy (cx - cx') is zero for equal principal points)
constexpr unsigned QMASK_STEREO0 = (1u << 0) | (1u << 3) | (1u << 5) | (1u << 7) | (1u << 11) | (1u << 14);
constexpr unsigned QMASK_STEREO = QMASK_STEREO0 | (1u << 15);

// float(a / w): one correctly rounded reciprocal shared by the three coordinates, see reproject_pack_kernel
__device__ __forceinline__ float div_to_float(double a, double w, double rw)
{
    const double q = __dmul_rn(a, rw);
    const unsigned lo = (unsigned)__double2loint(q) & 0x1FFFFFFFu;
    const unsigned hi = (unsigned)__double2hiint(q) & 0x7FFFFFFFu;
    // |q| in [2^-100, 2^100] (finite, far from float denormals / overflow) and away from the float rounding boundary
    const bool safe = hi > 0x39B00000u && hi < 0x46300000u && (lo - 0x0FFFFFF0u) > 0x20u;
    return __double2float_rn(safe ? q : __ddiv_rn(a, w));
}

constexpr uint32_t LUT_INVALID_Z = 0x7f800001u;     // a signalling NaN: no float conversion produces it

// w of the standard Q and the record terms that depend on the disparity only, exactly as the per-pixel path forms them
template <int STDQ>
__device__ __forceinline__ double stdq_w(double q32, double q33, double d)
{
    double w = __dadd_rn(0.0, __dmul_rn(q32, d));
    if (STDQ == 1) w = __dadd_rn(w, q33);
    return w;
}

template <int STDQ>
__global__ void __launch_bounds__(256) reproject_lut_kernel(uint4* __restrict__ lut, int n, int dmin, double cxd,
                                                            const double* __restrict__ Q)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int dv = dmin + i;
    const float dfv = disp_to_float(dv, cxd);
    const double d = (double)dfv, minDisp = (double)disp_to_float(dmin, cxd);
    const double az = (double)__double2float_rn(__dadd_rn(0.0, __ldg(Q + 11)));
    const double w = stdq_w<STDQ>(__ldg(Q + 14), __ldg(Q + 15), d);
    const double rw = __drcp_rn(w);
    float z = div_to_float(az, w, rw);
    if (fabs(__dadd_rn(d, -minDisp)) <= (double)1.1920928955078125e-07f) z = 10000.0f;
    const bool valid = (z != 10000.0f) && !isinf(z);
    lut[i] = make_uint4((uint32_t)__double2loint(rw), (uint32_t)__double2hiint(rw), valid ? __float_as_uint(z) : LUT_INVALID_Z,
                        __float_as_uint(dfv));
}

// cv::reprojectImageTo3D(handleMissingValues = true) + PointCloud2 records (+ the float disparity plane) in one pass.
// STDQ (1: with Q33, 2: Q33 == 0, 0: any Q): Q has the sparsity of image_geometry's stereo model (the only Q the reference
// can produce); the terms that depend on the column only are then computed once per thread and reused for the RP_RPW rows
// its warp walks.  The generic path evaluates the full 4-term products with zero entries skipped.  Both give the bytes of the CPU code: every FP64
// operation is an explicit round-to-nearest intrinsic in the CPU's order.
//
// float(a / w) for the three coordinates, a = (double)(float)h[r]: the correctly rounded FP64 division is a long
// instruction sequence and all three share the divisor, so one correctly rounded reciprocal and a multiplication give a
// quotient within 2 ulp (FP64) of a / w; rounding THAT to float gives the same float as rounding the exact quotient unless
// it lies within a few FP64 ulps of a float rounding boundary (29 dropped bits = 0x10000000).  Those rare lanes, and
// non-finite or tiny quotients, take the exact division: the result is bit-identical always.
// LUT (STDQ only, records without the xyz plane): the per-disparity terms come from the table of reproject_lut_kernel --
// the same operations, done once per value instead of once per pixel; disparities outside the table take the arithmetic.
template <int STDQ, bool LUT>
#if B200S_PACK_PREFETCH
#define B200S_PACK_BOUNDS __launch_bounds__(32 * RP_ROWS, 1536 / (32 * RP_ROWS))    /* 40 registers: six blocks per SM */
#else
#define B200S_PACK_BOUNDS __launch_bounds__(32 * RP_ROWS)
#endif
__global__ void B200S_PACK_BOUNDS reproject_pack_kernel(const int16_t* __restrict__ d16, int W, int H, double cxd,
                                                             const double* __restrict__ Q, const int* __restrict__ min_d16,
                                                             const uint8_t* __restrict__ color, int ch,
                                                             float* __restrict__ xyz, uint8_t* __restrict__ pc2, unsigned qmask,
                                                             size_t d_stride, size_t color_stride, size_t xyz_stride, const FrameDst fd,
                                                             int dmin_const, float* __restrict__ df, const FrameDst fdf,
                                                             const uint8_t* const* __restrict__ color_tab,
                                                             const uint4* __restrict__ lut, int lut_n)
{
    const int lane = threadIdx.x & 31;
    const int x0 = blockIdx.x * 32, x = x0 + lane;
    const int yw = (blockIdx.y * RP_ROWS + (threadIdx.x >> 5)) * RP_RPW;      // first row of this warp
    {
        const int f = blockIdx.z;          // frame of the batch
        d16 = (const int16_t*)((const uint8_t*)d16 + f * d_stride);
        if (color_tab) color = color_tab[f];
        else if (color) color += f * color_stride;
        if (xyz) xyz = (float*)((uint8_t*)xyz + f * xyz_stride);
        if (fd.use_list) pc2 = (uint8_t*)fd.list.p[f];
        else if (pc2) pc2 += f * fd.stride;
        if (fdf.use_list) df = (float*)fdf.list.p[f];
        else if (df) df = (float*)((uint8_t*)df + f * fdf.stride);
        if (min_d16) min_d16 += f;
    }
    if (yw >= H) return;
    const int dmin = min_d16 ? *min_d16 : dmin_const;
    const double minDisp = (double)disp_to_float(dmin, cxd);
    // column-only and constant terms of the standard Q
    double q11 = 0, q13 = 0, q32 = 0, q33 = 0, ax = 0, az = 0;
    if (STDQ) {
        ax = (double)__double2float_rn(__dadd_rn(__dmul_rn(__ldg(Q + 0), (double)x), __ldg(Q + 3)));
        q11 = __ldg(Q + 5); q13 = __ldg(Q + 7);
        az = (double)__double2float_rn(__dadd_rn(0.0, __ldg(Q + 11)));
        q32 = __ldg(Q + 14); q33 = __ldg(Q + 15);
    }
    const bool xin = x < W;
    const bool full_row = x0 + 32 <= W;
    // all global loads of the warp's rows first (the kernel waited on one dependent load per row otherwise)
    int dvs[RP_RPW];
    uint32_t bgrs[RP_RPW];
#pragma unroll
    for (int rr = 0; rr < RP_RPW; ++rr) {
        const int y = yw + rr;
        dvs[rr] = 0;
        bgrs[rr] = 0;
        if (B200S_PACK_PREFETCH && xin && y < H) {
            const size_t i = (size_t)y * W + x;
            dvs[rr] = (int)d16[i];
            if (pc2) {
                if (ch == 3) bgrs[rr] = (uint32_t)color[i * 3] | ((uint32_t)color[i * 3 + 1] << 8) | ((uint32_t)color[i * 3 + 2] << 16);
                else { const uint32_t g = color ? color[i] : 0; bgrs[rr] = g | (g << 8) | (g << 16); }
            }
        }
    }
#pragma unroll
    for (int rr = 0; rr < RP_RPW; ++rr) {
        const int y = yw + rr;
        if (y >= H) break;                                   // warp-uniform
        // One PointCloud2 record = 32 bytes = two 16-byte halves, [x y z 0] and [bgr 0 0 0].  The 32 records of a warp
        // row are 1 KiB contiguous; lanes exchange halves so that each of the two store instructions of the warp writes
        // 512 contiguous bytes (lane l: half (l & 1) of record (l >> 1) + 16 * instruction) -- full lines for HBM and
        // for posted PCIe writes when pc2 is pinned host memory.  Partial warps at the right edge store per thread.
        uint32_t ux = 0x7fc00000u, uy = 0x7fc00000u, uz = 0x7fc00000u;
        const size_t i = xin ? (size_t)y * W + x : 0;
        const int dv = B200S_PACK_PREFETCH ? dvs[rr] : (xin ? (int)d16[i] : 0);
        const unsigned li = (unsigned)(dv - dmin);
        if (LUT && li < (unsigned)lut_n) {
            const uint4 e = xin ? __ldg(lut + li) : make_uint4(0u, 0u, LUT_INVALID_Z, 0u);
            if (df && xin) df[i] = __uint_as_float(e.w);
            if (xin && e.z != LUT_INVALID_Z) {
                const double rw = __hiloint2double((int)e.y, (int)e.x);
                const double ay = (double)__double2float_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(q11, (double)y)), q13));
                double w = 0.0;
                float p0, p1;
                {
                    const double q = __dmul_rn(ax, rw);
                    const unsigned lo = (unsigned)__double2loint(q) & 0x1FFFFFFFu, hi = (unsigned)__double2hiint(q) & 0x7FFFFFFFu;
                    const bool safe = hi > 0x39B00000u && hi < 0x46300000u && (lo - 0x0FFFFFF0u) > 0x20u;
                    if (!safe) w = stdq_w<STDQ>(q32, q33, (double)__uint_as_float(e.w));
                    p0 = __double2float_rn(safe ? q : __ddiv_rn(ax, w));
                }
                {
                    const double q = __dmul_rn(ay, rw);
                    const unsigned lo = (unsigned)__double2loint(q) & 0x1FFFFFFFu, hi = (unsigned)__double2hiint(q) & 0x7FFFFFFFu;
                    const bool safe = hi > 0x39B00000u && hi < 0x46300000u && (lo - 0x0FFFFFF0u) > 0x20u;
                    if (!safe) w = stdq_w<STDQ>(q32, q33, (double)__uint_as_float(e.w));
                    p1 = __double2float_rn(safe ? q : __ddiv_rn(ay, w));
                }
                ux = __float_as_uint(p0); uy = __float_as_uint(p1); uz = e.z;
            }
        } else {
        if (df && xin) df[i] = disp_to_float(dv, cxd);       // the DisparityImage payload from the same pass (convertTo)
        // missing value (d == min over the image): cv::reprojectImageTo3D sets Z = 10000, which isValidPoint rejects, so
        // the record is NaN xyz + colour whatever X and Y were; no arithmetic needed unless the xyz plane is wanted too
        if (xin && (xyz || dv != dmin)) {
            const double d = (double)disp_to_float(dv, cxd);
            double a[3], w;
            if (STDQ) {
                a[0] = ax;
                a[1] = (double)__double2float_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(q11, (double)y)), q13));
                a[2] = az;
                w = stdq_w<STDQ>(q32, q33, d);
            } else {
                double h[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    // entries of Q that are exactly zero are skipped (qmask bit = entry is non-zero): 0 * v = +-0 and
                    // s + (+-0) = s, so the sum is bit-identical to the full 4-term product (finite v)
                    double s2 = (qmask >> (r * 4 + 0)) & 1u ? __dmul_rn(__ldg(Q + r * 4 + 0), (double)x) : 0.0;
                    if ((qmask >> (r * 4 + 1)) & 1u) s2 = __dadd_rn(s2, __dmul_rn(__ldg(Q + r * 4 + 1), (double)y));
                    if ((qmask >> (r * 4 + 2)) & 1u) s2 = __dadd_rn(s2, __dmul_rn(__ldg(Q + r * 4 + 2), d));
                    if ((qmask >> (r * 4 + 3)) & 1u) s2 = __dadd_rn(s2, __ldg(Q + r * 4 + 3));
                    h[r] = s2;
                }
                a[0] = (double)__double2float_rn(h[0]); a[1] = (double)__double2float_rn(h[1]); a[2] = (double)__double2float_rn(h[2]);
                w = h[3];
            }
            const double rw = __drcp_rn(w);
            float p[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) p[r] = div_to_float(a[r], w, rw);
            if (fabs(__dadd_rn(d, -minDisp)) <= (double)1.1920928955078125e-07f) p[2] = 10000.0f;
            if (xyz) {
                xyz[i * 3] = p[0]; xyz[i * 3 + 1] = p[1]; xyz[i * 3 + 2] = p[2];
            }
            // isValidPoint (src/GpuSenderPc2.cpp:84-89): z != MISSING_Z and not inf; invalid -> quiet NaN
            if ((p[2] != 10000.0f) && !isinf(p[2])) {
                ux = __float_as_uint(p[0]); uy = __float_as_uint(p[1]); uz = __float_as_uint(p[2]);
            }
        }
        }
        if (!pc2) continue;
        uint32_t bgr = bgrs[rr];
        if (!B200S_PACK_PREFETCH && xin) {
            if (ch == 3) bgr = (uint32_t)color[i * 3] | ((uint32_t)color[i * 3 + 1] << 8) | ((uint32_t)color[i * 3 + 2] << 16);
            else { const uint32_t g = color ? color[i] : 0; bgr = g | (g << 8) | (g << 16); }
        }
        if (full_row) {
            uint4* row = (uint4*)(pc2 + ((size_t)y * W + x0) * 32);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int src = (lane >> 1) + 16 * k;
                const uint32_t sx = __shfl_sync(0xffffffffu, ux, src), sy = __shfl_sync(0xffffffffu, uy, src);
                const uint32_t sz = __shfl_sync(0xffffffffu, uz, src), sc = __shfl_sync(0xffffffffu, bgr, src);
                row[32 * k + lane] = (lane & 1) ? make_uint4(sc, 0u, 0u, 0u) : make_uint4(sx, sy, sz, 0u);
            }
        } else if (xin) {
            uint4* o = (uint4*)(pc2 + i * 32);
            o[0] = make_uint4(ux, uy, uz, 0u);
            o[1] = make_uint4(bgr, 0u, 0u, 0u);
        }
    }
}

// The table path of the frame chain as a kernel of its own: PointCloud2 records (+ the float disparity plane) from the
// per-disparity table, eight rows per warp, with the loads hoisted -- the disparities and colour bytes of all eight rows are
// requested up front and the table entry of row r + 1 is in flight while row r is computed and stored (the generic kernel
// waited for two dependent loads in every row: 49 % of its issue slots were busy).  Same operations per value as reproject_pack_kernel.
constexpr int PL_RPW = 8;                        // consecutive image rows per warp
constexpr uint32_t LUT_MISS_Z = 0x7f800002u;     // pipeline marker: disparity outside the table, take the arithmetic

template <int STDQ>
__device__ __forceinline__ float stdq_coord(double a, double rw, double q32, double q33, float dfv)
{
    const double q = __dmul_rn(a, rw);
    const unsigned lo = (unsigned)__double2loint(q) & 0x1FFFFFFFu, hi = (unsigned)__double2hiint(q) & 0x7FFFFFFFu;
    const bool safe = hi > 0x39B00000u && hi < 0x46300000u && (lo - 0x0FFFFFF0u) > 0x20u;
    return __double2float_rn(safe ? q : __ddiv_rn(a, stdq_w<STDQ>(q32, q33, (double)dfv)));
}

template <int STDQ, int CH>
__global__ void __launch_bounds__(256, 4) pack_lut_kernel(const int16_t* __restrict__ d16, int W, int H, double cxd,
                                                       const double* __restrict__ Q, const uint8_t* __restrict__ color,
                                                       uint8_t* __restrict__ pc2, size_t d_stride, size_t color_stride, const FrameDst fd,
                                                       int dmin, float* __restrict__ df, const FrameDst fdf,
                                                       const uint8_t* const* __restrict__ color_tab,
                                                       const uint4* __restrict__ lut, int lut_n)
{
    const int lane = threadIdx.x & 31;
    const int x0 = blockIdx.x * 32, x = x0 + lane;
    const int yw = (blockIdx.y * 8 + (threadIdx.x >> 5)) * PL_RPW;      // first row of this warp
    {
        const int f = blockIdx.z;          // frame of the batch
        d16 = (const int16_t*)((const uint8_t*)d16 + f * d_stride);
        if (color_tab) color = color_tab[f];
        else if (color) color += f * color_stride;
        if (fd.use_list) pc2 = (uint8_t*)fd.list.p[f];
        else pc2 += f * fd.stride;
        if (fdf.use_list) df = (float*)fdf.list.p[f];
        else if (df) df = (float*)((uint8_t*)df + f * fdf.stride);
    }
    if (yw >= H) return;
    const double ax = (double)__double2float_rn(__dadd_rn(__dmul_rn(__ldg(Q + 0), (double)x), __ldg(Q + 3)));
    const double q11 = __ldg(Q + 5), q13 = __ldg(Q + 7), q32 = __ldg(Q + 14), q33 = __ldg(Q + 15);
    const bool xin = x < W;
    const bool full_row = x0 + 32 <= W;
    const int nrows = min(PL_RPW, H - yw);
    auto load_dv = [&](int y) { return xin ? (int)__ldg(d16 + (size_t)y * W + x) : dmin; };
    auto load_bgr = [&](int y) -> uint32_t {
        if (!xin || !color) return 0u;
        const size_t i = (size_t)y * W + x;
        if (CH == 3) return (uint32_t)__ldg(color + i * 3) | ((uint32_t)__ldg(color + i * 3 + 1) << 8) | ((uint32_t)__ldg(color + i * 3 + 2) << 16);
        const uint32_t g = __ldg(color + i);
        return g | (g << 8) | (g << 16);
    };
    auto load_entry = [&](int dv) {
        const unsigned li = (unsigned)(dv - dmin);
        return li < (unsigned)lut_n ? __ldg(lut + li) : make_uint4(0u, 0u, LUT_MISS_Z, 0u);
    };
    // all disparities and colour bytes of the warp's rows are requested first; the table entry of row r + 1 is requested
    // before row r is computed and stored
    int dvs[PL_RPW];
    uint32_t bgrs[PL_RPW];
#pragma unroll
    for (int rr = 0; rr < PL_RPW; ++rr) {
        dvs[rr] = rr < nrows ? load_dv(yw + rr) : dmin;
        bgrs[rr] = rr < nrows ? load_bgr(yw + rr) : 0u;
    }
    uint4 eB = load_entry(dvs[0]);
#pragma unroll
    for (int rr = 0; rr < PL_RPW; ++rr) {
        if (rr >= nrows) break;                              // warp-uniform
        const int y = yw + rr;
        const uint4 e = eB;
        const int dv = dvs[rr];
        const uint32_t bgr = bgrs[rr];
        if (rr + 1 < PL_RPW && rr + 1 < nrows) eB = load_entry(dvs[rr + 1]);
        uint32_t ux = 0x7fc00000u, uy = 0x7fc00000u, uz = 0x7fc00000u;
        const size_t i = xin ? (size_t)y * W + x : 0;
        if (e.z == LUT_MISS_Z) {
            // outside the table (cannot happen for planes this library's matcher wrote): the arithmetic of reproject_pack_kernel
            const float dfv = disp_to_float(dv, cxd);
            if (df && xin) df[i] = dfv;
            if (xin && dv != dmin) {
                const double d = (double)dfv, minDisp = (double)disp_to_float(dmin, cxd);
                const double az = (double)__double2float_rn(__dadd_rn(0.0, __ldg(Q + 11)));
                const double ay = (double)__double2float_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(q11, (double)y)), q13));
                const double w = stdq_w<STDQ>(q32, q33, d);
                const double rw = __drcp_rn(w);
                const float p0 = div_to_float(ax, w, rw), p1 = div_to_float(ay, w, rw);
                float p2 = div_to_float(az, w, rw);
                if (fabs(__dadd_rn(d, -minDisp)) <= (double)1.1920928955078125e-07f) p2 = 10000.0f;
                if ((p2 != 10000.0f) && !isinf(p2)) { ux = __float_as_uint(p0); uy = __float_as_uint(p1); uz = __float_as_uint(p2); }
            }
        } else {
            if (df && xin) df[i] = __uint_as_float(e.w);
            if (xin && e.z != LUT_INVALID_Z) {
                const double rw = __hiloint2double((int)e.y, (int)e.x);
                const double ay = (double)__double2float_rn(__dadd_rn(__dadd_rn(0.0, __dmul_rn(q11, (double)y)), q13));
                const float dfv = __uint_as_float(e.w);
                ux = __float_as_uint(stdq_coord<STDQ>(ax, rw, q32, q33, dfv));
                uy = __float_as_uint(stdq_coord<STDQ>(ay, rw, q32, q33, dfv));
                uz = e.z;
            }
        }
        // records: see reproject_pack_kernel (two 512-byte store instructions per warp row)
        if (full_row) {
            uint4* row = (uint4*)(pc2 + ((size_t)y * W + x0) * 32);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int src = (lane >> 1) + 16 * k;
                const uint32_t sx = __shfl_sync(0xffffffffu, ux, src), sy = __shfl_sync(0xffffffffu, uy, src);
                const uint32_t sz = __shfl_sync(0xffffffffu, uz, src), sc = __shfl_sync(0xffffffffu, bgr, src);
                row[32 * k + lane] = (lane & 1) ? make_uint4(sc, 0u, 0u, 0u) : make_uint4(sx, sy, sz, 0u);
            }
        } else if (xin) {
            uint4* o = (uint4*)(pc2 + i * 32);
            o[0] = make_uint4(ux, uy, uz, 0u);
            o[1] = make_uint4(bgr, 0u, 0u, 0u);
        }
    }
}

// cv::cuda::drawColorDisp (opencv_contrib cudastereo, util.cu cvtPixel) on the integer disparity d = clamp(d16 >> 4, 0, 255),
// the u8 value the reference's matcher plane holds (invalid = 0 -> hue 240 = blue).  Reference call site:
// computeDisparityImage, src/GPUStereoProcessor.cpp:323-330.  "next" row 1 of the scope table; the upstream kernel is
// not in /root/reference and cv2 has no CUDA modules, so this is a restatement without a golden (parity unpinned).
__global__ void __launch_bounds__(256) disparity_color_kernel(const int16_t* __restrict__ d16, uint8_t* __restrict__ bgra,
                                                              int n, int nd)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = min(max((int)d16[i] >> 4, 0), 255);
    const unsigned H = (unsigned)(((nd - d) * 240) / nd);   // int division, then unsigned, like upstream
    const unsigned hi = (H / 60u) % 6u;
    const float f = __fsub_rn(__fdiv_rn((float)H, 60.f), (float)(H / 60u));
    const float V = 1.f, p = 0.f, q = __fsub_rn(1.f, f), t = f;     // S = V = 1: p = 0, q = 1 - f, t = 1 - (1 - f) = f
    float x, y, z;                                                   // x = blue, y = green, z = red
    switch (hi) {
        case 0: x = p; y = t; z = V; break;
        case 1: x = p; y = V; z = q; break;
        case 2: x = t; y = V; z = p; break;
        case 3: x = V; y = q; z = p; break;
        case 4: x = V; y = p; z = t; break;
        default: x = q; y = p; z = V; break;
    }
    const unsigned bb = (unsigned)__fmul_rn(fmaxf(0.f, fminf(x, 1.f)), 255.f);
    const unsigned gg = (unsigned)__fmul_rn(fmaxf(0.f, fminf(y, 1.f)), 255.f);
    const unsigned rr = (unsigned)__fmul_rn(fmaxf(0.f, fminf(z, 1.f)), 255.f);
    ((uint32_t*)bgra)[i] = bb | (gg << 8) | (rr << 16) | (255u << 24);
}

static FrameDst frame_dst(size_t stride, const PtrList* list)
{
    FrameDst fd;
    fd.stride = stride;
    fd.use_list = list ? 1 : 0;
    if (list) fd.list = *list;
    else for (int i = 0; i < MAX_BATCH; ++i) fd.list.p[i] = nullptr;
    return fd;
}

int launch_disparity_to_float(const int16_t* d16, float* df, int n, double cxd, int* min_d16, cudaStream_t st, int nf,
                              size_t d_stride, size_t df_stride, const PtrList* df_list)
{
    cudaMemsetAsync(min_d16, 0x7f, nf * sizeof(int), st);     // 0x7f7f7f7f: larger than any int16
    disparity_to_float_kernel<<<dim3((n + 2047) / 2048, nf), 256, 0, st>>>(d16, df, n, cxd, min_d16, d_stride, frame_dst(df_stride, df_list));
    return 1;
}

int launch_reproject_pack(const int16_t* d16, int W, int H, double cxd, const double* Q, unsigned qmask, const int* min_d16,
                          const uint8_t* color, int ch, float* xyz, uint8_t* pc2, cudaStream_t st, int nf, size_t d_stride,
                          size_t color_stride, size_t xyz_stride, size_t pc2_stride, const PtrList* pc2_list, const ReprojectExtras* extra)
{
    dim3 g((W + 31) / 32, (H + RP_ROWS * RP_RPW - 1) / (RP_ROWS * RP_RPW), nf);
    if (pc2_list && !pc2) pc2 = (uint8_t*)pc2_list->p[0];     // the kernel tests pc2 for "records wanted"
    ReprojectExtras ex;
    if (extra) ex = *extra;
    const uint4* lut = (const uint4*)ex.lut;
    const int stdq = qmask == QMASK_STEREO ? 1 : (qmask == QMASK_STEREO0 ? 2 : 0);
    const bool use_lut = stdq && lut && ex.lut_n > 0 && !xyz && !min_d16;
#define B200S_PACK_ARGS d16, W, H, cxd, Q, min_d16, color, ch, xyz, pc2, qmask, d_stride, color_stride, xyz_stride, \
                        frame_dst(pc2_stride, pc2_list), ex.dmin_const, ex.df, frame_dst(ex.df_stride, ex.df_list), ex.color_tab, \
                        use_lut ? lut : nullptr, use_lut ? ex.lut_n : 0
    static const int lean = getenv("B200S_PACK_LEAN") ? atoi(getenv("B200S_PACK_LEAN")) : 1;
    if (use_lut && lean && pc2 && (ch == 1 || ch == 3)) {
        dim3 g2((W + 31) / 32, (H + 8 * PL_RPW - 1) / (8 * PL_RPW), nf);
#define B200S_LEAN_ARGS d16, W, H, cxd, Q, color, pc2, d_stride, color_stride, frame_dst(pc2_stride, pc2_list), ex.dmin_const, ex.df, \
                        frame_dst(ex.df_stride, ex.df_list), ex.color_tab, lut, ex.lut_n
        if (stdq == 1 && ch == 3) pack_lut_kernel<1, 3><<<g2, 256, 0, st>>>(B200S_LEAN_ARGS);
        else if (stdq == 1) pack_lut_kernel<1, 1><<<g2, 256, 0, st>>>(B200S_LEAN_ARGS);
        else if (ch == 3) pack_lut_kernel<2, 3><<<g2, 256, 0, st>>>(B200S_LEAN_ARGS);
        else pack_lut_kernel<2, 1><<<g2, 256, 0, st>>>(B200S_LEAN_ARGS);
#undef B200S_LEAN_ARGS
        return 1;
    }
    if (stdq == 1 && use_lut) reproject_pack_kernel<1, true><<<g, 32 * RP_ROWS, 0, st>>>(B200S_PACK_ARGS);
    else if (stdq == 2 && use_lut) reproject_pack_kernel<2, true><<<g, 32 * RP_ROWS, 0, st>>>(B200S_PACK_ARGS);
    else if (stdq == 1) reproject_pack_kernel<1, false><<<g, 32 * RP_ROWS, 0, st>>>(B200S_PACK_ARGS);
    else if (stdq == 2) reproject_pack_kernel<2, false><<<g, 32 * RP_ROWS, 0, st>>>(B200S_PACK_ARGS);
    else reproject_pack_kernel<0, false><<<g, 32 * RP_ROWS, 0, st>>>(B200S_PACK_ARGS);
#undef B200S_PACK_ARGS
    return 1;
}

int launch_reproject_lut(void* lut, int n, int dmin, double cxd, const double* Q, unsigned qmask, cudaStream_t st)
{
    if (n <= 0 || !lut) return 0;
    if (qmask == QMASK_STEREO) reproject_lut_kernel<1><<<(n + 255) / 256, 256, 0, st>>>((uint4*)lut, n, dmin, cxd, Q);
    else if (qmask == QMASK_STEREO0) reproject_lut_kernel<2><<<(n + 255) / 256, 256, 0, st>>>((uint4*)lut, n, dmin, cxd, Q);
    else return 0;
    return 1;
}

int launch_disparity_color(const int16_t* d16, uint8_t* bgra, int n, int nd, cudaStream_t st)
{
    disparity_color_kernel<<<(n + 255) / 256, 256, 0, st>>>(d16, bgra, n, nd);
    return 1;
}

}  // namespace b200s
