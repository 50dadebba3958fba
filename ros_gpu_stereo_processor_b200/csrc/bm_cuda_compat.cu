// "cuda-compat" matcher: reproduces the bytes of cv::cuda::StereoBM, i.e. what the reference's GPU path computes at
// src/GPUStereoProcessor.cpp:283 (block_matcher_gpu_->compute) -- SURVEY.md A.6 / 8(f)-3.  Known answer: the reference's
// own test_data/aloe-disp.png (createStereoBM(128, 19)).
//
//   optional prefilter   full 3x3 Sobel-x on clamp-addressed pixels, min(clip(v, -cap, cap) + cap, 255)
//   matching             SSD over wsz^2, integer disparity candidates 0..nd-1 scanned ascending in groups of 8; inside a
//                        group the HIGHEST index holding the minimum wins, across groups strict '<' (the earlier group
//                        wins); computed for x in [nd + r, W - r), y in [r, H - r); CV_8UC1, 0 elsewhere
//   textureness filter   window sum of |Sobel-x| below avergeTexThreshold * wsz^2 -> 0 (exact integer arithmetic; upstream
//                        sums normalised floats: parity unpinned at the rounding boundary)
//
// Written from scratch for sm_100a: a block marches down a band of rows for 128 output columns; per group of 8 candidates
// every thread keeps the vertical column sums of its window-left column in registers (sliding down the rows), the block
// shares them through shared memory and every thread adds the 2r+1 columns of its window (as sums of 4 adjacent columns
// plus up to 3 single ones) with 128-bit loads.  Running
// minima per pixel live in shared memory for the whole band, so the output is written once.
#include "kernels.h"

#include <algorithm>
#include <cstdlib>

namespace b200s {

namespace {

constexpr int CC_BW = 128;    // output columns per block
constexpr int CC_RPT = 32;    // output rows per block (upper bound of the band)

__global__ void __launch_bounds__(256) cc_sobel_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int W, int H, int cap)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const int xm = max(x - 1, 0), xp = min(x + 1, W - 1), ym = max(y - 1, 0), yp = min(y + 1, H - 1);
    const uint8_t *r0 = src + (size_t)ym * W, *r1 = src + (size_t)y * W, *r2 = src + (size_t)yp * W;
    const int conv = -(int)r0[xm] + (int)r0[xp] - 2 * (int)r1[xm] + 2 * (int)r1[xp] - (int)r2[xm] + (int)r2[xp];
    dst[(size_t)y * W + x] = (uint8_t)min(min(max(conv, -cap), cap) + cap, 255);
}

struct CcParams {
    const uint8_t* L;
    const uint8_t* R;
    uint8_t* disp;
    int W, H, nd, r, rpt;
};

__device__ __forceinline__ void cc_row_ssd(const uint8_t* __restrict__ lrow, const uint8_t* __restrict__ rrow, int x, int d0, int sign, uint32_t (&col)[8])
{
    const int l = (int)__ldg(lrow + x);
    const uint8_t* rp = rrow + x - d0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int df = l - (int)__ldg(rp - j);
        col[j] += (uint32_t)(sign * df * df);
    }
}

__global__ void __launch_bounds__(CC_BW) cc_ssd_kernel(const CcParams P)
{
    extern __shared__ __align__(16) uint32_t cc_sm[];
    const int r = P.r, ncol = CC_BW + 2 * r;
    uint4* colS = (uint4*)cc_sm;                          // [ncol][2]  column sums, 8 candidates per column
    uint4* qS = colS + 2 * ncol;                          // [ncol][2]  sums of 4 adjacent columns
    uint32_t* bestS = cc_sm + 16 * ncol;                  // [rpt][CC_BW]
    uint8_t* bestD = (uint8_t*)(bestS + P.rpt * CC_BW);   // [rpt][CC_BW]
    const int t = threadIdx.x;
    const int X = P.nd + r + blockIdx.x * CC_BW + t;      // output column of this thread
    const int xt = X - r;                                 // left column of its window
    const int Y0 = r + blockIdx.y * P.rpt;
    const int rows = min(P.rpt, P.H - r - Y0);
    const bool extra = t < 2 * r;                         // these threads also own column xt + CC_BW
    const bool own_ok = xt < P.W, extra_ok = extra && xt + CC_BW < P.W;
    const bool out_ok = X < P.W - r;
    for (int i = t; i < P.rpt * CC_BW; i += CC_BW) { bestS[i] = 0xFFFFFFFFu; bestD[i] = 0; }
    __syncthreads();
    for (int g = 0; g < P.nd; g += 8) {
        uint32_t col[8], colx[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) col[j] = colx[j] = 0;
        for (int yy = Y0 - r; yy <= Y0 + r; ++yy) {
            const uint8_t *lr = P.L + (size_t)yy * P.W, *rr = P.R + (size_t)yy * P.W;
            if (own_ok) cc_row_ssd(lr, rr, xt, g, 1, col);
            if (extra_ok) cc_row_ssd(lr, rr, xt + CC_BW, g, 1, colx);
        }
        for (int row = 0; row < rows; ++row) {
            if (row > 0) {
                const int ya = Y0 + row + r, yl = Y0 + row - r - 1;
                const uint8_t *la = P.L + (size_t)ya * P.W, *ra = P.R + (size_t)ya * P.W;
                const uint8_t *ll = P.L + (size_t)yl * P.W, *rl = P.R + (size_t)yl * P.W;
                if (own_ok) { cc_row_ssd(la, ra, xt, g, 1, col); cc_row_ssd(ll, rl, xt, g, -1, col); }
                if (extra_ok) { cc_row_ssd(la, ra, xt + CC_BW, g, 1, colx); cc_row_ssd(ll, rl, xt + CC_BW, g, -1, colx); }
            }
            colS[2 * t] = make_uint4(col[0], col[1], col[2], col[3]);
            colS[2 * t + 1] = make_uint4(col[4], col[5], col[6], col[7]);
            if (extra) {
                colS[2 * (t + CC_BW)] = make_uint4(colx[0], colx[1], colx[2], colx[3]);
                colS[2 * (t + CC_BW) + 1] = make_uint4(colx[4], colx[5], colx[6], colx[7]);
            }
            __syncthreads();
            // two-level window sum: quad sums q[c] = col[c] + .. + col[c+3] first (threads that own an extra column also
            // form q[CC_BW + t]), then the 2r+1 = 4a + b columns of a window are a quads and b single columns
            {
                auto quad = [&](int c) {
                    uint4 a = colS[2 * c], b = colS[2 * c + 1];
#pragma unroll
                    for (int i = 1; i < 4; ++i) {
                        const uint4 a2 = colS[2 * (c + i)], b2 = colS[2 * (c + i) + 1];
                        a.x += a2.x; a.y += a2.y; a.z += a2.z; a.w += a2.w;
                        b.x += b2.x; b.y += b2.y; b.z += b2.z; b.w += b2.w;
                    }
                    qS[2 * c] = a;
                    qS[2 * c + 1] = b;
                };
                quad(t);
                if (t + CC_BW + 3 < ncol) quad(t + CC_BW);
            }
            __syncthreads();
            if (out_ok) {
                uint32_t s[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) s[j] = 0;
                const int na = (2 * r + 1) >> 2, nb = (2 * r + 1) & 3;
                for (int i = 0; i < na; ++i) {
                    const uint4 a = qS[2 * (t + 4 * i)], b = qS[2 * (t + 4 * i) + 1];
                    s[0] += a.x; s[1] += a.y; s[2] += a.z; s[3] += a.w;
                    s[4] += b.x; s[5] += b.y; s[6] += b.z; s[7] += b.w;
                }
                for (int i = 0; i < nb; ++i) {
                    const uint4 a = colS[2 * (t + 4 * na + i)], b = colS[2 * (t + 4 * na + i) + 1];
                    s[0] += a.x; s[1] += a.y; s[2] += a.z; s[3] += a.w;
                    s[4] += b.x; s[5] += b.y; s[6] += b.z; s[7] += b.w;
                }
                uint32_t m = s[0];
#pragma unroll
                for (int j = 1; j < 8; ++j) m = min(m, s[j]);
                int bi = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (s[j] == m) bi = j;                  // the highest index holding the minimum
                if (m < bestS[row * CC_BW + t]) {
                    bestS[row * CC_BW + t] = m;
                    bestD[row * CC_BW + t] = (uint8_t)(g + bi);
                }
            }
            __syncthreads();
        }
    }
    if (out_ok)
        for (int row = 0; row < rows; ++row) P.disp[(size_t)(Y0 + row) * P.W + X] = bestD[row * CC_BW + t];
}

// |Sobel-x| of the (clamp-addressed) image at possibly out-of-image coordinates
__device__ __forceinline__ int cc_abs_sobel(const uint8_t* __restrict__ img, int W, int H, int x, int y)
{
    const int xm = min(max(x - 1, 0), W - 1), xp = min(max(x + 1, 0), W - 1);
    const int y0 = min(max(y - 1, 0), H - 1), y1 = min(max(y, 0), H - 1), y2 = min(max(y + 1, 0), H - 1);
    const uint8_t *r0 = img + (size_t)y0 * W, *r1 = img + (size_t)y1 * W, *r2 = img + (size_t)y2 * W;
    return abs(-(int)__ldg(r0 + xm) + (int)__ldg(r0 + xp) - 2 * (int)__ldg(r1 + xm) + 2 * (int)__ldg(r1 + xp) - (int)__ldg(r2 + xm) + (int)__ldg(r2 + xp));
}

// one block = 32 x 8 pixels: |Sobel| tile with a halo of r in shared memory, horizontal window sums, vertical sums
__global__ void __launch_bounds__(256) cc_texture_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ disp, int W, int H, int r, int thr)
{
    extern __shared__ int tx_sm[];
    const int TWd = 32 + 2 * r, THt = 8 + 2 * r;
    int* sob = tx_sm;                 // [THt][TWd]
    int* hs = tx_sm + THt * TWd;      // [THt][32]
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
    for (int i = threadIdx.x; i < THt * TWd; i += 256) {
        const int ty = i / TWd, tx = i - ty * TWd;
        sob[i] = cc_abs_sobel(img, W, H, x0 + tx - r, y0 + ty - r);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < THt * 32; i += 256) {
        const int ty = i >> 5, tx = i & 31;
        int s = 0;
        for (int c = 0; c <= 2 * r; ++c) s += sob[ty * TWd + tx + c];
        hs[i] = s;
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int x = x0 + tx, y = y0 + ty;
    if (x >= W || y >= H) return;
    int s = 0;
    for (int c = 0; c <= 2 * r; ++c) s += hs[(ty + c) * 32 + tx];
    if (s < thr) disp[(size_t)y * W + x] = 0;
}

__global__ void __launch_bounds__(256) cc_u8_to_s16_kernel(const uint8_t* __restrict__ a, int16_t* __restrict__ b, size_t n)
{
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) b[i] = (int16_t)a[i];
}

__global__ void __launch_bounds__(256) cc_s16_to_u8_kernel(const int16_t* __restrict__ a, uint8_t* __restrict__ b, size_t n)
{
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) b[i] = (uint8_t)a[i];
}

}  // namespace

// L, R: tightly packed u8 planes; tmpL/tmpR: W*H scratch planes (only used with xsobel); disp: W*H u8
int launch_cuda_compat_bm(const uint8_t* L, const uint8_t* R, uint8_t* tmpL, uint8_t* tmpR, int W, int H, int nd, int wsz,
                          bool xsobel, int cap, int tex_threshold, uint8_t* disp, cudaStream_t st)
{
    int launches = 0;
    const int r = wsz / 2;
    cudaMemsetAsync(disp, 0, (size_t)W * H, st);
    if (xsobel) {
        dim3 g((W + 31) / 32, (H + 7) / 8);
        cc_sobel_kernel<<<g, 256, 0, st>>>(L, tmpL, W, H, cap);
        cc_sobel_kernel<<<g, 256, 0, st>>>(R, tmpR, W, H, cap);
        launches += 2;
        L = tmpL;
        R = tmpR;
    }
    const int ncx = W - r - (nd + r), ncy = H - 2 * r;
    if (ncx > 0 && ncy > 0) {
        CcParams P{L, R, disp, W, H, nd, r, CC_RPT};
        const int gx = (ncx + CC_BW - 1) / CC_BW;
        // shorter bands until the grid covers the SMs about four times (measured: 1080p 2.06 ms at 32 rows, 1.56 ms at 16;
        // each block re-reads 2r rows per group, so 8 rows is the floor)
        static const int rpt_env = getenv("B200S_CC_RPT") ? atoi(getenv("B200S_CC_RPT")) : 0;
        while (P.rpt > 8 && gx * ((ncy + P.rpt - 1) / P.rpt) < 4 * 148) P.rpt /= 2;
        if (rpt_env > 0) P.rpt = rpt_env;
        const size_t smem = (size_t)16 * (CC_BW + 2 * r) * 4 + (size_t)P.rpt * CC_BW * 5;
        cc_ssd_kernel<<<dim3(gx, (ncy + P.rpt - 1) / P.rpt), CC_BW, smem, st>>>(P);
        ++launches;
    }
    if (tex_threshold > 0) {
        const size_t smem = ((size_t)(8 + 2 * r) * (32 + 2 * r) + (size_t)(8 + 2 * r) * 32) * 4;
        cc_texture_kernel<<<dim3((W + 31) / 32, (H + 7) / 8), 256, smem, st>>>(L, disp, W, H, r, tex_threshold * wsz * wsz);
        ++launches;
    }
    return launches;
}

int launch_u8_to_s16(const uint8_t* a, int16_t* b, size_t n, cudaStream_t st)
{
    cc_u8_to_s16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, b, n);
    return 1;
}

int launch_s16_to_u8(const int16_t* a, uint8_t* b, size_t n, cudaStream_t st)
{
    cc_s16_to_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, b, n);
    return 1;
}

}  // namespace b200s
