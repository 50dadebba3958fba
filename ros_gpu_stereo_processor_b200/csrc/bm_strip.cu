// Border strips of cv::StereoBM for sm_100a.
//
// With disp12MaxDiff >= 0 (the state the reference's matcher is in out of the box, src/GPUStereoProcessor.cpp:23) OpenCV
// also needs the disparities of the r columns left and right of the valid ROI, where its window sums run against the
// clamped image border: they feed the left-right check (cv::validateDisparity) before the ROI mask
// removes them.  The warp-specialised matcher (bm_vh.cu) only covers the clamp-free rectangle; the generic kernels
// (bm_sad.cu) compute the two strips exactly but through a global cost volume with two latency-bound launches (56 us at
// 1080p).  This kernel does both strips of a frame (of a whole batch) in one launch and keeps everything on chip:
//
//   block  = one strip x a chunk of output rows x a frame; the chunk's rows of the left strip window and of the matching right
//            segment are staged in shared memory once (the border clamps are applied while staging, so the inner loops are
//            clamp-free);
//   C warps = one thread per disparity: vertical column sums of its 3r window columns in registers, sliding down the rows;
//            per row the r window sums go to one of two shared-memory buffers;
//   winner warps = one warp per strip pixel: argmin (lowest index on ties), uniqueness, texture (sliding column sums, one
//            lane per window column), sub-pixel -- the same selection as gen_winner_warp_body.
// The two roles hand the buffers over through named barriers (bar.arrive / bar.sync), so the selection of row o overlaps
// the column-sum update of row o + 1.  Integer arithmetic in int32 throughout: bit-exact for any cap / window.
#include "kernels.h"
#include "bm_common.cuh"

#include <algorithm>
#include <climits>

namespace b200s {

struct StripParams {
    const uint8_t* Lp;
    const uint8_t* Rp;
    size_t pitch, pre_stride, disp_stride;
    int16_t* disp;
    int16_t* cost;
    int W, H, nd, minD, cap, texThr, uniq, lofs;
    int xa[2];       // first strip column (x = X - lofs) of the left / right strip; each strip is R columns wide
    int ya, yb;      // output rows
    int rch;         // output rows per block
};

constexpr int STRIP_MAXC = 256;       // one C thread per disparity: nd <= 256
constexpr int STRIP_MAXT = STRIP_MAXC + 32 * 10;

namespace strip {
enum { BAR_FULL = 1, BAR_EMPTY = 3 };       // named barriers FULL + buffer, EMPTY + buffer
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
}  // namespace strip

template <int R>
__global__ void __launch_bounds__(STRIP_MAXT, 2) bm_strip_fused_kernel(const StripParams P)
{
    using namespace strip;
    constexpr int B = 2 * R + 1;
    constexpr int WC = 3 * R;                    // window columns of a strip: R output columns + 2R halo
    constexpr int LW = (WC + 3) & ~3;            // staged left bytes per row
    extern __shared__ __align__(16) uint8_t smem[];
    const int nd = P.nd;
    const int RW = (WC + nd + 3) & ~3;           // staged right bytes per row
    const int y0 = P.ya + blockIdx.x * P.rch, y1 = min(y0 + P.rch, P.yb);
    const int nrows = y1 - y0, NR = nrows + 2 * R;
    const int side = blockIdx.y;
    const int xa = side ? P.xa[1] : P.xa[0];
    const uint8_t* __restrict__ Lp = P.Lp + blockIdx.z * P.pre_stride;
    const uint8_t* __restrict__ Rp = P.Rp + blockIdx.z * P.pre_stride;
    int16_t* __restrict__ disp = (int16_t*)((uint8_t*)P.disp + blockIdx.z * P.disp_stride);
    int16_t* __restrict__ cost = P.cost ? (int16_t*)((uint8_t*)P.cost + blockIdx.z * P.disp_stride) : nullptr;
    uint8_t* Ls = smem;                                           // [NRmax][LW]
    uint8_t* Rs = smem + (size_t)(P.rch + 2 * R) * LW;            // [NRmax][RW]
    int* Bo = (int*)(smem + (((size_t)(P.rch + 2 * R) * (LW + RW) + 15) & ~(size_t)15));   // [LW]: right-segment offset of a window column
    int* Ssm = Bo + LW;                                           // [2][R][nd]
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int NTC = nthreads - 32 * R;                            // C threads (nd rounded up to whole warps)
    // right-image column of window column `col` at disparity index 0 (rofs == 0 here)
    const int bmin = min(max(xa - R, 0), P.W - nd);
    // ---- stage: clamps applied here ----
    for (int i = tid; i < NR * LW; i += nthreads) {
        const int row = i / LW, col = i - row * LW;
        const int lc = min(max(xa + col - R, -P.lofs), P.W - 1 - P.lofs) + P.lofs;
        Ls[i] = col < WC ? __ldg(Lp + (size_t)(y0 - R + row) * P.pitch + lc) : 0;
    }
    const int rbytes = min(max(xa + WC - 1 - R, 0), P.W - nd) - bmin + nd;      // distinct right bytes a row needs
    for (int i = tid; i < NR * RW; i += nthreads) {
        const int row = i / RW, j = i - row * RW;
        Rs[i] = j < rbytes ? __ldg(Rp + (size_t)(y0 - R + row) * P.pitch + bmin + j) : 0;
    }
    if (tid < WC) Bo[tid] = min(max(xa + tid - R, 0), P.W - nd) - bmin;
    __syncthreads();
    if (tid < NTC) {
        // =============================== C role: one disparity per thread ====================================
        const int k = tid;
        const bool kact = k < nd;
        const uint8_t* rk = Rs + (kact ? k : 0);
        int C[WC];
#pragma unroll
        for (int col = 0; col < WC; ++col) C[col] = 0;
        for (int i = 0; i < B; ++i) {
            const uint8_t* lr = Ls + i * LW;
            const uint8_t* rr = rk + i * RW;
#pragma unroll
            for (int col = 0; col < WC; ++col) C[col] += abs((int)lr[col] - (int)rr[Bo[col]]);
        }
        for (int o = 0; o < nrows; ++o) {
            const int buf = o & 1;
            if (o >= 2) bar_sync(BAR_EMPTY + buf, nthreads);      // the winner warps are done with row o - 2
            if (kact) {
                int* S = Ssm + buf * R * nd + k;
                int s = 0;
#pragma unroll
                for (int col = 0; col < B; ++col) s += C[col];
#pragma unroll
                for (int c = 0; c < R; ++c) {
                    S[c * nd] = s;
                    if (c + 1 < R) s += C[c + B] - C[c];
                }
            }
            __threadfence_block();
            bar_arrive(BAR_FULL + buf, nthreads);
            if (o + 1 < nrows) {
                const uint8_t* ln = Ls + (o + B) * LW;
                const uint8_t* rn = rk + (o + B) * RW;
                const uint8_t* lo = Ls + o * LW;
                const uint8_t* ro = rk + o * RW;
#pragma unroll
                for (int col = 0; col < WC; ++col) {
                    const int bo = Bo[col];
                    C[col] += abs((int)ln[col] - (int)rn[bo]) - abs((int)lo[col] - (int)ro[bo]);
                }
            }
        }
    } else {
        // =============================== winner role: one warp per strip pixel ================================
        const int c = (tid - NTC) >> 5, lane = tid & 31;
        const int16_t FILTERED = (int16_t)((P.minD - 1) * 16);
        const int X = xa + c + P.lofs;
        const bool inside = X >= 0 && X < P.W;
        // texture: lane l < B slides the column sum of window column c + l
        const bool tact = lane < B;
        const uint8_t* lt = Ls + c + (tact ? lane : 0);
        int tcol = 0;
        for (int i = 0; i < B; ++i) tcol += abs((int)lt[i * LW] - P.cap);
        if (!tact) tcol = 0;
        constexpr int NV = STRIP_MAXC / 32;
        for (int o = 0; o < nrows; ++o) {
            const int buf = o & 1;
            const int tsum = __reduce_add_sync(0xffffffffu, tcol);
            if (tact && o + 1 < nrows) tcol += abs((int)lt[(o + B) * LW] - P.cap) - abs((int)lt[o * LW] - P.cap);
            bar_sync(BAR_FULL + buf, nthreads);
            const int* S = Ssm + (buf * R + c) * nd;
            int v[NV];
            int best = INT_MAX, bk = nd;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int kk = lane + 32 * i;
                v[i] = kk < nd ? S[kk] : INT_MAX;
                if (v[i] < best) { best = v[i]; bk = kk; }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const int ov = __shfl_xor_sync(0xffffffffu, best, d), ok2 = __shfl_xor_sync(0xffffffffu, bk, d);
                if (ov < best || (ov == best && ok2 < bk)) { best = ov; bk = ok2; }
            }
            const int minsad = best, mind = bk;
            const int p = S[mind + 1 < nd ? mind + 1 : nd - 2], n = S[mind > 0 ? mind - 1 : 1];
            if (o + 2 < nrows) {
                __threadfence_block();
                bar_arrive(BAR_EMPTY + buf, nthreads);            // everything this warp needs of the buffer is in registers
            }
            bool ok = tsum >= P.texThr;
            if (ok && P.uniq > 0) {
                const int thr = minsad + (minsad * P.uniq / 100);
                bool hit = false;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const int kk = lane + 32 * i;
                    if ((kk < mind - 1 || kk > mind + 1) && v[i] <= thr) hit = true;     // lanes past nd hold INT_MAX
                }
                if (__any_sync(0xffffffffu, hit)) ok = false;
            }
            if (lane == 0 && inside) {
                const int y = y0 + o;
                int16_t out = FILTERED;
                if (ok) {
                    out = subpixel_disp(minsad, mind, p, n, nd, P.minD);
                    if (cost) cost[(size_t)y * P.W + X] = (int16_t)minsad;
                }
                disp[(size_t)y * P.W + X] = out;
            }
        }
    }
}

template <int R>
static cudaError_t launch_strip(const StripParams& P, dim3 grid, int nt, size_t smem, cudaStream_t st)
{
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(bm_strip_fused_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    bm_strip_fused_kernel<R><<<grid, nt, smem, st>>>(P);
    return cudaGetLastError();
}

// Both r-wide strips [xa0, xa0 + r) and [xa1, xa1 + r) (x = X - lofs) of rows [ya, yb) for nf frames.
// returns 1 when launched, 0 when the configuration is not handled (nd > 256, r outside 2..10), < 0 on CUDA errors
int launch_bm_strips(const uint8_t* Lp, const uint8_t* Rp, size_t pitch, int W, int H, const BMConfig& cfg, int r, int lofs,
                     int xa0, int xa1, int ya, int yb, int16_t* disp, int16_t* cost, cudaStream_t st, int nf, size_t pre_stride,
                     size_t disp_stride)
{
    if (cfg.nd > STRIP_MAXC || r < 2 || r > 10 || yb <= ya || W - cfg.nd < 0) return 0;
    StripParams P;
    P.Lp = Lp; P.Rp = Rp; P.pitch = pitch; P.pre_stride = pre_stride; P.disp_stride = disp_stride;
    P.disp = disp; P.cost = cost;
    P.W = W; P.H = H; P.nd = cfg.nd; P.minD = cfg.minD; P.cap = cfg.cap; P.texThr = cfg.textureThreshold; P.uniq = cfg.uniquenessRatio;
    P.lofs = lofs; P.xa[0] = xa0; P.xa[1] = xa1; P.ya = ya; P.yb = yb;
    // rows per block: about two waves of blocks over the SMs, at least 8 rows (the 2r warm-up rows are amortised)
    const int rows = yb - ya;
    int rch = (2 * rows * nf + 295) / 296;
    rch = std::max(8, std::min(rch, 48));
    P.rch = rch;
    const int WC = 3 * r, LW = (WC + 3) & ~3, RW = (WC + cfg.nd + 3) & ~3;
    const size_t smem = (((size_t)(rch + 2 * r) * (LW + RW) + 15) & ~(size_t)15) + ((size_t)LW + (size_t)2 * r * cfg.nd) * sizeof(int);
    const int nt = ((cfg.nd + 31) / 32) * 32 + 32 * r;      // C warps + one winner warp per strip column
    dim3 grid((rows + rch - 1) / rch, 2, nf);
    cudaError_t e;
    switch (r) {
    case 2: e = launch_strip<2>(P, grid, nt, smem, st); break;
    case 3: e = launch_strip<3>(P, grid, nt, smem, st); break;
    case 4: e = launch_strip<4>(P, grid, nt, smem, st); break;
    case 5: e = launch_strip<5>(P, grid, nt, smem, st); break;
    case 6: e = launch_strip<6>(P, grid, nt, smem, st); break;
    case 7: e = launch_strip<7>(P, grid, nt, smem, st); break;
    case 8: e = launch_strip<8>(P, grid, nt, smem, st); break;
    case 9: e = launch_strip<9>(P, grid, nt, smem, st); break;
    case 10: e = launch_strip<10>(P, grid, nt, smem, st); break;
    default: return 0;
    }
    return e == cudaSuccess ? 1 : -1;
}

}  // namespace b200s
