// Border strips of cv::StereoBM for sm_100a.
//
// With disp12MaxDiff >= 0 (the state the reference's matcher is in out of the box, src/GPUStereoProcessor.cpp:23) OpenCV
// also needs the disparities of the r columns left and right of the valid ROI, where its window sums run against the
// clamped image border: they feed the left-right check (cv::validateDisparity) before the ROI mask
// removes them.  The warp-specialised matcher (bm_vh.cu) only covers the clamp-free rectangle; the generic kernels
// (bm_sad.cu) compute the two strips exactly but through a global cost volume with two latency-bound launches (56 us at
// 1080p).  This kernel does both strips of a frame (of a whole batch) in one launch and keeps everything on chip:
//
//   block = one strip x a chunk of RCH output rows; the chunk's rows of the left strip window and of the matching right
//           segment are staged in shared memory once (the border clamps are applied while staging, so the inner loops are
//           clamp-free);
//   thread = one disparity: vertical column sums of its 3r window columns in registers, sliding down the rows; per row
//           the r window sums go to shared memory;
//   warp   = one strip pixel of the row: argmin (lowest index on ties), uniqueness, texture, sub-pixel -- the same
//           selection as gen_winner_warp_body, from shared memory.
// Integer arithmetic in int32 throughout: bit-exact for any cap / window the 16-bit matcher accepts.
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

constexpr int STRIP_NT = 256;        // one thread per disparity: nd <= 256

template <int R>
__global__ void __launch_bounds__(STRIP_NT) bm_strip_fused_kernel(const StripParams P)
{
    constexpr int B = 2 * R + 1;
    constexpr int WC = 3 * R;                    // window columns of a strip: R output columns + 2R halo
    constexpr int LW = (WC + 3) & ~3;            // staged left bytes per row
    extern __shared__ __align__(16) uint8_t smem[];
    const int nd = P.nd;
    const int RW = (WC + nd + 3) & ~3;           // staged right bytes per row
    const int y0 = P.ya + blockIdx.x * P.rch, y1 = min(y0 + P.rch, P.yb);
    const int nrows = y1 - y0, NR = nrows + 2 * R;
    const int side = blockIdx.y;
    const int xa = P.xa[side];
    const uint8_t* __restrict__ Lp = P.Lp + blockIdx.z * P.pre_stride;
    const uint8_t* __restrict__ Rp = P.Rp + blockIdx.z * P.pre_stride;
    int16_t* __restrict__ disp = (int16_t*)((uint8_t*)P.disp + blockIdx.z * P.disp_stride);
    int16_t* __restrict__ cost = P.cost ? (int16_t*)((uint8_t*)P.cost + blockIdx.z * P.disp_stride) : nullptr;
    uint8_t* Ls = smem;                                           // [NRmax][LW]
    uint8_t* Rs = smem + (size_t)(P.rch + 2 * R) * LW;            // [NRmax][RW]
    int* Ssm = (int*)(smem + (((size_t)(P.rch + 2 * R) * (LW + RW) + 15) & ~(size_t)15));   // [R][nd]
    const int tid = threadIdx.x;
    // right-image column of window column `col` at disparity index 0 (rofs == 0 here)
    const int bmin = min(max(xa - R, 0), P.W - nd);
    // ---- stage: clamps applied here ----
    for (int i = tid; i < NR * LW; i += STRIP_NT) {
        const int row = i / LW, col = i - row * LW;
        const int lc = min(max(xa + col - R, -P.lofs), P.W - 1 - P.lofs) + P.lofs;
        Ls[i] = col < WC ? __ldg(Lp + (size_t)(y0 - R + row) * P.pitch + lc) : 0;
    }
    const int rbytes = min(max(xa + WC - 1 - R, 0), P.W - nd) - bmin + nd;      // distinct right bytes a row needs
    for (int i = tid; i < NR * RW; i += STRIP_NT) {
        const int row = i / RW, j = i - row * RW;
        Rs[i] = j < rbytes ? __ldg(Rp + (size_t)(y0 - R + row) * P.pitch + bmin + j) : 0;
    }
    __syncthreads();
    const int k = tid;
    const bool kact = k < nd;
    int boff[WC];
#pragma unroll
    for (int col = 0; col < WC; ++col) boff[col] = min(max(xa + col - R, 0), P.W - nd) - bmin + k;
    int C[WC];
#pragma unroll
    for (int col = 0; col < WC; ++col) C[col] = 0;
    if (kact) {
        for (int i = 0; i < B; ++i) {
            const uint8_t* lr = Ls + i * LW;
            const uint8_t* rr = Rs + i * RW;
#pragma unroll
            for (int col = 0; col < WC; ++col) C[col] += abs((int)lr[col] - (int)rr[boff[col]]);
        }
    }
    const int warp = tid >> 5, lane = tid & 31;
    const int16_t FILTERED = (int16_t)((P.minD - 1) * 16);
    for (int o = 0; o < nrows; ++o) {
        if (kact) {
            int s = 0;
#pragma unroll
            for (int col = 0; col < B; ++col) s += C[col];
#pragma unroll
            for (int c = 0; c < R; ++c) {
                Ssm[c * nd + k] = s;
                if (c + 1 < R) s += C[c + B] - C[c];
            }
        }
        __syncthreads();
        // ---- winner selection: one warp per strip pixel of this row ----
        for (int c = warp; c < R; c += STRIP_NT / 32) {
            const int* S = Ssm + c * nd;
            int best = INT_MAX, bk = nd;
            for (int kk = lane; kk < nd; kk += 32) {
                const int v = S[kk];
                if (v < best) { best = v; bk = kk; }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                const int ov = __shfl_xor_sync(0xffffffffu, best, d), ok2 = __shfl_xor_sync(0xffffffffu, bk, d);
                if (ov < best || (ov == best && ok2 < bk)) { best = ov; bk = ok2; }
            }
            const int minsad = best, mind = bk;
            const int X = xa + c + P.lofs, y = y0 + o;
            if (X < 0 || X >= P.W) continue;
            int tsum = 0;
            for (int t = lane; t < B * B; t += 32) {
                const int dy = t / B, dx = t - dy * B;
                tsum += abs((int)Ls[(o + dy) * LW + c + dx] - P.cap);
            }
            tsum = __reduce_add_sync(0xffffffffu, tsum);
            bool ok = tsum >= P.texThr;
            if (ok && P.uniq > 0) {
                const int thr = minsad + (minsad * P.uniq / 100);
                bool hit = false;
                for (int kk = lane; kk < nd; kk += 32)
                    if ((kk < mind - 1 || kk > mind + 1) && S[kk] <= thr) hit = true;
                if (__any_sync(0xffffffffu, hit)) ok = false;
            }
            if (lane == 0) {
                int16_t out = FILTERED;
                if (ok) {
                    const int p = S[mind + 1 < nd ? mind + 1 : nd - 2], n = S[mind > 0 ? mind - 1 : 1];
                    out = subpixel_disp(minsad, mind, p, n, nd, P.minD);
                    if (cost) cost[(size_t)y * P.W + X] = (int16_t)minsad;
                }
                disp[(size_t)y * P.W + X] = out;
            }
        }
        __syncthreads();
        if (kact && o + 1 < nrows) {
            const uint8_t* ln = Ls + (o + B) * LW;
            const uint8_t* rn = Rs + (o + B) * RW;
            const uint8_t* lo = Ls + o * LW;
            const uint8_t* ro = Rs + o * RW;
#pragma unroll
            for (int col = 0; col < WC; ++col)
                C[col] += abs((int)ln[col] - (int)rn[boff[col]]) - abs((int)lo[col] - (int)ro[boff[col]]);
        }
    }
}

template <int R>
static cudaError_t launch_strip(const StripParams& P, dim3 grid, size_t smem, cudaStream_t st)
{
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(bm_strip_fused_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    bm_strip_fused_kernel<R><<<grid, STRIP_NT, smem, st>>>(P);
    return cudaGetLastError();
}

// Both r-wide strips [xa0, xa0 + r) and [xa1, xa1 + r) (x = X - lofs) of rows [ya, yb) for nf frames.
// returns 1 when launched, 0 when the configuration is not handled (nd > 256, r outside 2..10), < 0 on CUDA errors
int launch_bm_strips(const uint8_t* Lp, const uint8_t* Rp, size_t pitch, int W, int H, const BMConfig& cfg, int r, int lofs,
                     int xa0, int xa1, int ya, int yb, int16_t* disp, int16_t* cost, cudaStream_t st, int nf, size_t pre_stride,
                     size_t disp_stride)
{
    if (cfg.nd > STRIP_NT || r < 2 || r > 10 || yb <= ya || W - cfg.nd < 0) return 0;
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
    const size_t smem = (((size_t)(rch + 2 * r) * (LW + RW) + 15) & ~(size_t)15) + (size_t)r * cfg.nd * sizeof(int);
    dim3 grid((rows + rch - 1) / rch, 2, nf);
    cudaError_t e;
    switch (r) {
    case 2: e = launch_strip<2>(P, grid, smem, st); break;
    case 3: e = launch_strip<3>(P, grid, smem, st); break;
    case 4: e = launch_strip<4>(P, grid, smem, st); break;
    case 5: e = launch_strip<5>(P, grid, smem, st); break;
    case 6: e = launch_strip<6>(P, grid, smem, st); break;
    case 7: e = launch_strip<7>(P, grid, smem, st); break;
    case 8: e = launch_strip<8>(P, grid, smem, st); break;
    case 9: e = launch_strip<9>(P, grid, smem, st); break;
    case 10: e = launch_strip<10>(P, grid, smem, st); break;
    default: return 0;
    }
    return e == cudaSuccess ? 1 : -1;
}

}  // namespace b200s
