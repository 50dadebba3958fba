// SAD block matcher with cv::StereoBM semantics for sm_100a (reference call site
// GpuStereoProcessor::computeDisparity, src/GPUStereoProcessor.cpp:264-321).
//
//   bm_fast_kernel     interior of the valid ROI (no border clamp active): packed-byte abs-diff
//                      (VABSDIFF4.U8), vertical sliding column sums in registers as u16x2 lanes, horizontal
//                      sliding row sums staged through shared memory, winner selection per pixel
//                      (argmin with largest-disparity tie rule, texture, uniqueness, sub-pixel fit).
//   bm_generic_*       exact clamped semantics for the border bands (needed when disp12MaxDiff >= 0 or
//                      minDisparity < 0) and for parameter sets outside the fast kernel's 16-bit range.
//
// Internal disparity index k in [0, nd): d = nd-1-k+minD; right column of (X, k) is X - lofs + k.
#include "kernels.h"
#include "bm_common.cuh"

#include <algorithm>
#include <climits>
#include <cstdlib>

namespace b200s {

int launch_bm_ws(const uint8_t* Lp, const uint8_t* Rp, size_t pitch, int W, int H, const BMConfig& cfg, int r, int lofs,
                 int XA, int XB, int YA, int YB, int16_t* disp, int16_t* cost, cudaStream_t st);
int launch_bm_vh(const uint8_t* Lp, const uint8_t* Rp, size_t pitch, int W, int H, const BMConfig& cfg, int r, int lofs,
                 int XA, int XB, int YA, int YB, int16_t* disp, int16_t* cost, cudaStream_t st, int nf, size_t pre_stride,
                 size_t disp_stride);
int launch_bm_strips(const uint8_t* Lp, const uint8_t* Rp, size_t pitch, int W, int H, const BMConfig& cfg, int r, int lofs,
                     int xa0, int xa1, int ya, int yb, int16_t* disp, int16_t* cost, cudaStream_t st, int nf, size_t pre_stride,
                     size_t disp_stride);

__global__ void fill_s16_kernel(int16_t* p, size_t n, int16_t v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// FILTERED everywhere outside [x0, x1) x [y0, y1) (that rectangle is written completely by the matcher kernels).
// The thread index space is compact: first the nb = x0 + (W - x1) border columns of every row, then the interior
// columns of the rows above y0 and below y1.
__global__ void __launch_bounds__(256) fill_border_kernel(int16_t* __restrict__ p, int W, int H, int x0, int x1, int y0, int y1, int16_t v,
                                                          size_t frame_stride)
{
    p = (int16_t*)((uint8_t*)p + blockIdx.y * frame_stride);      // blockIdx.y = frame of the batch
    const int nb = x0 + (W - x1), ni = x1 - x0, nrows_tb = y0 + (H - y1);
    const long long n_side = (long long)nb * H, total = n_side + (long long)ni * nrows_tb;
    long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    int x, y;
    if (i < n_side) {
        y = (int)(i / nb);
        const int c = (int)(i - (long long)y * nb);
        x = c < x0 ? c : x1 + (c - x0);
    } else {
        i -= n_side;
        const int rr = (int)(i / ni);
        x = x0 + (int)(i - (long long)rr * ni);
        y = rr < y0 ? rr : y1 + (rr - y0);
    }
    p[(size_t)y * W + x] = v;
}

// ------------------------------------------------------------------------------------------------------
// fast kernel (v3)
//
// One block = TW output columns x BH output rows, all nd disparities; it marches down the rows.  Per row:
//   V  thread = (4 adjacent window columns, two groups of 8 disparities): abs-diff of the entering and the leaving
//      row (VABSDIFF4.U8, 4 disparities per instruction; the right-image window of column i is the thread's
//      register window funnel-shifted by i bytes), biased byte delta, widened to u16x2 lanes and added to the 32
//      column-sum registers; the sums go to shared memory (Cbuf, one row per window column, linear).
//   H  thread = (8 disparities, strip of columns): horizontal sliding sum over the 2r+1 window columns -> Sbuf,
//      plus one 32-bit key per (pixel, 8 disparities): packed-u16x2 minimum << 16 | group -> Kbuf.
//   W  one thread per pixel: min over the keys -> winning group, exact index inside it (lowest index wins ties =
//      largest disparity), uniqueness from the keys with the three neighbouring groups re-scanned exactly, texture,
//      sub-pixel fit.  The threads that own no pixel stage the next two image rows meanwhile.
// All shared-memory layouts are linear with compile-time strides (ND > 0) so that unrolled loops use immediate
// offsets; row strides are odd multiples of 16 bytes where a warp walks over rows.
// ------------------------------------------------------------------------------------------------------
struct FastParams {
    const uint8_t* Lp;    // prefiltered planes, row pitch `pitch` (multiple of 16), with slack before/after
    const uint8_t* Rp;
    size_t pitch;
    int16_t* disp;        // tightly packed W
    int16_t* cost;        // may be null
    int W, H, nd, minD, r, cap, texThr, uniq, lofs;
    int X0base, XA, XB, YA, YB;   // first tile origin (<= XA, aligned), output columns [XA, XB), output rows
    int TW, BH, ncols;            // tile width, band height, window columns (= 4 * NCQ)
    int NCQ, NK;                  // V items: column quads x (nd / 16) disparity groups
    int NGH, NS, SWD;             // H items: (nd / 8) groups x strips of SWD columns
    int CWb, SWb, KWb;            // row strides of Cbuf / Sbuf / Kbuf in bytes
    int NK4;                      // uint4 loads per Kbuf row
    int CSB, RLW;                 // bytes per right-row copy, words staged per right row
    int stage_from;               // threads >= stage_from stage rows during phase W
    int oLb, oRc, oT, oK, oC, oS; // shared memory byte offsets
};

__device__ __forceinline__ void stage_rows(const FastParams& P, uint8_t* smem, int t, int nt, int yi, bool has_old,
                                           int Xl0, int Xr0)
{
    uint32_t* sLb = (uint32_t*)(smem + P.oLb);
    uint8_t* sRc = smem + P.oRc;
    const int b = 2 * P.r + 1;
    const uint8_t* ln = P.Lp + (size_t)yi * P.pitch + Xl0;
    const uint8_t* lo = P.Lp + (size_t)max(yi - b, 0) * P.pitch + Xl0;
    for (int c = t; c < P.ncols; c += nt) {
        sLb[c] = (uint32_t)__ldg(ln + c) * 0x01010101u;
        sLb[P.ncols + c] = has_old ? (uint32_t)__ldg(lo + c) * 0x01010101u : 0u;
    }
    const uint32_t* rn = (const uint32_t*)(P.Rp + (size_t)yi * P.pitch + Xr0);          // Xr0 % 4 == 0, pitch % 16 == 0
    const uint32_t* ro = (const uint32_t*)(P.Rp + (size_t)max(yi - b, 0) * P.pitch + Xr0);
    for (int i = t; i < 2 * P.RLW; i += nt) {
        const int s = i >= P.RLW;
        const int wi = i - s * P.RLW;
        uint32_t v = s ? (has_old ? __ldg(ro + wi) : 0u) : __ldg(rn + wi);
        uint8_t* cp = sRc + (size_t)(s * 4) * P.CSB + 4 * wi;
        // copy j holds row[a + 4j] at byte a
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (wi >= j) *(uint32_t*)(cp + (size_t)j * P.CSB - 4 * j) = v;
    }
}

template <int ND>
__global__ void __launch_bounds__(384, 2) bm_fast_kernel(const FastParams P)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t* sLb = (const uint32_t*)(smem + P.oLb);   // [2][ncols]  left bytes x 0x01010101
    const uint8_t* sRc = smem + P.oRc;                       // [2][4][CSB] right row, copy j shifted by 4j bytes
    uint32_t* sT = (uint32_t*)(smem + P.oT);                 // [ncols]     texture column sums
    uint8_t* sK = smem + P.oK;                               // [rowsS][KWb] keys
    uint8_t* sC = smem + P.oC;                               // [rowsC][CWb] column sums, u16x2
    uint8_t* sS = smem + P.oS;                               // [rowsS][SWb] window sums, u16x2

    const int nd = ND > 0 ? ND : P.nd;
    const int CWb = ND > 0 ? ND * 2 : P.CWb;
    const int SWb = ND > 0 ? ND * 2 + 16 : P.SWb;
    const int KWb = ND > 0 ? (((ND / 8 + 3) / 4) * 4 + 4) * 4 : P.KWb;
    const int NK = ND > 0 ? ND / 16 : P.NK;
    const int NGH = ND > 0 ? ND / 8 : P.NGH;
    const int NK4 = ND > 0 ? (ND / 8 + 3) / 4 : P.NK4;

    const int tid = threadIdx.x, NT = blockDim.x;
    const int X0 = P.X0base + blockIdx.x * P.TW;             // first output column of the tile
    const int yb0 = P.YA + blockIdx.y * P.BH;
    const int yb1 = min(yb0 + P.BH, P.YB);
    const int r = P.r, b = 2 * r + 1;
    const int Xl0 = X0 - r;               // left image column of window column c = 0
    const int Xr0 = X0 - r - P.lofs;      // right image column of (c = 0, k = 0); multiple of 4 by construction

    // V identity: lanes run over the disparity groups of one column quad (contiguous 16-byte units per store)
    const int cq = tid / NK, kg = tid - cq * NK;
    const bool vact = cq < P.NCQ;
    uint32_t C[4][2][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int w = 0; w < 4; ++w) C[i][h][w] = 0;
    const uint8_t* vpn[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int u = kg + h * NK;
        const int j = (cq + 2 * u) & 3;
        vpn[h] = sRc + (size_t)j * P.CSB + (4 * cq + 8 * u - 4 * j);
    }
    uint8_t* vst = sC + (size_t)(4 * cq) * CWb + 16 * kg;
    // H identity
    const int hs = tid / NGH, gh = tid - hs * NGH;
    const bool hact = hs < P.NS;

    for (int c = tid; c < P.ncols; c += NT) sT[c] = 0;
    for (int i = tid; i < (P.NS * P.SWD * KWb) / 4; i += NT) ((uint32_t*)sK)[i] = 0xFFFFFFFFu;   // pad keys stay "infinite"
    stage_rows(P, smem, tid, NT, yb0 - r, false, Xl0, Xr0);
    __syncthreads();

    int nrow = 0;   // rows accumulated so far (bias bookkeeping: every row adds 128 per u16 lane)
    for (int yi = yb0 - r; yi < yb1 + r; ++yi) {
        const bool has_old = (yi - b) >= yb0 - r;
        const bool do_out = yi >= yb0 + r;
        // ---- phase V ------------------------------------------------------------------------------------
        if (vact) {
            const uint4 ln4 = *(const uint4*)(sLb + 4 * cq);
            const uint4 lo4 = *(const uint4*)(sLb + P.ncols + 4 * cq);
            const uint32_t ln[4] = {ln4.x, ln4.y, ln4.z, ln4.w};
            const uint32_t lo[4] = {lo4.x, lo4.y, lo4.z, lo4.w};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint4 rn4 = *(const uint4*)vpn[h];
                const uint4 ro4 = *(const uint4*)(vpn[h] + (size_t)4 * P.CSB);
                const uint32_t rn[3] = {rn4.x, rn4.y, rn4.z};
                const uint32_t ro[3] = {ro4.x, ro4.y, ro4.z};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
#pragma unroll
                    for (int w = 0; w < 2; ++w) {
                        const uint32_t wn = i ? __funnelshift_r(rn[w], rn[w + 1], 8 * i) : rn[w];
                        const uint32_t wo = i ? __funnelshift_r(ro[w], ro[w + 1], 8 * i) : ro[w];
                        const uint32_t an = __vabsdiffu4(ln[i], wn);
                        const uint32_t ao = __vabsdiffu4(lo[i], wo);
                        const uint32_t t = an + 0x80808080u - ao;          // per byte: 128 + new - old, no borrow
                        C[i][h][2 * w] += t & 0x00ff00ffu;                 // lanes k+0, k+2   (bias 128 per lane kept)
                        C[i][h][2 * w + 1] += __byte_perm(t, 0, 0x4341);   // lanes k+1, k+3
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    *(uint4*)(vst + (size_t)i * CWb + (size_t)h * (16 * NK)) = make_uint4(C[i][h][0], C[i][h][1], C[i][h][2], C[i][h][3]);
        }
        for (int c = tid; c < P.ncols; c += NT) {   // texture column sums live in shared memory (same owner every row)
            const int cap = P.cap;
            const int lnv = (int)(sLb[c] & 0xffu), lov = (int)(sLb[P.ncols + c] & 0xffu);
            sT[c] += (uint32_t)(abs(lnv - cap) - (has_old ? abs(lov - cap) : 0));
        }
        ++nrow;
        __syncthreads();
        if (!do_out) {                    // warm-up rows: only stage the next pair of rows
            stage_rows(P, smem, tid, NT, yi + 1, (yi + 1 - b) >= yb0 - r, Xl0, Xr0);
            __syncthreads();
            continue;
        }
        // ---- phase H ------------------------------------------------------------------------------------
        if (hact) {
            const int xs = hs * P.SWD;
            const uint8_t* pc = sC + (size_t)xs * CWb + 16 * gh;
            // every column sum carries a bias of 128 * nrow per lane; remove b of them from the window sum
            const uint32_t bias = (uint32_t)(128 * nrow * b) * 0x00010001u;
            uint4 S = make_uint4(0u - bias, 0u - bias, 0u - bias, 0u - bias);
            for (int c = 0; c < b; ++c) {
                const uint4 v = *(const uint4*)(pc + (size_t)c * CWb);
                S.x += v.x; S.y += v.y; S.z += v.z; S.w += v.w;
            }
            const uint8_t* pa = pc + (size_t)b * CWb;
            uint8_t* ps = sS + (size_t)xs * SWb + 16 * gh;
            uint8_t* pk = sK + (size_t)xs * KWb + 4 * gh;
#pragma unroll 2
            for (int x = 0; x < P.SWD; ++x) {
                *(uint4*)ps = S;
                uint32_t m = __vimin3_u16x2(S.x, S.y, S.z);
                m = __vminu2(m, S.w);
                m = __vminu2(m, m >> 16);
                *(uint32_t*)pk = (m << 16) | (uint32_t)gh;
                const uint4 a = *(const uint4*)pa;
                const uint4 o = *(const uint4*)pc;
                S.x += a.x - o.x; S.y += a.y - o.y; S.z += a.z - o.z; S.w += a.w - o.w;
                pa += CWb; pc += CWb; ps += SWb; pk += KWb;
            }
        }
        __syncthreads();
        // ---- phase W (+ staging of the next rows by the threads that own no pixel) ----------------------------
        if (tid >= P.stage_from && yi + 1 < yb1 + r)
            stage_rows(P, smem, tid - P.stage_from, NT - P.stage_from, yi + 1, (yi + 1 - b) >= yb0 - r, Xl0, Xr0);
        if (tid < P.TW) {
            const int px = tid;
            const int y = yi - r;
            const int16_t FILTERED = (int16_t)((P.minD - 1) * 16);
            uint8_t* krow = sK + (size_t)px * KWb;
            uint8_t* srow = sS + (size_t)px * SWb;
            uint32_t best = 0xFFFFFFFFu;
#pragma unroll
            for (int i = 0; i < (ND > 0 ? NK4 : 1); ++i) {
                if (ND > 0) {
                    const uint4 k4 = *(const uint4*)(krow + 16 * i);
                    best = min(min(best, k4.x), min(k4.y, min(k4.z, k4.w)));
                }
            }
            if (ND == 0)
                for (int i = 0; i < NK4; ++i) {
                    const uint4 k4 = *(const uint4*)(krow + 16 * i);
                    best = min(min(best, k4.x), min(k4.y, min(k4.z, k4.w)));
                }
            const int minsad = (int)(best >> 16), gs = (int)(best & 0xffffu);
            // exact index inside the winning group, in disparity-index order (lowest k wins ties)
            int mind;
            {
                const uint4 u = *(const uint4*)(srow + 16 * gs);
                const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
                int loc = 7;
#pragma unroll
                for (int kk = 7; kk >= 0; --kk) {
                    const uint32_t w = wv[2 * (kk >> 2) + (kk & 1)];
                    const uint32_t v = (kk & 2) ? (w >> 16) : (w & 0xffffu);
                    if ((int)v == minsad) loc = kk;
                }
                mind = 8 * gs + loc;
            }
            uint16_t* s16 = (uint16_t*)srow;
            const int pv = s16[kpos(mind + 1 < nd ? mind + 1 : nd - 2)];
            const int nv = s16[kpos(mind > 0 ? mind - 1 : 1)];
            bool filtered = false;
            if (P.uniq > 0) {
                // group level: the keys of the winner's group and of both neighbours are taken out ...
                const int g0 = max(gs - 1, 0), g2 = min(gs + 1, NGH - 1);
                ((uint32_t*)krow)[g0] = 0xFFFFFFFFu;
                ((uint32_t*)krow)[gs] = 0xFFFFFFFFu;
                ((uint32_t*)krow)[g2] = 0xFFFFFFFFu;
                // ... and those three groups are re-scanned exactly with mind-1, mind, mind+1 masked
                s16[kpos(mind)] = 0xFFFFu;
                if (mind > 0) s16[kpos(mind - 1)] = 0xFFFFu;
                if (mind + 1 < nd) s16[kpos(mind + 1)] = 0xFFFFu;
                uint32_t m2k = 0xFFFFFFFFu;
                if (ND > 0) {
#pragma unroll
                    for (int i = 0; i < NK4; ++i) {
                        const uint4 k4 = *(const uint4*)(krow + 16 * i);
                        m2k = min(min(m2k, k4.x), min(k4.y, min(k4.z, k4.w)));
                    }
                } else {
                    for (int i = 0; i < NK4; ++i) {
                        const uint4 k4 = *(const uint4*)(krow + 16 * i);
                        m2k = min(min(m2k, k4.x), min(k4.y, min(k4.z, k4.w)));
                    }
                }
                const uint4 e0 = *(const uint4*)(srow + 16 * g0);
                const uint4 e1 = *(const uint4*)(srow + 16 * gs);
                const uint4 e2 = *(const uint4*)(srow + 16 * g2);
                uint32_t acc = __vimin3_u16x2(e0.x, e0.y, e0.z);
                acc = __vimin3_u16x2(acc, e0.w, e1.x);
                acc = __vimin3_u16x2(acc, e1.y, e1.z);
                acc = __vimin3_u16x2(acc, e1.w, e2.x);
                acc = __vimin3_u16x2(acc, e2.y, e2.z);
                acc = __vminu2(acc, e2.w);
                const uint32_t m2 = min(min(acc & 0xffffu, acc >> 16), m2k >> 16);
                const int thr = minsad + (minsad * P.uniq / 100);
                filtered = (int)m2 <= thr;
            }
            const int X = X0 + px;
            if (X >= P.XA && X < P.XB) {
                int tsum = 0;
                for (int c = px; c < px + b; ++c) tsum += (int)sT[c];
                int16_t out = FILTERED;
                if (tsum >= P.texThr && !filtered) out = subpixel_disp(minsad, mind, pv, nv, nd, P.minD);
                P.disp[(size_t)y * P.W + X] = out;
                if (P.cost) P.cost[(size_t)y * P.W + X] = (int16_t)minsad;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------
// generic path: exact clamped semantics, int32 sums, cost volume in global scratch
// ------------------------------------------------------------------------------------------------------
struct GenPlanes { const uint8_t* Lp; const uint8_t* Rp; size_t pitch; };

struct GenParams {
    const uint8_t* Lp;
    const uint8_t* Rp;
    size_t pitch;
    int W, H, nd, minD, r, cap, texThr, uniq, lofs, rofs;
    int xa, xb;   // x = X - lofs range handled
    int ya, yb;   // rows handled by this launch (chunk)
    int RCH;      // rows per thread-march
    int* vol;     // [(y-ya)][(x-xa)][k]
    int16_t* disp;
    int16_t* cost;
};

__device__ __forceinline__ int gen_row_sum(const GenParams& P, int y, int x, int k)
{
    const uint8_t* lr = P.Lp + (size_t)y * P.pitch;
    const uint8_t* rr = P.Rp + (size_t)y * P.pitch;
    int s = 0;
    for (int dx = -P.r; dx <= P.r; ++dx) {
        int xp = x + dx;
        int lc = min(max(xp, -P.lofs), P.W - 1 - P.lofs) + P.lofs;
        int rc = min(max(xp, -P.rofs), P.W - P.nd - P.rofs) + P.rofs + k;
        s += abs((int)__ldg(lr + lc) - (int)__ldg(rr + rc));
    }
    return s;
}

__global__ void __launch_bounds__(128) bm_generic_cost_kernel(const GenParams P)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int x = P.xa + blockIdx.y;
    const int y0 = P.ya + blockIdx.z * P.RCH, y1 = min(y0 + P.RCH, P.yb);
    if (k >= P.nd || y0 >= y1) return;
    const int ncx = P.xb - P.xa;
    int S = 0;
    for (int yy = y0 - P.r; yy <= y0 + P.r; ++yy) S += gen_row_sum(P, yy, x, k);
    for (int y = y0; y < y1; ++y) {
        P.vol[((size_t)(y - P.ya) * ncx + (x - P.xa)) * P.nd + k] = S;
        if (y + 1 < y1) S += gen_row_sum(P, y + 1 + P.r, x, k) - gen_row_sum(P, y - P.r, x, k);
    }
}

// Narrow bands (at most GEN_STRIP_MAXC columns, e.g. the r-wide strips next to the fast rectangle): one thread per
// (disparity, row march) walks the columns itself, so the horizontal window sum slides (2 terms per further column
// instead of 2r+1) and the vertical sums of all strip columns stay in registers.
constexpr int GEN_STRIP_MAXC = 16;

__device__ __forceinline__ int gen_term(const GenParams& P, const uint8_t* __restrict__ lr, const uint8_t* __restrict__ rr, int xp, int k)
{
    const int lc = min(max(xp, -P.lofs), P.W - 1 - P.lofs) + P.lofs;
    const int rc = min(max(xp, -P.rofs), P.W - P.nd - P.rofs) + P.rofs + k;
    return abs((int)__ldg(lr + lc) - (int)__ldg(rr + rc));
}

// adds sign * (window sums of row y for the strip columns) to V.  RT > 0: compile-time radius, so that the 2r+1 terms of
// the first window are unrolled and their loads overlap (the kernel is latency-bound otherwise).
template <int RT>
__device__ __forceinline__ void gen_strip_row(const GenParams& P, int y, int k, int ncx, int sign, int (&V)[GEN_STRIP_MAXC])
{
    const uint8_t* lr = P.Lp + (size_t)y * P.pitch;
    const uint8_t* rr = P.Rp + (size_t)y * P.pitch;
    const int r = RT > 0 ? RT : P.r;
    int s = 0;
    if (RT > 0) {
#pragma unroll
        for (int dx = -RT; dx <= RT; ++dx) s += gen_term(P, lr, rr, P.xa + dx, k);
    } else {
        for (int dx = -r; dx <= r; ++dx) s += gen_term(P, lr, rr, P.xa + dx, k);
    }
#pragma unroll
    for (int c = 0; c < GEN_STRIP_MAXC; ++c) {
        if (c < ncx) {
            V[c] += sign * s;
            if (c + 1 < ncx) s += gen_term(P, lr, rr, P.xa + c + 1 + r, k) - gen_term(P, lr, rr, P.xa + c - r, k);
        }
    }
}

template <int RT> __device__ __forceinline__ void gen_strip_cost_body(const GenParams& P);
__device__ __forceinline__ void gen_winner_warp_body(const GenParams& P);

// both border strips of a frame in one launch each (blockIdx.z selects the strip): one strip alone leaves most SMs idle
template <int RT>
__global__ void __launch_bounds__(128) bm_generic_strip_cost_pair_kernel(const GenParams P0, const GenParams P1)
{
    gen_strip_cost_body<RT>(blockIdx.z ? P1 : P0);
}
__global__ void __launch_bounds__(256) bm_generic_winner_warp_pair_kernel(const GenParams P0, const GenParams P1)
{
    gen_winner_warp_body(blockIdx.y ? P1 : P0);
}

__global__ void __launch_bounds__(128) bm_generic_strip_cost_kernel(const GenParams P) { gen_strip_cost_body<0>(P); }

template <int RT>
__device__ __forceinline__ void gen_strip_cost_body(const GenParams& P)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int y0 = P.ya + blockIdx.y * P.RCH, y1 = min(y0 + P.RCH, P.yb);
    if (k >= P.nd || y0 >= y1) return;
    const int ncx = P.xb - P.xa;
    int V[GEN_STRIP_MAXC];
#pragma unroll
    for (int c = 0; c < GEN_STRIP_MAXC; ++c) V[c] = 0;
    for (int yy = y0 - P.r; yy <= y0 + P.r; ++yy) gen_strip_row<RT>(P, yy, k, ncx, 1, V);
    for (int y = y0; y < y1; ++y) {
#pragma unroll
        for (int c = 0; c < GEN_STRIP_MAXC; ++c)
            if (c < ncx) P.vol[((size_t)(y - P.ya) * ncx + c) * P.nd + k] = V[c];
        if (y + 1 < y1) {
            gen_strip_row<RT>(P, y + 1 + P.r, k, ncx, 1, V);
            gen_strip_row<RT>(P, y - P.r, k, ncx, -1, V);
        }
    }
}

__global__ void __launch_bounds__(128) bm_generic_winner_kernel(const GenParams P)
{
    const int ncx = P.xb - P.xa;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int npx = ncx * (P.yb - P.ya);
    if (idx >= npx) return;
    const int y = P.ya + idx / ncx, x = P.xa + idx % ncx;
    const int* S = P.vol + (size_t)idx * P.nd;
    const int nd = P.nd;
    int minsad = INT_MAX, mind = -1;
    for (int k = 0; k < nd; ++k) {
        int v = S[k];
        if (v < minsad) { minsad = v; mind = k; }
    }
    const int X = x + P.lofs;
    const int16_t FILTERED = (int16_t)((P.minD - 1) * 16);
    if (X < 0 || X >= P.W) return;
    int tsum = 0;
    for (int dy = -P.r; dy <= P.r; ++dy) {
        const uint8_t* lr = P.Lp + (size_t)(y + dy) * P.pitch;
        for (int dx = -P.r; dx <= P.r; ++dx) {
            int lc = min(max(x + dx, -P.lofs), P.W - 1 - P.lofs) + P.lofs;
            tsum += abs((int)__ldg(lr + lc) - P.cap);
        }
    }
    int16_t out = FILTERED;
    bool ok = tsum >= P.texThr;
    if (ok && P.uniq > 0) {
        int thr = minsad + (minsad * P.uniq / 100);
        for (int k = 0; k < nd; ++k)
            if ((k < mind - 1 || k > mind + 1) && S[k] <= thr) { ok = false; break; }
    }
    if (ok) {
        int p = S[mind + 1 < nd ? mind + 1 : nd - 2], n = S[mind > 0 ? mind - 1 : 1];
        out = subpixel_disp(minsad, mind, p, n, nd, P.minD);
        if (P.cost) P.cost[(size_t)y * P.W + X] = (int16_t)minsad;
    }
    P.disp[(size_t)y * P.W + X] = out;
}

// Same selection with one warp per pixel (lanes stride over the disparities): the border bands that disp12MaxDiff >= 0
// needs are only r columns wide, far too few pixels to fill the GPU with one thread each.
__global__ void __launch_bounds__(256) bm_generic_winner_warp_kernel(const GenParams P) { gen_winner_warp_body(P); }

__device__ __forceinline__ void gen_winner_warp_body(const GenParams& P)
{
    const int ncx = P.xb - P.xa;
    const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int npx = ncx * (P.yb - P.ya);
    if (idx >= npx) return;
    const int y = P.ya + idx / ncx, x = P.xa + idx % ncx;
    const int* S = P.vol + (size_t)idx * P.nd;
    const int nd = P.nd;
    // argmin with the lowest index on ties: key = (sad << 9 | k) would overflow for big windows, so compare pairs
    int best = INT_MAX, bk = nd;
    for (int k = lane; k < nd; k += 32) {
        const int v = S[k];
        if (v < best) { best = v; bk = k; }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const int ov = __shfl_xor_sync(0xffffffffu, best, d), ok = __shfl_xor_sync(0xffffffffu, bk, d);
        if (ov < best || (ov == best && ok < bk)) { best = ov; bk = ok; }
    }
    const int minsad = best, mind = bk;
    const int X = x + P.lofs;
    const int16_t FILTERED = (int16_t)((P.minD - 1) * 16);
    if (X < 0 || X >= P.W) return;
    const int b = 2 * P.r + 1;
    int tsum = 0;
    for (int t = lane; t < b * b; t += 32) {
        const int dy = t / b - P.r, dx = t % b - P.r;
        const int lc = min(max(x + dx, -P.lofs), P.W - 1 - P.lofs) + P.lofs;
        tsum += abs((int)__ldg(P.Lp + (size_t)(y + dy) * P.pitch + lc) - P.cap);
    }
    tsum = __reduce_add_sync(0xffffffffu, tsum);
    bool ok = tsum >= P.texThr;
    if (ok && P.uniq > 0) {
        const int thr = minsad + (minsad * P.uniq / 100);
        bool hit = false;
        for (int k = lane; k < nd; k += 32)
            if ((k < mind - 1 || k > mind + 1) && S[k] <= thr) hit = true;
        if (__any_sync(0xffffffffu, hit)) ok = false;
    }
    if (lane == 0) {
        int16_t out = FILTERED;
        if (ok) {
            const int p = S[mind + 1 < nd ? mind + 1 : nd - 2], n = S[mind > 0 ? mind - 1 : 1];
            out = subpixel_disp(minsad, mind, p, n, nd, P.minD);
            if (P.cost) P.cost[(size_t)y * P.W + X] = (int16_t)minsad;
        }
        P.disp[(size_t)y * P.W + X] = out;
    }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
struct Geom {
    int lofs, rofs, width1, r;
    int roiX0, roiX1, roiY0, roiY1;   // valid-disparity ROI (getValidDisparityROI with full-image rois)
    bool degenerate;
};

static Geom geom(int W, int H, const BMConfig& c)
{
    Geom g;
    int t = c.nd - 1 + c.minD;
    g.lofs = t > 0 ? t : 0;
    g.rofs = t < 0 ? -t : 0;
    g.width1 = W - g.rofs - c.nd + 1;
    g.r = c.wsz / 2;
    g.roiX0 = std::max(c.minD + c.nd - 1, 0) + g.r;
    g.roiX1 = W - g.r;
    g.roiY0 = g.r;
    g.roiY1 = H - g.r;
    g.degenerate = g.lofs >= W || g.rofs >= W || g.width1 < 1 || g.roiY1 <= g.roiY0;
    return g;
}

static const size_t GEN_VOL_BUDGET = (size_t)192 << 20;

size_t bm_scratch_bytes(int W, int H, const BMConfig& cfg)
{
    (void)W; (void)H; (void)cfg;
    return GEN_VOL_BUDGET;
}

static int run_generic(const GenPlanes& pl, int W, int H, const BMConfig& cfg, const Geom& g,
                       int xa, int xb, int16_t* disp, int16_t* cost, BMScratch* sc, cudaStream_t st)
{
    if (xb <= xa) return 0;
    if (!sc || !sc->vol) return -2;
    int launches = 0;
    const int ncx = xb - xa;
    size_t per_row = (size_t)ncx * cfg.nd * sizeof(int);
    int rows_per_chunk = (int)std::max<size_t>(1, std::min<size_t>(sc->vol_bytes / per_row, (size_t)(g.roiY1 - g.roiY0)));
    if (per_row > sc->vol_bytes) return -3;
    GenParams P;
    P.Lp = pl.Lp; P.Rp = pl.Rp; P.pitch = pl.pitch; P.W = W; P.H = H; P.nd = cfg.nd; P.minD = cfg.minD; P.r = g.r; P.cap = cfg.cap;
    P.texThr = cfg.textureThreshold; P.uniq = cfg.uniquenessRatio; P.lofs = g.lofs; P.rofs = g.rofs;
    P.xa = xa; P.xb = xb; P.vol = sc->vol; P.disp = disp; P.cost = cost;
    // rows per thread-march: narrow bands (the r-wide strips next to the fast rectangle) would leave most SMs idle with
    // long marches, so they get short ones (more threads, a few more warm-up rows each)
    const long long warps_at_32 = (long long)((cfg.nd + 31) / 32) * ncx * ((g.roiY1 - g.roiY0 + 31) / 32);
    P.RCH = warps_at_32 >= 8192 ? 32 : (warps_at_32 >= 4096 ? 16 : 8);
    for (int ya = g.roiY0; ya < g.roiY1; ya += rows_per_chunk) {
        P.ya = ya;
        P.yb = std::min(g.roiY1, ya + rows_per_chunk);
        int kt = std::min(128, cfg.nd);
        dim3 grid((cfg.nd + kt - 1) / kt, ncx, (P.yb - P.ya + P.RCH - 1) / P.RCH);
        // gridDim.y/z limits (65535) are far above any supported image size
        if (ncx <= GEN_STRIP_MAXC) bm_generic_strip_cost_kernel<<<dim3(grid.x, grid.z), kt, 0, st>>>(P);
        else bm_generic_cost_kernel<<<grid, kt, 0, st>>>(P);
        int npx = ncx * (P.yb - P.ya);
        if (npx < 262144) bm_generic_winner_warp_kernel<<<(npx + 7) / 8, 256, 0, st>>>(P);
        else bm_generic_winner_kernel<<<(npx + 127) / 128, 128, 0, st>>>(P);
        launches += 2;
    }
    return launches;
}

// the two r-wide strips left and right of the fast rectangle, one cost launch and one winner launch for both
static int run_generic_pair(const GenPlanes& pl, int W, int H, const BMConfig& cfg, const Geom& g, int xa0, int xb0, int xa1,
                            int xb1, int16_t* disp, int16_t* cost, BMScratch* sc, cudaStream_t st)
{
    const int n0 = xb0 - xa0, n1 = xb1 - xa1, rows = g.roiY1 - g.roiY0;
    const size_t need = (size_t)(n0 + n1) * cfg.nd * sizeof(int) * (size_t)rows;
    if (n0 <= 0 || n1 <= 0 || n0 > GEN_STRIP_MAXC || n1 > GEN_STRIP_MAXC || !sc || !sc->vol || need > sc->vol_bytes || rows <= 0) {
        int l0 = run_generic(pl, W, H, cfg, g, xa0, xb0, disp, cost, sc, st);
        if (l0 < 0) return l0;
        int l1 = run_generic(pl, W, H, cfg, g, xa1, xb1, disp, cost, sc, st);
        return l1 < 0 ? l1 : l0 + l1;
    }
    GenParams P[2];
    for (int i = 0; i < 2; ++i) {
        GenParams& Q = P[i];
        Q.Lp = pl.Lp; Q.Rp = pl.Rp; Q.pitch = pl.pitch; Q.W = W; Q.H = H; Q.nd = cfg.nd; Q.minD = cfg.minD; Q.r = g.r; Q.cap = cfg.cap;
        Q.texThr = cfg.textureThreshold; Q.uniq = cfg.uniquenessRatio; Q.lofs = g.lofs; Q.rofs = g.rofs;
        Q.xa = i ? xa1 : xa0; Q.xb = i ? xb1 : xb0; Q.ya = g.roiY0; Q.yb = g.roiY1; Q.RCH = 8;
        Q.vol = sc->vol + (i ? (size_t)n0 * cfg.nd * rows : 0); Q.disp = disp; Q.cost = cost;
    }
    const int kt = std::min(128, cfg.nd);
    const dim3 gc((cfg.nd + kt - 1) / kt, (rows + 7) / 8, 2);
    switch (g.r) {
    case 2: bm_generic_strip_cost_pair_kernel<2><<<gc, kt, 0, st>>>(P[0], P[1]); break;
    case 3: bm_generic_strip_cost_pair_kernel<3><<<gc, kt, 0, st>>>(P[0], P[1]); break;
    case 4: bm_generic_strip_cost_pair_kernel<4><<<gc, kt, 0, st>>>(P[0], P[1]); break;
    case 5: bm_generic_strip_cost_pair_kernel<5><<<gc, kt, 0, st>>>(P[0], P[1]); break;
    case 6: bm_generic_strip_cost_pair_kernel<6><<<gc, kt, 0, st>>>(P[0], P[1]); break;
    case 7: bm_generic_strip_cost_pair_kernel<7><<<gc, kt, 0, st>>>(P[0], P[1]); break;
    case 8: bm_generic_strip_cost_pair_kernel<8><<<gc, kt, 0, st>>>(P[0], P[1]); break;
    case 9: bm_generic_strip_cost_pair_kernel<9><<<gc, kt, 0, st>>>(P[0], P[1]); break;
    case 10: bm_generic_strip_cost_pair_kernel<10><<<gc, kt, 0, st>>>(P[0], P[1]); break;
    default: bm_generic_strip_cost_pair_kernel<0><<<gc, kt, 0, st>>>(P[0], P[1]); break;
    }
    const int npx = std::max(n0, n1) * rows;
    bm_generic_winner_warp_pair_kernel<<<dim3((npx + 7) / 8, 2), 256, 0, st>>>(P[0], P[1]);
    return 2;
}

template <int ND>
static cudaError_t launch_fast(const FastParams& P, dim3 grid, int nt, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(bm_fast_kernel<ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    bm_fast_kernel<ND><<<grid, nt, smem, st>>>(P);
    return cudaGetLastError();
}

// one frame (nf == 1), or a whole batch when the warp-specialised matcher covers everything the configuration needs;
// NOT_BATCHABLE tells the caller to go frame by frame
static const int NOT_BATCHABLE = -1000;

static int block_match_impl(const uint8_t* Lp, const uint8_t* Rp, size_t pitch, int W, int H, const BMConfig& cfg,
                            int16_t* disp, int16_t* cost, BMScratch* sc, cudaStream_t st, double* evals, int nf,
                            size_t pre_stride, size_t disp_stride, bool border_is_filled)
{
    const Geom g = geom(W, H, cfg);
    const int16_t FILTERED = (int16_t)((cfg.minD - 1) * 16);
    int launches = 0;
    size_t n = (size_t)W * H;
    if (evals) *evals = 0;
    if (nf > 1 && g.degenerate) return NOT_BATCHABLE;
    if (cost)
        for (int f = 0; f < nf; ++f) cudaMemsetAsync((uint8_t*)cost + f * disp_stride, 0, n * sizeof(int16_t), st);
    if (g.degenerate) {
        fill_s16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(disp, n, FILTERED);
        return launches + 1;
    }
    const bool need_bands = cfg.disp12MaxDiff >= 0;
    // columns (in X) the matcher produces at all, and the ones that can reach the output
    const int compX0 = g.lofs, compX1 = std::min(W, g.lofs + g.width1);
    const int outX0 = need_bands ? compX0 : std::max(compX0, g.roiX0);
    const int outX1 = need_bands ? compX1 : std::min(compX1, g.roiX1);
    if (outX1 <= outX0) {
        if (nf > 1) return NOT_BATCHABLE;
        fill_s16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(disp, n, FILTERED);
        return launches + 1;
    }
    {
        const long long nfill = (long long)(outX0 + (W - outX1)) * H + (long long)(outX1 - outX0) * (g.roiY0 + (H - g.roiY1));
        if (nfill > 0 && !border_is_filled)
            fill_border_kernel<<<dim3((unsigned)((nfill + 255) / 256), nf), 256, 0, st>>>(disp, W, H, outX0, outX1, g.roiY0, g.roiY1,
                                                                                        FILTERED, disp_stride);
    }
    if (!border_is_filled) ++launches;
    if (evals) {
        int ex0 = std::max(compX0, g.roiX0), ex1 = std::min(compX1, g.roiX1);
        *evals = ex1 > ex0 ? (double)(ex1 - ex0) * (g.roiY1 - g.roiY0) * cfg.nd : 0.0;
    }
    // fast-path rectangle: clamp-free interior
    int XA = g.lofs + g.r, XB = W - g.r + std::min(cfg.minD, 0);
    XA = std::max(XA, outX0);
    XB = std::min(XB, outX1);
    bool fast_ok = g.rofs == 0 && (2 * cfg.cap * cfg.wsz * cfg.wsz < 65535) && XB > XA && (pitch % 16 == 0);
    static const int force_generic = getenv("B200S_FORCE_GENERIC") ? atoi(getenv("B200S_FORCE_GENERIC")) : 0;
    if (force_generic) fast_ok = false;
    const bool fast_ok_base = fast_ok;
    FastParams P;
    size_t smem = 0;
    int nt = 0;
    dim3 grid;
    if (fast_ok) {
        const int NK = cfg.nd / 16, NGH = cfg.nd / 8;
        static const int nt_target = getenv("B200S_NT") ? atoi(getenv("B200S_NT")) : 384;
        int NCQ = std::max(nt_target / NK, (2 * g.r + 8 + 3) / 4);
        NCQ = std::min(NCQ, (384 + 2 * g.r) / 4);  // one W thread per output column, at most 384 threads
        int TW = (4 * NCQ - 2 * g.r) & ~3;
        const int X0base = XA - ((XA - g.r - g.lofs) & 3);
        // do not make the tile wider than the work (keeps shared memory small for narrow images)
        int need = ((XB - X0base + 3) / 4) * 4;
        if (TW > need) {
            TW = need;
            NCQ = (TW + 2 * g.r + 3) / 4;
        }
        nt = ((std::max(NCQ * NK, TW) + 31) / 32) * 32;
        if (TW < 4 || nt > 384) fast_ok = false;
        else {
            static const int swd_min = getenv("B200S_SWD") ? atoi(getenv("B200S_SWD")) : 12;
            int NS = std::max(1, std::min(nt / NGH, (TW + swd_min - 1) / swd_min));
            int SWD = (((TW + NS - 1) / NS) + 1) & ~1;          // even strip width (the H loop is unrolled by 2)
            NS = (TW + SWD - 1) / SWD;
            const int rowsS = NS * SWD;
            const int ncols = 4 * NCQ;
            const int rowsC = std::max(ncols, rowsS + 2 * g.r + 2);
            P.Lp = Lp; P.Rp = Rp; P.pitch = pitch; P.disp = disp; P.cost = cost;
            P.W = W; P.H = H; P.nd = cfg.nd; P.minD = cfg.minD; P.r = g.r; P.cap = cfg.cap;
            P.texThr = cfg.textureThreshold; P.uniq = cfg.uniquenessRatio; P.lofs = g.lofs;
            P.X0base = X0base; P.XA = XA; P.XB = XB; P.YA = g.roiY0; P.YB = g.roiY1;
            P.TW = TW; P.ncols = ncols; P.NCQ = NCQ; P.NK = NK;
            P.NGH = NGH; P.NS = NS; P.SWD = SWD;
            P.CWb = cfg.nd * 2;
            P.SWb = cfg.nd * 2 + 16;
            P.NK4 = (NGH + 3) / 4;
            P.KWb = (P.NK4 * 4 + 4) * 4;
            P.RLW = (ncols + cfg.nd) / 4 + 1;
            int units = (4 * P.RLW + 15) / 16;
            while ((units & 3) != 2) ++units;
            P.CSB = units * 16;
            P.stage_from = (nt - TW >= 64) ? TW : 0;
            size_t o = 0;
            P.oLb = (int)o; o += 2 * (size_t)ncols * 4;
            o = (o + 15) & ~(size_t)15;
            P.oRc = (int)o; o += 2 * 4 * (size_t)P.CSB;
            P.oT = (int)o; o += (size_t)ncols * 4;
            o = (o + 15) & ~(size_t)15;
            P.oK = (int)o; o += (size_t)rowsS * P.KWb;
            P.oC = (int)o; o += (size_t)rowsC * P.CWb;
            P.oS = (int)o; o += (size_t)rowsS * P.SWb;
            smem = o;
            if (smem > 200 * 1024) fast_ok = false;
            // band height: about two blocks per SM in one wave; the column-sum bias (128 per row) bounds it
            int tilesX = (XB - X0base + TW - 1) / TW;
            int rows = g.roiY1 - g.roiY0;
            static const int want_blocks = getenv("B200S_BLOCKS") ? atoi(getenv("B200S_BLOCKS")) : 148 * 2;
            int bands = std::max(1, std::min(want_blocks / tilesX, std::max(1, rows / (4 * g.r + 8))));
            P.BH = (rows + bands - 1) / bands;
            int bh_max = 480 - 2 * g.r;
            if (P.BH > bh_max) P.BH = bh_max;
            grid = dim3(tilesX, (rows + P.BH - 1) / P.BH);
        }
    }
    GenPlanes gp{Lp, Rp, pitch};
    static const int kernel_sel = getenv("B200S_KERNEL") ? atoi(getenv("B200S_KERNEL")) : 7;
    bool ws_done = false;
    if (fast_ok_base && kernel_sel >= 7) {
        int rc = launch_bm_vh(Lp, Rp, pitch, W, H, cfg, g.r, g.lofs, XA, XB, g.roiY0, g.roiY1, disp, cost, st, nf, pre_stride,
                              disp_stride);
        if (rc < 0) return -1;
        ws_done = rc == 1;
    }
    // the two r-wide border strips of the L/R-check path: one fused launch (bm_strip.cu) when it applies
    bool strips_done = false;
    if (ws_done && XA - outX0 == g.r && outX1 - XB == g.r && g.rofs == 0) {
        static const int use_strips = getenv("B200S_STRIPS") ? atoi(getenv("B200S_STRIPS")) : 1;
        if (use_strips) {
            int rc = launch_bm_strips(Lp, Rp, pitch, W, H, cfg, g.r, g.lofs, outX0 - g.lofs, XB - g.lofs, g.roiY0, g.roiY1, disp, cost, st,
                                      nf, pre_stride, disp_stride);
            if (rc < 0) return -1;
            strips_done = rc == 1;
        }
    }
    if (strips_done) return launches + 2;
    if (nf > 1) {
        // a batch needs bm_vh (the border fill above is harmless to repeat frame by frame otherwise); strips the fused
        // kernel does not handle go through the generic kernels frame by frame
        if (!ws_done) return NOT_BATCHABLE;
        ++launches;
        for (int f = 0; f < nf; ++f) {
            GenPlanes gf{Lp + f * pre_stride, Rp + f * pre_stride, pitch};
            int l = run_generic_pair(gf, W, H, cfg, g, outX0 - g.lofs, XA - g.lofs, XB - g.lofs, outX1 - g.lofs,
                                     (int16_t*)((uint8_t*)disp + f * disp_stride),
                                     cost ? (int16_t*)((uint8_t*)cost + f * disp_stride) : nullptr, sc, st);
            if (l < 0) return l;
            launches += l;
        }
        return launches;
    }
    if (!ws_done && fast_ok_base && kernel_sel >= 4) {
        int rc = launch_bm_ws(Lp, Rp, pitch, W, H, cfg, g.r, g.lofs, XA, XB, g.roiY0, g.roiY1, disp, cost, st);
        if (rc < 0) return -1;
        ws_done = rc == 1;
    }
    if (ws_done) {
        ++launches;
        int l = run_generic_pair(gp, W, H, cfg, g, outX0 - g.lofs, XA - g.lofs, XB - g.lofs, outX1 - g.lofs, disp, cost, sc, st);
        if (l < 0) return l;
        launches += l;
    } else if (fast_ok) {
        cudaError_t e;
        if (cfg.nd == 256) e = launch_fast<256>(P, grid, nt, smem, st);
        else if (cfg.nd == 128) e = launch_fast<128>(P, grid, nt, smem, st);
        else if (cfg.nd == 64) e = launch_fast<64>(P, grid, nt, smem, st);
        else e = launch_fast<0>(P, grid, nt, smem, st);
        if (e != cudaSuccess) return -1;
        ++launches;
        // border bands that the fast kernel does not cover
        int l = run_generic(gp, W, H, cfg, g, outX0 - g.lofs, XA - g.lofs, disp, cost, sc, st);
        if (l < 0) return l;
        launches += l;
        l = run_generic(gp, W, H, cfg, g, XB - g.lofs, outX1 - g.lofs, disp, cost, sc, st);
        if (l < 0) return l;
        launches += l;
    } else {
        int l = run_generic(gp, W, H, cfg, g, outX0 - g.lofs, outX1 - g.lofs, disp, cost, sc, st);
        if (l < 0) return l;
        launches += l;
    }
    return launches;
}

int launch_block_match(const uint8_t* Lp, const uint8_t* Rp, size_t pitch, int W, int H, const BMConfig& cfg,
                       int16_t* disp, int16_t* cost, BMScratch* sc, cudaStream_t st, double* evals, int nf, size_t pre_stride,
                       size_t disp_stride, bool border_is_filled)
{
    if (nf <= 1) return block_match_impl(Lp, Rp, pitch, W, H, cfg, disp, cost, sc, st, evals, 1, 0, 0, border_is_filled);
    int rc = block_match_impl(Lp, Rp, pitch, W, H, cfg, disp, cost, sc, st, evals, nf, pre_stride, disp_stride, border_is_filled);
    if (rc != NOT_BATCHABLE) return rc;
    int total = 0;
    for (int f = 0; f < nf; ++f) {
        rc = block_match_impl(Lp + f * pre_stride, Rp + f * pre_stride, pitch, W, H, cfg, (int16_t*)((uint8_t*)disp + f * disp_stride),
                              cost ? (int16_t*)((uint8_t*)cost + f * disp_stride) : nullptr, sc, st, evals, 1, 0, 0, false);
        if (rc < 0) return rc;
        total += rc;
    }
    return total;
}

}  // namespace b200s
