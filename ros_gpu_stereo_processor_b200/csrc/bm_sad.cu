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

#include <algorithm>
#include <climits>
#include <cstdlib>

namespace b200s {

// ------------------------------------------------------------------------------------------------------
// shared winner arithmetic (A.2.5)
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int16_t subpixel_disp(int minsad, int mind, int p, int n, int nd, int minD)
{
    int d = p + n - 2 * minsad + abs(p - n);
    int v = ((nd - mind - 1 + minD) * 256 + (d != 0 ? (p - n) * 256 / d : 0) + 15) >> 4;
    return (int16_t)v;
}

__global__ void fill_s16_kernel(int16_t* p, size_t n, int16_t v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ------------------------------------------------------------------------------------------------------
// fast kernel (v2)
//
// One block = TW output columns x BH output rows, all nd disparities; it marches down the rows.  Per row:
//   V  thread = (4 adjacent window columns, 16 disparities): abs-diff of the entering and the leaving row
//      (VABSDIFF4.U8 on 4 disparities at a time; the right-image window of column i is the thread's 20-byte
//      register window funnel-shifted by i bytes), biased byte delta, widened to u16x2 lanes and added to the
//      32 column-sum registers; the sums go to shared memory (Cbuf, 16-byte units XOR-swizzled by column).
//   H  thread = (8 disparities, strip of columns): horizontal sliding sum over the 2r+1 window columns -> Sbuf.
//   W  LPP lanes per pixel: packed u16x2 minima per 16 disparities -> 32-bit key (sad << 16 | chunk) -> argmin;
//      exact index inside the winning chunk; uniqueness by a second packed-min pass with the winner and its two
//      neighbours masked; texture; sub-pixel fit.  While W runs, the next two image rows are staged.
// Shared memory per row pass: right rows as 4 word-shifted copies (so that every 20-byte window is one
// aligned LDS.128 + LDS.32), left rows pre-broadcast to 4 bytes.
// ------------------------------------------------------------------------------------------------------
struct FastParams {
    const uint8_t* Lp;    // prefiltered planes, row pitch `pitch` (multiple of 16), with slack before/after
    const uint8_t* Rp;
    size_t pitch;
    int16_t* disp;        // tightly packed W
    int16_t* cost;        // may be null
    int W, H, nd, minD, r, cap, texThr, uniq, lofs;
    int X0base, XB, YA, YB;   // first tile origin (<= XA, aligned), end of output columns, output rows
    int XA;
    int TW, BH, ncols;        // tile width, band height, window columns (multiple of 4)
    int NCQ, NKG;             // V items: column quads x 16-disparity groups
    int NGH, NS, SWD;         // H items: 8-disparity groups x strips of SWD columns
    int CW;                   // words per Cbuf/Sbuf row (multiple of 32)
    int CSB, RLW;             // bytes per right-row copy, words staged per right row
    int oLb, oRc, oC, oS, oT; // shared memory byte offsets
};

// position of disparity index k inside a group of four u16 lanes (V-phase lane order is k, k+2, k+1, k+3)
__device__ __forceinline__ int kpos(int k) { return (k & ~3) | ((k & 1) << 1) | ((k >> 1) & 1); }
// XOR swizzle of the 16-byte unit index inside an Sbuf row (conflict-free for 1, 2 and 4 lanes per pixel)
__device__ __forceinline__ int swzS(int x) { return (x & 1) | ((x & 2) << 1) | ((x & 4) >> 1); }
__device__ __forceinline__ int swzC(int c) { return (c >> 2) & 7; }

__device__ __forceinline__ void stage_rows(const FastParams& P, uint8_t* smem, int tid, int NT, int yi, bool has_old,
                                           int Xl0, int Xr0)
{
    uint32_t* sLb = (uint32_t*)(smem + P.oLb);
    uint8_t* sRc = smem + P.oRc;
    const int b = 2 * P.r + 1;
    const uint8_t* ln = P.Lp + (size_t)yi * P.pitch + Xl0;
    const uint8_t* lo = P.Lp + (size_t)max(yi - b, 0) * P.pitch + Xl0;
    for (int c = tid; c < P.ncols; c += NT) {
        sLb[c] = (uint32_t)__ldg(ln + c) * 0x01010101u;
        sLb[P.ncols + c] = has_old ? (uint32_t)__ldg(lo + c) * 0x01010101u : 0u;
    }
    const uint32_t* rn = (const uint32_t*)(P.Rp + (size_t)yi * P.pitch + Xr0);          // Xr0 % 4 == 0, pitch % 16 == 0
    const uint32_t* ro = (const uint32_t*)(P.Rp + (size_t)max(yi - b, 0) * P.pitch + Xr0);
    for (int i = tid; i < 2 * P.RLW; i += NT) {
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

template <int LPP>
__global__ void __launch_bounds__(384, 2) bm_fast_kernel(const FastParams P)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t* sLb = (const uint32_t*)(smem + P.oLb);   // [2][ncols]  left bytes x 0x01010101
    const uint8_t* sRc = smem + P.oRc;                       // [2][4][CSB] right row, copy j shifted by 4j bytes
    uint32_t* sC = (uint32_t*)(smem + P.oC);                 // [ncols][CW] column sums, u16x2
    uint32_t* sS = (uint32_t*)(smem + P.oS);                 // [TW][CW]    window sums, u16x2
    uint32_t* sT = (uint32_t*)(smem + P.oT);                 // [ncols]     texture column sums

    const int tid = threadIdx.x, NT = blockDim.x;
    const int X0 = P.X0base + blockIdx.x * P.TW;             // first output column of the tile
    const int yb0 = P.YA + blockIdx.y * P.BH;
    const int yb1 = min(yb0 + P.BH, P.YB);
    const int r = P.r, b = 2 * r + 1, nd = P.nd;
    const int Xl0 = X0 - r;               // left image column of window column c = 0
    const int Xr0 = X0 - r - P.lofs;      // right image column of (c = 0, k = 0); multiple of 4 by construction

    // V identity: lanes run over column quads first (bank-conflict-free stores with the column swizzle)
    const int cq = tid % P.NCQ, kg = tid / P.NCQ;
    const bool vact = kg < P.NKG;
    uint32_t C[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int w = 0; w < 8; ++w) C[i][w] = 0;
    for (int c = tid; c < P.ncols; c += NT) sT[c] = 0;
    // H identity
    const int gh = tid % P.NGH, hs = tid / P.NGH;
    const bool hact = hs < P.NS;

    stage_rows(P, smem, tid, NT, yb0 - r, false, Xl0, Xr0);
    __syncthreads();

    int nrow = 0;   // rows accumulated so far (bias bookkeeping: every row adds 128 per u16 lane)
    for (int yi = yb0 - r; yi < yb1 + r; ++yi) {
        const bool has_old = (yi - b) >= yb0 - r;
        const bool do_out = yi >= yb0 + r;
        // ---- phase V ------------------------------------------------------------------------------------
        if (vact) {
            const int j = cq & 3;
            const int A = 4 * cq + 16 * kg;
            const uint8_t* pn = sRc + (size_t)j * P.CSB + (A - 4 * j);
            const uint8_t* po = pn + (size_t)4 * P.CSB;
            const uint4 rn4 = *(const uint4*)pn;
            const uint32_t rn5 = *(const uint32_t*)(pn + 16);
            const uint4 ro4 = *(const uint4*)po;
            const uint32_t ro5 = *(const uint32_t*)(po + 16);
            const uint4 ln4 = *(const uint4*)(sLb + 4 * cq);
            const uint4 lo4 = *(const uint4*)(sLb + P.ncols + 4 * cq);
            const uint32_t rn[5] = {rn4.x, rn4.y, rn4.z, rn4.w, rn5};
            const uint32_t ro[5] = {ro4.x, ro4.y, ro4.z, ro4.w, ro5};
            const uint32_t ln[4] = {ln4.x, ln4.y, ln4.z, ln4.w};
            const uint32_t lo[4] = {lo4.x, lo4.y, lo4.z, lo4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const uint32_t wn = i ? __funnelshift_r(rn[w], rn[w + 1], 8 * i) : rn[w];
                    const uint32_t wo = i ? __funnelshift_r(ro[w], ro[w + 1], 8 * i) : ro[w];
                    const uint32_t an = __vabsdiffu4(ln[i], wn);
                    const uint32_t ao = __vabsdiffu4(lo[i], wo);
                    const uint32_t t = an + 0x80808080u - ao;        // per byte: 128 + new - old, no borrow
                    C[i][2 * w] += t & 0x00ff00ffu;                  // lanes k+0, k+2   (bias 128 per lane kept)
                    C[i][2 * w + 1] += __byte_perm(t, 0, 0x4341);    // lanes k+1, k+3
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = 4 * cq + i;
                uint32_t* row = sC + (size_t)c * P.CW;
                const int sw = swzC(c);
                *(uint4*)(row + 4 * ((2 * kg) ^ sw)) = make_uint4(C[i][0], C[i][1], C[i][2], C[i][3]);
                *(uint4*)(row + 4 * ((2 * kg + 1) ^ sw)) = make_uint4(C[i][4], C[i][5], C[i][6], C[i][7]);
            }
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
            const int xs = hs * P.SWD, xe = min(P.TW, xs + P.SWD);
            if (xs < xe) {
                // every column sum carries a bias of 128 * nrow per lane; remove b of them from the window sum
                const uint32_t bias = (uint32_t)(128 * nrow * b) * 0x00010001u;
                uint4 S = make_uint4(0u - bias, 0u - bias, 0u - bias, 0u - bias);
                for (int c = xs; c < xs + b; ++c) {
                    const uint4 v = *(const uint4*)(sC + (size_t)c * P.CW + 4 * (gh ^ swzC(c)));
                    S.x += v.x; S.y += v.y; S.z += v.z; S.w += v.w;
                }
                for (int x = xs; x < xe; ++x) {
                    *(uint4*)(sS + (size_t)x * P.CW + 4 * (gh ^ swzS(x))) = S;
                    if (x + 1 < xe) {
                        const int ca = x + b;
                        const uint4 a = *(const uint4*)(sC + (size_t)ca * P.CW + 4 * (gh ^ swzC(ca)));
                        const uint4 o = *(const uint4*)(sC + (size_t)x * P.CW + 4 * (gh ^ swzC(x)));
                        S.x += a.x - o.x; S.y += a.y - o.y; S.z += a.z - o.z; S.w += a.w - o.w;
                    }
                }
            }
        }
        __syncthreads();
        // ---- phase W (+ staging of the next rows) ---------------------------------------------------------
        if (yi + 1 < yb1 + r) stage_rows(P, smem, tid, NT, yi + 1, (yi + 1 - b) >= yb0 - r, Xl0, Xr0);
        {
            const int y = yi - r;
            const int16_t FILTERED = (int16_t)((P.minD - 1) * 16);
            const int q = tid & (LPP - 1);
            const int npair = nd >> 4;                 // 16-disparity chunks per pixel
            for (int base = 0; base < P.TW; base += NT / LPP) {
                const int px = base + tid / LPP;
                const int x = min(px, P.TW - 1);
                uint32_t* srow = sS + (size_t)x * P.CW;
                const int sw = swzS(x);
                // pass A: per 16 disparities a packed minimum, then key = sad << 16 | chunk (lowest chunk wins ties)
                uint32_t best = 0xFFFFFFFFu;
                for (int pc = q; pc < npair; pc += LPP) {
                    const uint4 u0 = *(const uint4*)(srow + 4 * ((2 * pc) ^ sw));
                    const uint4 u1 = *(const uint4*)(srow + 4 * ((2 * pc + 1) ^ sw));
                    uint32_t m = __vimin3_u16x2(u0.x, u0.y, u0.z);
                    m = __vimin3_u16x2(m, u0.w, u1.x);
                    m = __vimin3_u16x2(m, u1.y, u1.z);
                    m = __vminu2(m, u1.w);
                    m = __vminu2(m, m >> 16);
                    best = min(best, (m << 16) | (uint32_t)pc);
                }
                if (LPP >= 2) best = min(best, __shfl_xor_sync(0xffffffffu, best, 1));
                if (LPP >= 4) best = min(best, __shfl_xor_sync(0xffffffffu, best, 2));
                const int minsad = (int)(best >> 16), pcs = (int)(best & 0xffffu);
                // exact index inside the winning chunk, in disparity-index order (lowest k wins ties)
                int mind;
                {
                    const uint4 u0 = *(const uint4*)(srow + 4 * ((2 * pcs) ^ sw));
                    const uint4 u1 = *(const uint4*)(srow + 4 * ((2 * pcs + 1) ^ sw));
                    const uint32_t wv[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
                    int loc = 15;
#pragma unroll
                    for (int kk = 15; kk >= 0; --kk) {
                        // word of k-offset kk: group g4 = kk / 4, inside: (0: w0.lo, 1: w1.lo, 2: w0.hi, 3: w1.hi)
                        const uint32_t w = wv[2 * (kk >> 2) + (kk & 1)];
                        const uint32_t v = (kk & 2) ? (w >> 16) : (w & 0xffffu);
                        if ((int)v == minsad) loc = kk;
                    }
                    mind = 16 * pcs + loc;
                }
                int pv = 0, nv = 0;
                const uint16_t* s16c = (const uint16_t*)srow;
                if (q == 0) {
                    const int kp = kpos(mind + 1 < nd ? mind + 1 : nd - 2), kn = kpos(mind > 0 ? mind - 1 : 1);
                    pv = s16c[8 * ((kp >> 3) ^ sw) + (kp & 7)];
                    nv = s16c[8 * ((kn >> 3) ^ sw) + (kn & 7)];
                }
                bool filtered = false;
                if (P.uniq > 0) {
                    if (LPP > 1) __syncwarp();
                    if (q == 0 && px < P.TW) {   // inactive lanes alias the last row and must not patch it
                        uint16_t* s16 = (uint16_t*)srow;
                        int kp = kpos(mind);
                        s16[8 * ((kp >> 3) ^ sw) + (kp & 7)] = 0xFFFFu;
                        if (mind > 0) { kp = kpos(mind - 1); s16[8 * ((kp >> 3) ^ sw) + (kp & 7)] = 0xFFFFu; }
                        if (mind + 1 < nd) { kp = kpos(mind + 1); s16[8 * ((kp >> 3) ^ sw) + (kp & 7)] = 0xFFFFu; }
                    }
                    if (LPP > 1) __syncwarp();
                    uint32_t acc = 0xFFFFFFFFu;
                    for (int pc = q; pc < npair; pc += LPP) {
                        const uint4 u0 = *(const uint4*)(srow + 4 * ((2 * pc) ^ sw));
                        const uint4 u1 = *(const uint4*)(srow + 4 * ((2 * pc + 1) ^ sw));
                        acc = __vimin3_u16x2(acc, u0.x, u0.y);
                        acc = __vimin3_u16x2(acc, u0.z, u0.w);
                        acc = __vimin3_u16x2(acc, u1.x, u1.y);
                        acc = __vimin3_u16x2(acc, u1.z, u1.w);
                    }
                    uint32_t m2 = min(acc & 0xffffu, acc >> 16);
                    if (LPP >= 2) m2 = min(m2, __shfl_xor_sync(0xffffffffu, m2, 1));
                    if (LPP >= 4) m2 = min(m2, __shfl_xor_sync(0xffffffffu, m2, 2));
                    const int thr = minsad + (minsad * P.uniq / 100);
                    filtered = (int)m2 <= thr;
                }
                if (q == 0 && px < P.TW) {
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
    P.RCH = 32;
    for (int ya = g.roiY0; ya < g.roiY1; ya += rows_per_chunk) {
        P.ya = ya;
        P.yb = std::min(g.roiY1, ya + rows_per_chunk);
        int kt = std::min(128, cfg.nd);
        dim3 grid((cfg.nd + kt - 1) / kt, ncx, (P.yb - P.ya + P.RCH - 1) / P.RCH);
        // gridDim.y/z limits (65535) are far above any supported image size
        bm_generic_cost_kernel<<<grid, kt, 0, st>>>(P);
        int npx = ncx * (P.yb - P.ya);
        bm_generic_winner_kernel<<<(npx + 127) / 128, 128, 0, st>>>(P);
        launches += 2;
    }
    return launches;
}

template <int LPP>
static cudaError_t launch_fast(const FastParams& P, dim3 grid, int nt, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(bm_fast_kernel<LPP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    bm_fast_kernel<LPP><<<grid, nt, smem, st>>>(P);
    return cudaGetLastError();
}

int launch_block_match(const uint8_t* Lp, const uint8_t* Rp, size_t pitch, int W, int H, const BMConfig& cfg,
                       int16_t* disp, int16_t* cost, BMScratch* sc, cudaStream_t st, double* evals)
{
    const Geom g = geom(W, H, cfg);
    const int16_t FILTERED = (int16_t)((cfg.minD - 1) * 16);
    int launches = 0;
    size_t n = (size_t)W * H;
    fill_s16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(disp, n, FILTERED);
    ++launches;
    if (cost) { cudaMemsetAsync(cost, 0, n * sizeof(int16_t), st); }
    if (evals) *evals = 0;
    if (g.degenerate) return launches;
    const bool need_bands = cfg.disp12MaxDiff >= 0;
    // columns (in X) the matcher produces at all, and the ones that can reach the output
    const int compX0 = g.lofs, compX1 = std::min(W, g.lofs + g.width1);
    const int outX0 = need_bands ? compX0 : std::max(compX0, g.roiX0);
    const int outX1 = need_bands ? compX1 : std::min(compX1, g.roiX1);
    if (outX1 <= outX0) return launches;
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
    FastParams P;
    size_t smem = 0;
    int nt = 0, LPP = 1;
    dim3 grid;
    if (fast_ok) {
        const int NKG = cfg.nd / 16;
        static const int nt_target = getenv("B200S_NT") ? atoi(getenv("B200S_NT")) : 320;
        int NCQ = std::max(nt_target / NKG, (2 * g.r + 8 + 3) / 4);
        NCQ = std::min(NCQ, 128);                 // at most 512 window columns per tile
        int ncols = 4 * NCQ;
        int TW = (ncols - 2 * g.r) & ~3;
        nt = ((NCQ * NKG + 31) / 32) * 32;
        if (TW < 4 || nt > 384) fast_ok = false;
        else {
            const int X0base = XA - ((XA - g.r - g.lofs) & 3);
            // do not make the tile wider than the work (keeps shared memory small for narrow images)
            int need = ((XB - X0base + 3) / 4) * 4;
            if (TW > need) {
                TW = need;
                NCQ = (TW + 2 * g.r + 3) / 4;
                ncols = 4 * NCQ;
                nt = ((NCQ * NKG + 31) / 32) * 32;
            }
            P.Lp = Lp; P.Rp = Rp; P.pitch = pitch; P.disp = disp; P.cost = cost;
            P.W = W; P.H = H; P.nd = cfg.nd; P.minD = cfg.minD; P.r = g.r; P.cap = cfg.cap;
            P.texThr = cfg.textureThreshold; P.uniq = cfg.uniquenessRatio; P.lofs = g.lofs;
            P.X0base = X0base; P.XA = XA; P.XB = XB; P.YA = g.roiY0; P.YB = g.roiY1;
            P.TW = TW; P.ncols = ncols; P.NCQ = NCQ; P.NKG = NKG;
            P.NGH = cfg.nd / 8;
            static const int swd_min = getenv("B200S_SWD") ? atoi(getenv("B200S_SWD")) : 12;
            P.NS = std::max(1, std::min(nt / P.NGH, (TW + swd_min - 1) / swd_min));
            P.SWD = (TW + P.NS - 1) / P.NS;
            P.CW = ((cfg.nd / 2 + 31) / 32) * 32;
            P.RLW = (ncols + cfg.nd) / 4 + 1;
            int units = (4 * P.RLW + 15) / 16;
            while ((units & 7) != 2) ++units;
            P.CSB = units * 16;
            size_t o = 0;
            P.oLb = (int)o; o += 2 * (size_t)ncols * 4;
            o = (o + 15) & ~(size_t)15;
            P.oRc = (int)o; o += 2 * 4 * (size_t)P.CSB;
            P.oC = (int)o; o += (size_t)ncols * P.CW * 4;
            P.oS = (int)o; o += (size_t)TW * P.CW * 4;
            P.oT = (int)o; o += (size_t)ncols * 4;
            smem = o;
            if (smem > 200 * 1024) fast_ok = false;
            static const int lpp_env = getenv("B200S_LPP") ? atoi(getenv("B200S_LPP")) : 0;
            LPP = (cfg.nd % 64 == 0 && 4 * TW <= nt + nt / 2) ? 4 : ((cfg.nd % 32 == 0 && 2 * TW <= nt + nt / 2) ? 2 : 1);
            if (lpp_env == 1 || (lpp_env == 2 && cfg.nd % 32 == 0) || (lpp_env == 4 && cfg.nd % 64 == 0)) LPP = lpp_env;
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
    if (fast_ok) {
        cudaError_t e;
        if (LPP == 4) e = launch_fast<4>(P, grid, nt, smem, st);
        else if (LPP == 2) e = launch_fast<2>(P, grid, nt, smem, st);
        else e = launch_fast<1>(P, grid, nt, smem, st);
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

}  // namespace b200s
