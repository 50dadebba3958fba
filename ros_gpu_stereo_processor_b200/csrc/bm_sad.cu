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
// fast kernel
// ------------------------------------------------------------------------------------------------------
struct FastParams {
    const uint8_t* Lp;
    const uint8_t* Rp;
    int16_t* disp;
    int16_t* cost;  // may be null
    int W, H, nd, minD, r, cap, texThr, uniq, lofs;
    int XA, XB, YA, YB;   // output rectangle handled by the fast kernel
    int TW, BH, ncols;    // tile width, band height, TW + 2r
    int NG;               // V-thread groups per column (nd / KPT)
    int NGH, NS, SWD;     // H-phase: k-groups of 8, strips, strip width
    int CW;               // words per column row in Cbuf/Sbuf (nd/2 + 4)
    int RL, CS;           // right row segment length, copy stride (bytes)
    // shared memory byte offsets
    int oL, oRb, oRc, oC, oS, oT;
    int RbS;              // bytes per Rbase row
    int LS;               // bytes per L row
};

// position of disparity index k inside a group of four u16 lanes (V-phase lane order is k, k+2, k+1, k+3)
__device__ __forceinline__ int kpos(int k) { return (k & ~3) | ((k & 1) << 1) | ((k >> 1) & 1); }

template <int KPT>
__global__ void __launch_bounds__(384, 2) bm_fast_kernel(const FastParams P)
{
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* sL = smem + P.oL;                       // [2][LS]
    uint8_t* sRb = smem + P.oRb;                     // [2][RbS]
    uint8_t* sRc = smem + P.oRc;                     // [2][16][CS]
    uint32_t* sC = (uint32_t*)(smem + P.oC);         // [ncols][CW]
    uint32_t* sS = (uint32_t*)(smem + P.oS);         // [TW][CW]
    uint32_t* sT = (uint32_t*)(smem + P.oT);         // [ncols]

    const int tid = threadIdx.x, NT = blockDim.x;
    const int X0 = P.XA + blockIdx.x * P.TW;         // first output column of the tile
    const int yb0 = P.YA + blockIdx.y * P.BH;
    const int yb1 = min(yb0 + P.BH, P.YB);
    const int r = P.r, b = 2 * r + 1, nd = P.nd, ncols = P.ncols;

    // V-phase identity
    const int vc = tid % ncols, vg = tid / ncols;
    const bool vact = vg < P.NG;
    uint32_t C[KPT / 2];
#pragma unroll
    for (int i = 0; i < KPT / 2; ++i) C[i] = 0;
    uint32_t tcol = 0;

    const int Xl0 = X0 - r;               // left image column of window column c = 0
    const int Xr0 = X0 - r - P.lofs;      // right image column of (c = 0, k = 0)

    for (int yi = yb0 - r; yi < yb1 + r; ++yi) {
        const int yo = yi - b;
        const bool has_old = yo >= yb0 - r;
        const bool do_out = yi >= yb0 + r;
        // ---- phase 0: stage the new and the old row ------------------------------------------------
        {
            const uint8_t* ln = P.Lp + (size_t)yi * P.W;
            const uint8_t* rn = P.Rp + (size_t)yi * P.W;
            const uint8_t* lo = P.Lp + (size_t)max(yo, 0) * P.W;
            const uint8_t* ro = P.Rp + (size_t)max(yo, 0) * P.W;
            for (int i = tid; i < P.LS; i += NT) {
                int X = min(max(Xl0 + i, 0), P.W - 1);
                sL[i] = __ldg(ln + X);
                sL[P.LS + i] = has_old ? __ldg(lo + X) : (uint8_t)0;
            }
            for (int i = tid; i < P.RbS; i += NT) {
                int X = min(max(Xr0 + i, 0), P.W - 1);
                sRb[i] = __ldg(rn + X);
                sRb[P.RbS + i] = has_old ? __ldg(ro + X) : (uint8_t)0;
            }
        }
        __syncthreads();
        {   // 16 byte-shifted copies of each row: copy[m][j] = base[j + m]
            const int nq = P.CS >> 4;
            const int items = 2 * 16 * nq;
            for (int it = tid; it < items; it += NT) {
                int s = it / (16 * nq);
                int rem = it - s * 16 * nq;
                int m = rem / nq, q = rem - m * nq;
                int boff = 16 * q + m;
                const uint32_t* w = (const uint32_t*)(sRb + s * P.RbS + (boff & ~3));
                int sh = (boff & 3) * 8;
                uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
                uint4 o;
                o.x = __funnelshift_r(w0, w1, sh);
                o.y = __funnelshift_r(w1, w2, sh);
                o.z = __funnelshift_r(w2, w3, sh);
                o.w = __funnelshift_r(w3, w4, sh);
                *(uint4*)(sRc + ((size_t)(s * 16 + m) * P.CS) + 16 * q) = o;
            }
        }
        __syncthreads();
        // ---- phase V: vertical sliding column sums --------------------------------------------------
        if (vact) {
            const int k0 = vg * KPT;
            const int s0 = vc + k0;
            const int m = s0 & 15, off = s0 - m;
            const uint8_t* pn = sRc + (size_t)m * P.CS + off;
            const uint8_t* po = sRc + (size_t)(16 + m) * P.CS + off;
            const uint32_t ln = sL[vc], lo = sL[P.LS + vc];
            const uint32_t Ln4 = ln * 0x01010101u, Lo4 = lo * 0x01010101u;
#pragma unroll
            for (int q = 0; q < KPT / 16; ++q) {
                const uint4 rn = *(const uint4*)(pn + 16 * q);
                const uint4 ro = *(const uint4*)(po + 16 * q);
                const uint32_t rnw[4] = {rn.x, rn.y, rn.z, rn.w};
                const uint32_t row[4] = {ro.x, ro.y, ro.z, ro.w};
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    uint32_t an = __vabsdiffu4(Ln4, rnw[w]);
                    uint32_t ao = __vabsdiffu4(Lo4, row[w]);
                    uint32_t t = an + 0x80808080u - ao;          // per byte: 128 + new - old, no borrow
                    uint32_t e = t & 0x00ff00ffu;                // lanes k+0, k+2
                    uint32_t o = __byte_perm(t, 0, 0x4341);      // lanes k+1, k+3
                    C[8 * q + 2 * w] += e - 0x00800080u;
                    C[8 * q + 2 * w + 1] += o - 0x00800080u;
                }
            }
            uint32_t* dst = sC + (size_t)vc * P.CW + (k0 >> 1);
#pragma unroll
            for (int i = 0; i < KPT / 8; ++i)
                *(uint4*)(dst + 4 * i) = make_uint4(C[4 * i], C[4 * i + 1], C[4 * i + 2], C[4 * i + 3]);
            if (vg == 0) {
                int cap = P.cap;
                tcol += abs((int)ln - cap) - (has_old ? abs((int)lo - cap) : 0);
                sT[vc] = tcol;
            }
        }
        __syncthreads();
        if (!do_out) continue;   // uniform across the block
        // ---- phase H: horizontal sliding sums ---------------------------------------------------------
        if (tid < P.NGH * P.NS) {
            const int gh = tid % P.NGH, s = tid / P.NGH;
            const int xs = s * P.SWD, xe = min(P.TW, xs + P.SWD);
            if (xs < xe) {
                const uint32_t* cb = sC + 4 * gh;
                uint4 S = make_uint4(0, 0, 0, 0);
                for (int c = xs; c < xs + b; ++c) {
                    uint4 v = *(const uint4*)(cb + (size_t)c * P.CW);
                    S.x += v.x; S.y += v.y; S.z += v.z; S.w += v.w;
                }
                for (int x = xs; x < xe; ++x) {
                    *(uint4*)(sS + (size_t)x * P.CW + 4 * gh) = S;
                    if (x + 1 < xe) {
                        uint4 a = *(const uint4*)(cb + (size_t)(x + b) * P.CW);
                        uint4 o = *(const uint4*)(cb + (size_t)x * P.CW);
                        S.x += a.x - o.x; S.y += a.y - o.y; S.z += a.z - o.z; S.w += a.w - o.w;
                    }
                }
            }
        }
        __syncthreads();
        // ---- phase W: winner per pixel, 4 lanes per pixel ---------------------------------------------
        {
            const int y = yi - r;
            const int16_t FILTERED = (int16_t)((P.minD - 1) * 16);
            const int q4 = tid & 3;
            const int per = nd >> 2;                 // positions per lane (multiple of 4)
            for (int base = 0; base < P.TW; base += (NT >> 2)) {
                const int px = base + (tid >> 2);
                const int x = min(px, P.TW - 1);
                uint32_t* srow = sS + (size_t)x * P.CW;
                const uint2* sp = (const uint2*)srow + (q4 * per >> 2);
                // pass A: argmin with key = sad << 12 | k  (lowest k wins ties)
                uint32_t best = 0xFFFFFFFFu;
                const int kb = q4 * per;
                for (int j = 0; j < (per >> 2); ++j) {
                    uint2 v = sp[j];
                    int k = kb + 4 * j;
                    uint32_t k0v = ((v.x & 0xffffu) << 12) | (uint32_t)(k);
                    uint32_t k2v = ((v.x >> 16) << 12) | (uint32_t)(k + 2);
                    uint32_t k1v = ((v.y & 0xffffu) << 12) | (uint32_t)(k + 1);
                    uint32_t k3v = ((v.y >> 16) << 12) | (uint32_t)(k + 3);
                    best = min(best, min(min(k0v, k1v), min(k2v, k3v)));
                }
                best = min(best, __shfl_xor_sync(0xffffffffu, best, 1));
                best = min(best, __shfl_xor_sync(0xffffffffu, best, 2));
                const int minsad = (int)(best >> 12), mind = (int)(best & 0xfffu);
                int pv = 0, nv = 0;
                if (q4 == 0) {
                    const uint16_t* s16 = (const uint16_t*)srow;
                    pv = s16[kpos(mind + 1 < nd ? mind + 1 : nd - 2)];
                    nv = s16[kpos(mind > 0 ? mind - 1 : 1)];
                }
                bool filtered = false;
                if (P.uniq > 0) {
                    __syncwarp();
                    if (q4 == 0 && px < P.TW) {   // inactive quads alias the last row and must not patch it
                        uint16_t* s16 = (uint16_t*)srow;
                        s16[kpos(mind)] = 0xFFFFu;
                        if (mind > 0) s16[kpos(mind - 1)] = 0xFFFFu;
                        if (mind + 1 < nd) s16[kpos(mind + 1)] = 0xFFFFu;
                    }
                    __syncwarp();
                    uint32_t acc = 0xFFFFFFFFu;
                    for (int j = 0; j < (per >> 2); ++j) {
                        uint2 v = sp[j];
                        acc = __vimin3_u16x2(acc, v.x, v.y);
                    }
                    uint32_t m2 = min(acc & 0xffffu, acc >> 16);
                    m2 = min(m2, __shfl_xor_sync(0xffffffffu, m2, 1));
                    m2 = min(m2, __shfl_xor_sync(0xffffffffu, m2, 2));
                    int thr = minsad + (minsad * P.uniq / 100);
                    filtered = (int)m2 <= thr;
                }
                if (q4 == 0 && px < P.TW) {
                    const int X = X0 + px;
                    if (X < P.XB) {
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
        // the next iteration's phase 0 only touches sL/sRb, which phase V (behind a barrier) is done with;
        // sS is rewritten only after two more barriers.
    }
}

// ------------------------------------------------------------------------------------------------------
// generic path: exact clamped semantics, int32 sums, cost volume in global scratch
// ------------------------------------------------------------------------------------------------------
struct GenParams {
    const uint8_t* Lp;
    const uint8_t* Rp;
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
    const uint8_t* lr = P.Lp + (size_t)y * P.W;
    const uint8_t* rr = P.Rp + (size_t)y * P.W;
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
        const uint8_t* lr = P.Lp + (size_t)(y + dy) * P.W;
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

static int run_generic(const uint8_t* Lp, const uint8_t* Rp, int W, int H, const BMConfig& cfg, const Geom& g,
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
    P.Lp = Lp; P.Rp = Rp; P.W = W; P.H = H; P.nd = cfg.nd; P.minD = cfg.minD; P.r = g.r; P.cap = cfg.cap;
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

template <int KPT>
static cudaError_t launch_fast(const FastParams& P, dim3 grid, int nt, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(bm_fast_kernel<KPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    bm_fast_kernel<KPT><<<grid, nt, smem, st>>>(P);
    return cudaGetLastError();
}

int launch_block_match(const uint8_t* Lp, const uint8_t* Rp, int W, int H, const BMConfig& cfg, int16_t* disp,
                       int16_t* cost, BMScratch* sc, cudaStream_t st, double* evals)
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
    bool fast_ok = g.rofs == 0 && (2 * cfg.cap * cfg.wsz * cfg.wsz < 65535) && cfg.nd <= 1024 && XB > XA;
    FastParams P;
    size_t smem = 0;
    int nt = 0, KPT = 0;
    dim3 grid;
    if (fast_ok) {
        KPT = (cfg.nd % 64 == 0) ? 64 : (cfg.nd % 32 == 0 ? 32 : 16);
        int NG = cfg.nd / KPT;
        // threads: one per (window column, k-group); aim for ~288 threads, tile width >= 16
        int ncols = std::max(320 / NG, 2 * g.r + 16);
        ncols = std::min(ncols, 384 / NG);
        int TW = ncols - 2 * g.r;
        if (TW < 8) fast_ok = false;
        else {
            TW = std::min(TW, ((XB - XA + 7) / 8) * 8);
            ncols = TW + 2 * g.r;
            nt = ((ncols * NG + 31) / 32) * 32;
            P.Lp = Lp; P.Rp = Rp; P.disp = disp; P.cost = cost;
            P.W = W; P.H = H; P.nd = cfg.nd; P.minD = cfg.minD; P.r = g.r; P.cap = cfg.cap;
            P.texThr = cfg.textureThreshold; P.uniq = cfg.uniquenessRatio; P.lofs = g.lofs;
            P.XA = XA; P.XB = XB; P.YA = g.roiY0; P.YB = g.roiY1;
            P.TW = TW; P.ncols = ncols; P.NG = NG;
            P.NGH = cfg.nd / 8;
            P.NS = std::max(1, std::min(nt / P.NGH, (TW + 15) / 16));
            P.SWD = (TW + P.NS - 1) / P.NS;
            P.CW = cfg.nd / 2 + 4;
            P.RL = ncols + cfg.nd - 1;
            int cs = ((P.RL + 15) / 16) * 16 + 16;
            if (((cs / 16) & 1) == 0) cs += 16;
            P.CS = cs;
            P.RbS = cs + 32;
            P.LS = ((ncols + 15) / 16) * 16;
            size_t o = 0;
            P.oL = (int)o; o += 2 * (size_t)P.LS;
            P.oRb = (int)o; o += 2 * (size_t)P.RbS;
            o = (o + 15) & ~(size_t)15;
            P.oRc = (int)o; o += 2 * 16 * (size_t)P.CS;
            P.oC = (int)o; o += (size_t)ncols * P.CW * 4;
            P.oS = (int)o; o += (size_t)TW * P.CW * 4;
            P.oT = (int)o; o += (size_t)ncols * 4;
            smem = o;
            if (smem > 200 * 1024 || nt > 384) fast_ok = false;
            // band height: fill the machine with ~2 blocks per SM, warm-up overhead bounded
            int tilesX = (XB - XA + TW - 1) / TW;
            int rows = g.roiY1 - g.roiY0;
            int want_blocks = 148 * 2;
            int bands = std::max(1, std::min(want_blocks / tilesX, std::max(1, rows / (4 * g.r + 8))));
            P.BH = (rows + bands - 1) / bands;
            grid = dim3(tilesX, (rows + P.BH - 1) / P.BH);
        }
    }
    if (fast_ok) {
        cudaError_t e;
        if (KPT == 64) e = launch_fast<64>(P, grid, nt, smem, st);
        else if (KPT == 32) e = launch_fast<32>(P, grid, nt, smem, st);
        else e = launch_fast<16>(P, grid, nt, smem, st);
        if (e != cudaSuccess) return -1;
        ++launches;
        // border bands that the fast kernel does not cover
        int l = run_generic(Lp, Rp, W, H, cfg, g, outX0 - g.lofs, XA - g.lofs, disp, cost, sc, st);
        if (l < 0) return l;
        launches += l;
        l = run_generic(Lp, Rp, W, H, cfg, g, XB - g.lofs, outX1 - g.lofs, disp, cost, sc, st);
        if (l < 0) return l;
        launches += l;
    } else {
        int l = run_generic(Lp, Rp, W, H, cfg, g, outX0 - g.lofs, outX1 - g.lofs, disp, cost, sc, st);
        if (l < 0) return l;
        launches += l;
    }
    return launches;
}

}  // namespace b200s
