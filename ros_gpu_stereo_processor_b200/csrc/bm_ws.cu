// Warp-specialised SAD block matcher (v4) for sm_100a: same arithmetic as bm_fast_kernel (bm_sad.cu) but the three
// per-row phases run concurrently on different rows, connected by double-buffered shared memory and named barriers:
//
//   stager warps (2)  global rows -> stage[j & 1]         (left bytes pre-broadcast, right row as 4 word-shifted copies)
//   V warps           stage -> 32 column-sum registers -> Cbuf[o & 1]      (VABSDIFF4.U8, u16x2 lanes, vertical sliding)
//   H warps           Cbuf -> Sbuf[o & 1] + Kbuf[o & 1]                    (horizontal sliding sums, per-8 min keys)
//   W warps           Sbuf/Kbuf -> disparity                               (argmin, uniqueness, texture, sub-pixel)
//
// One block per SM (up to ~222 KB of shared memory), 24 warps.  Producer/consumer hand-over uses PTX named barriers:
// producers `bar.arrive full[b]`, consumers `bar.sync full[b]`; consumers `bar.arrive empty[b]`, producers
// `bar.sync empty[b]` before overwriting.  No role ever waits on a block-wide barrier inside the row loop.
#include "kernels.h"
#include "bm_common.cuh"

#include <algorithm>
#include <cstdlib>

namespace b200s {

struct WsParams {
    const uint8_t* Lp;
    const uint8_t* Rp;
    size_t pitch;
    int16_t* disp;
    int16_t* cost;
    int W, H, nd, minD, r, cap, texThr, uniq, lofs;
    int X0base, XA, XB, YA, YB;
    int TW, BH, ncols, NCQ, NK, NGH, NS, SWD;
    int CWb, SWb, KWb, NK4, CSB, RLW;
    int nVw, nHw, nWw;     // warps per role; nSw (1..4) stager warps follow
    int nSw;
    int rowsS, rowsC;
    int oStage[2];         // per buffer: Lb [2][ncols] words, then Rc [2][4][CSB] bytes
    int oTc;               // [8][ncols] words: texture column sums, ring indexed by output row & 7
    int oC[2], oS[2], oK[2];
};

enum { B_FULL_STAGE = 1, B_EMPTY_STAGE = 3, B_FULL_C = 5, B_EMPTY_C = 7, B_FULL_S = 9, B_EMPTY_S = 11 };

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// producer side: publish the shared-memory writes, then arrive
__device__ __forceinline__ void bar_arrive(int id, int n)
{
    asm volatile("fence.acq_rel.cta;" ::: "memory");
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
}
// consumer side: the buffer was only read (and the loaded values already consumed), nothing to publish
__device__ __forceinline__ void bar_release(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// register budget: ptxas rounds the block up to a multiple of 128 threads, so 640 threads -> 96 registers (the ring
// variant of the H role needs them), 768 threads -> 80 registers
template <int RB> struct WsBounds { static constexpr int threads = RB > 0 ? 640 : 768; };

template <int ND, int RB>
__global__ void __launch_bounds__(WsBounds<RB>::threads, 1) bm_ws_kernel(const WsParams P)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int nd = ND > 0 ? ND : P.nd;
    const int CWb = ND > 0 ? ND * 2 : P.CWb;
    const int SWb = ND > 0 ? ND * 2 + 16 : P.SWb;
    const int KWb = ND > 0 ? (((ND / 8 + 3) / 4) * 4 + 4) * 4 : P.KWb;
    const int NK = ND > 0 ? ND / 16 : P.NK;
    const int NGH = ND > 0 ? ND / 8 : P.NGH;
    const int NK4 = ND > 0 ? (ND / 8 + 3) / 4 : P.NK4;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int NVt = P.nVw * 32, NHt = P.nHw * 32, NWt = P.nWw * 32, NSt = P.nSw * 32;
    const int X0 = P.X0base + blockIdx.x * P.TW;
    const int yb0 = P.YA + blockIdx.y * P.BH;
    const int yb1 = min(yb0 + P.BH, P.YB);
    const int r = P.r, b = 2 * r + 1;
    const int nIn = yb1 - yb0 + 2 * r;     // input rows consumed
    const int nOut = yb1 - yb0;            // output rows produced
    const int y_in0 = yb0 - r;             // image row of input index 0

    // pad keys stay "infinite" for the whole kernel
    for (int i = tid; i < (P.rowsS * KWb) / 4; i += blockDim.x) {
        ((uint32_t*)(smem + P.oK[0]))[i] = 0xFFFFFFFFu;
        ((uint32_t*)(smem + P.oK[1]))[i] = 0xFFFFFFFFu;
    }
    __syncthreads();

    if (warp < P.nVw) {
        // =============================== V role ===============================================================
        const int cq = tid / NK, kg = tid - cq * NK;
        const bool vact = cq < P.NCQ;
        uint32_t C[4][2][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int w = 0; w < 4; ++w) C[i][h][w] = 0;
        int voff[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int u = kg + h * NK;
            const int j = (cq + 2 * u) & 3;
            voff[h] = 2 * P.ncols * 4 + j * P.CSB + (4 * cq + 8 * u - 4 * j);   // inside a stage buffer, new row
        }
        const int cstore = (4 * cq) * CWb + 16 * kg;
        for (int j = 0; j < nIn; ++j) {
            const int sb = j & 1;
            bar_sync(B_FULL_STAGE + sb, NVt + NSt);
            if (vact) {
                const uint8_t* st = smem + P.oStage[sb];
                const uint4 ln4 = *(const uint4*)(st + 16 * cq);
                const uint4 lo4 = *(const uint4*)(st + 4 * P.ncols + 16 * cq);
                const uint32_t ln[4] = {ln4.x, ln4.y, ln4.z, ln4.w};
                const uint32_t lo[4] = {lo4.x, lo4.y, lo4.z, lo4.w};
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint4 rn4 = *(const uint4*)(st + voff[h]);
                    const uint4 ro4 = *(const uint4*)(st + voff[h] + 4 * P.CSB);
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
                            C[i][h][2 * w] += t & 0x00ff00ffu;                 // lanes k+0, k+2 (bias 128 per lane kept)
                            C[i][h][2 * w + 1] += __byte_perm(t, 0, 0x4341);   // lanes k+1, k+3
                        }
                    }
                }
            }
            bar_release(B_EMPTY_STAGE + sb, NVt + NSt);   // after the loaded values were consumed
            if (j >= 2 * r) {
                const int o = j - 2 * r, cb = o & 1;
                if (o >= 2) bar_sync(B_EMPTY_C + cb, NVt + NHt);
                if (vact) {
                    uint8_t* dst = smem + P.oC[cb] + cstore;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int h = 0; h < 2; ++h)
                            *(uint4*)(dst + i * CWb + h * (16 * NK)) = make_uint4(C[i][h][0], C[i][h][1], C[i][h][2], C[i][h][3]);
                }
                bar_arrive(B_FULL_C + cb, NVt + NHt);
            }
        }
    } else if (warp < P.nVw + P.nHw) {
        // =============================== H role ===============================================================
        const int ht = tid - NVt;
        const int hs = ht / NGH, gh = ht - hs * NGH;
        const bool hact = hs < P.NS;
        const int xs = hs * P.SWD;
        for (int o = 0; o < nOut; ++o) {
            const int cb = o & 1;
            bar_sync(B_FULL_C + cb, NVt + NHt);
            if (o >= 2) bar_sync(B_EMPTY_S + cb, NHt + NWt);
            if (hact) {
                const uint8_t* pc = smem + P.oC[cb] + xs * CWb + 16 * gh;
                // every column sum carries a bias of 128 per accumulated row and lane; remove b columns' worth
                const uint32_t bias = (uint32_t)(128 * (o + 2 * r + 1) * b) * 0x00010001u;
                uint4 S = make_uint4(0u - bias, 0u - bias, 0u - bias, 0u - bias);
                uint8_t* ps = smem + P.oS[cb] + xs * SWb + 16 * gh;
                uint8_t* pk = smem + P.oK[cb] + xs * KWb + 4 * gh;
#define B200S_H_EMIT(OFFS, OFFK)                                                     \
    {                                                                                \
        *(uint4*)(ps + (OFFS)) = S;                                                  \
        uint32_t m_ = __vimin3_u16x2(S.x, S.y, S.z);                                 \
        m_ = __vminu2(m_, S.w);                                                      \
        m_ = __vminu2(m_, m_ >> 16);                                                 \
        *(uint32_t*)(pk + (OFFK)) = (m_ << 16) | (uint32_t)gh;                       \
    }
                if (RB > 0) {
                    // the leaving column comes from a register ring (window width RB = 2r+1 is compile time):
                    // one LDS.128 per column instead of two; strips are whole multiples of RB
                    uint4 ring[RB > 0 ? RB : 1];
#pragma unroll
                    for (int c = 0; c < RB; ++c) {
                        ring[c] = *(const uint4*)(pc + c * CWb);
                        S.x += ring[c].x; S.y += ring[c].y; S.z += ring[c].z; S.w += ring[c].w;
                    }
                    const uint8_t* pa = pc + RB * CWb;
                    uint4 a = *(const uint4*)pa;                 // entering column of the first pixel, loaded ahead
                    for (int x0 = 0; x0 < P.SWD; x0 += RB) {
#pragma unroll
                        for (int i = 0; i < RB; ++i) {
                            const uint4 an = *(const uint4*)(pa + (i + 1) * CWb);   // next pixel's entering column
                            B200S_H_EMIT(i * SWb, i * KWb)
                            S.x += a.x - ring[i].x; S.y += a.y - ring[i].y; S.z += a.z - ring[i].z; S.w += a.w - ring[i].w;
                            ring[i] = a;
                            a = an;
                        }
                        pa += RB * CWb; ps += RB * SWb; pk += RB * KWb;
                    }
                } else {
                    for (int c = 0; c < b; ++c) {
                        const uint4 v = *(const uint4*)(pc + c * CWb);
                        S.x += v.x; S.y += v.y; S.z += v.z; S.w += v.w;
                    }
                    const uint8_t* pa = pc + b * CWb;
                    // software pipeline: the entering/leaving columns of the next pair are loaded before the stores of
                    // the current pair (the compiler will not hoist shared loads over shared stores by itself)
                    uint4 a0 = *(const uint4*)(pa), o0 = *(const uint4*)(pc);
                    uint4 a1 = *(const uint4*)(pa + CWb), o1 = *(const uint4*)(pc + CWb);
                    for (int x0 = 0; x0 < P.SWD; x0 += 2) {     // SWD is a multiple of 4; Cbuf has spare rows behind
                        pa += 2 * CWb; pc += 2 * CWb;
                        const uint4 na0 = *(const uint4*)(pa), no0 = *(const uint4*)(pc);
                        const uint4 na1 = *(const uint4*)(pa + CWb), no1 = *(const uint4*)(pc + CWb);
                        B200S_H_EMIT(0, 0)
                        S.x += a0.x - o0.x; S.y += a0.y - o0.y; S.z += a0.z - o0.z; S.w += a0.w - o0.w;
                        B200S_H_EMIT(SWb, KWb)
                        S.x += a1.x - o1.x; S.y += a1.y - o1.y; S.z += a1.z - o1.z; S.w += a1.w - o1.w;
                        a0 = na0; o0 = no0; a1 = na1; o1 = no1;
                        ps += 2 * SWb; pk += 2 * KWb;
                    }
                }
#undef B200S_H_EMIT
            }
            bar_release(B_EMPTY_C + cb, NVt + NHt);
            bar_arrive(B_FULL_S + cb, NHt + NWt);
        }
    } else if (warp < P.nVw + P.nHw + P.nWw) {
        // =============================== W role ===============================================================
        const int px = tid - NVt - NHt;
        const bool wact = px < P.TW;
        const int16_t FILTERED = (int16_t)((P.minD - 1) * 16);
        const int X = X0 + px;
        const bool wout = wact && X >= P.XA && X < P.XB;
        for (int o = 0; o < nOut; ++o) {
            const int cb = o & 1;
            bar_sync(B_FULL_S + cb, NHt + NWt);
            if (wact) {
                uint8_t* krow = smem + P.oK[cb] + px * KWb;
                uint8_t* srow = smem + P.oS[cb] + px * SWb;
                uint32_t best = 0xFFFFFFFFu;
                if (ND > 0) {
#pragma unroll
                    for (int i = 0; i < NK4; ++i) {
                        const uint4 k4 = *(const uint4*)(krow + 16 * i);
                        best = min(min(best, k4.x), min(k4.y, min(k4.z, k4.w)));
                    }
                } else {
                    for (int i = 0; i < NK4; ++i) {
                        const uint4 k4 = *(const uint4*)(krow + 16 * i);
                        best = min(min(best, k4.x), min(k4.y, min(k4.z, k4.w)));
                    }
                }
                const int minsad = (int)(best >> 16), gs = (int)(best & 0xffffu);
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
                    const int g0 = max(gs - 1, 0), g2 = min(gs + 1, NGH - 1);
                    ((uint32_t*)krow)[g0] = 0xFFFFFFFFu;
                    ((uint32_t*)krow)[gs] = 0xFFFFFFFFu;
                    ((uint32_t*)krow)[g2] = 0xFFFFFFFFu;
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
                if (wout) {
                    const int* tc = (const int*)(smem + P.oTc) + (o & 7) * P.ncols + px;
                    int tsum = 0;
                    for (int c = 0; c < b; ++c) tsum += tc[c];
                    int16_t out = FILTERED;
                    if (tsum >= P.texThr && !filtered) out = subpixel_disp(minsad, mind, pv, nv, nd, P.minD);
                    const int y = yb0 + o;
                    P.disp[(size_t)y * P.W + X] = out;
                    if (P.cost) P.cost[(size_t)y * P.W + X] = (int16_t)minsad;
                }
            }
            bar_release(B_EMPTY_S + cb, NHt + NWt);
        }
    } else {
        // =============================== stager role ===========================================================
        // stager 0: left rows (pre-broadcast) + running texture column sums; the others: right rows (4 shifted copies).
        // The global loads of row j + 1 are issued into registers before the wait for the stage buffer of row j, so
        // their latency is off the hand-over path (the planner keeps ncols <= 32 * MAXL, RLW <= 32 * MAXR * (nSw - 1)).
        constexpr int MAXL = 5, MAXR = 4;
        const int s = warp - (P.nVw + P.nHw + P.nWw);
        const int lane = tid & 31;
        const int Xl0 = X0 - r;
        const int Xr0 = X0 - r - P.lofs;     // multiple of 4 by construction
        if (s == 0) {
            int an[MAXL], ao[MAXL], trun[MAXL];
#pragma unroll
            for (int m = 0; m < MAXL; ++m) {
                const int c = lane + 32 * m;
                an[m] = c < P.ncols ? (int)__ldg(P.Lp + (size_t)y_in0 * P.pitch + Xl0 + c) : 0;
                ao[m] = 0;
                trun[m] = 0;
            }
            for (int j = 0; j < nIn; ++j) {
                const int sb = j & 1;
                int cn[MAXL], co[MAXL];
#pragma unroll
                for (int m = 0; m < MAXL; ++m) { cn[m] = an[m]; co[m] = ao[m]; }
                if (j + 1 < nIn) {
                    const int yi = y_in0 + j + 1;
                    const bool has_old = j + 1 >= b;
                    const uint8_t* ln = P.Lp + (size_t)yi * P.pitch + Xl0;
                    const uint8_t* lo = P.Lp + (size_t)max(yi - b, 0) * P.pitch + Xl0;
#pragma unroll
                    for (int m = 0; m < MAXL; ++m) {
                        const int c = lane + 32 * m;
                        if (c < P.ncols) {
                            an[m] = (int)__ldg(ln + c);
                            ao[m] = has_old ? (int)__ldg(lo + c) : 0;
                        }
                    }
                }
                if (j >= 2) bar_sync(B_EMPTY_STAGE + sb, NVt + NSt);
                uint32_t* sLb = (uint32_t*)(smem + P.oStage[sb]);
                const bool publish = j >= 2 * r, has_old = j >= b;
                int* tpub = (int*)(smem + P.oTc) + ((j - 2 * r) & 7) * P.ncols;
#pragma unroll
                for (int m = 0; m < MAXL; ++m) {
                    const int c = lane + 32 * m;
                    if (c < P.ncols) {
                        sLb[c] = (uint32_t)cn[m] * 0x01010101u;
                        sLb[P.ncols + c] = (uint32_t)co[m] * 0x01010101u;
                        trun[m] += abs(cn[m] - P.cap) - (has_old ? abs(co[m] - P.cap) : 0);
                        if (publish) tpub[c] = trun[m];   // texture column sums of output row j - 2r (ring of 8 rows)
                    }
                }
                bar_arrive(B_FULL_STAGE + sb, NVt + NSt);
            }
        } else {
            const int nrw = P.nSw - 1, rw = s - 1;
            const int w0 = rw * 32 + lane, wstep = 32 * nrw;
            uint32_t vn[MAXR], vo[MAXR];
#pragma unroll
            for (int m = 0; m < MAXR; ++m) {
                const int wi = w0 + wstep * m;
                vn[m] = wi < P.RLW ? __ldg((const uint32_t*)(P.Rp + (size_t)y_in0 * P.pitch + Xr0) + wi) : 0u;
                vo[m] = 0u;
            }
            for (int j = 0; j < nIn; ++j) {
                const int sb = j & 1;
                uint32_t cn[MAXR], co[MAXR];
#pragma unroll
                for (int m = 0; m < MAXR; ++m) { cn[m] = vn[m]; co[m] = vo[m]; }
                if (j + 1 < nIn) {
                    const int yi = y_in0 + j + 1;
                    const bool has_old = j + 1 >= b;
                    const uint32_t* rn = (const uint32_t*)(P.Rp + (size_t)yi * P.pitch + Xr0);
                    const uint32_t* ro = (const uint32_t*)(P.Rp + (size_t)max(yi - b, 0) * P.pitch + Xr0);
#pragma unroll
                    for (int m = 0; m < MAXR; ++m) {
                        const int wi = w0 + wstep * m;
                        if (wi < P.RLW) {
                            vn[m] = __ldg(rn + wi);
                            vo[m] = has_old ? __ldg(ro + wi) : 0u;
                        }
                    }
                }
                if (j >= 2) bar_sync(B_EMPTY_STAGE + sb, NVt + NSt);
                uint8_t* sRc = smem + P.oStage[sb] + 2 * P.ncols * 4;
#pragma unroll
                for (int m = 0; m < MAXR; ++m) {
                    const int wi = w0 + wstep * m;
                    if (wi < P.RLW) {
                        uint8_t* cp = sRc + 4 * wi;
                        // copy jj holds row[a + 4 jj] at byte a
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
                            if (wi >= jj) {
                                *(uint32_t*)(cp + jj * P.CSB - 4 * jj) = cn[m];
                                *(uint32_t*)(cp + (4 + jj) * P.CSB - 4 * jj) = co[m];
                            }
                    }
                }
                bar_arrive(B_FULL_STAGE + sb, NVt + NSt);
            }
        }
    }
}

template <int ND, int RB>
static cudaError_t launch_ws2(const WsParams& P, dim3 grid, int nt, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(bm_ws_kernel<ND, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    bm_ws_kernel<ND, RB><<<grid, nt, smem, st>>>(P);
    return cudaGetLastError();
}

template <int ND>
static cudaError_t launch_ws(const WsParams& P, int rb, dim3 grid, int nt, size_t smem, cudaStream_t st)
{
    if (rb == 11) return launch_ws2<ND, 11>(P, grid, nt, smem, st);
    return launch_ws2<ND, 0>(P, grid, nt, smem, st);
}

static inline bool ring_width_supported(int b) { return b == 11; }

// returns 1 when launched, 0 when this configuration is not handled (caller falls back), < 0 on CUDA errors
int launch_bm_ws(const uint8_t* Lp, const uint8_t* Rp, size_t pitch, int W, int H, const BMConfig& cfg, int r, int lofs,
                 int XA, int XB, int YA, int YB, int16_t* disp, int16_t* cost, cudaStream_t st)
{
    const int nd = cfg.nd;
    const int NK = nd / 16, NGH = nd / 8;
    static const int max_warps = getenv("B200S_WS_WARPS") ? atoi(getenv("B200S_WS_WARPS")) : 24;
    // H strip width: longer strips amortise the 2r+1 warm-up loads (the kernel is bound by shared-memory wavefronts),
    // shorter ones give more H warps; measured best on B200: 21 columns at nd = 256, ~14 below
    static const int swd_env = getenv("B200S_SWD") ? atoi(getenv("B200S_SWD")) : 0;
    const int swd_min = swd_env > 0 ? swd_env : (nd >= 256 ? 21 : 14);
    static const int ring_mult = getenv("B200S_RING_MULT") ? atoi(getenv("B200S_RING_MULT")) : 2;
    // the register-ring variant of the H role (one LDS.128 per column) measured slower on B200 (fewer, longer strips:
    // 323 us vs 274 us on C4) and is off unless B200S_RING=1
    static const int ring_on = getenv("B200S_RING") ? atoi(getenv("B200S_RING")) : 0;
    const int rb = (ring_on && ring_width_supported(2 * r + 1)) ? 2 * r + 1 : 0;
    const size_t smem_max = 227 * 1024 - 1024;
    const int X0base = XA - ((XA - r - lofs) & 3);
    const int need = ((XB - X0base + 3) / 4) * 4;
    WsParams P, best;
    size_t smem = 0;
    int nt = 0;
    bool ok = false;
    double best_cost = 1e300;
    int best_bands = 1;
    const int rows = YB - YA;
    static const int n_sm = getenv("B200S_WS_BLOCKS") ? atoi(getenv("B200S_WS_BLOCKS")) : 148;
    // Enumerate tile widths; cost model: waves x (band rows + warm-up rows) x window columns (the V role dominates).
    // Small images prefer narrow tiles and short bands so that the grid still covers the SMs.
    for (int NCQ = std::min(128, (need + 2 * r + 3) / 4); 4 * NCQ - 2 * r >= 8; --NCQ) {
        int TW = std::min((4 * NCQ - 2 * r) & ~3, need);
        int nVw = (NCQ * NK + 31) / 32, nWw = (TW + 31) / 32;
        int NS, SWD;
        if (rb) {                               // strips are whole multiples of the window width (register ring)
            SWD = rb * ring_mult;
            NS = (TW + SWD - 1) / SWD;
        } else {
            NS = std::max(1, (TW + swd_min - 1) / swd_min);
            SWD = (((TW + NS - 1) / NS) + 1) & ~1;  // even (the H loop handles two columns per iteration)
            NS = (TW + SWD - 1) / SWD;
        }
        int nHw = (NS * NGH + 31) / 32;
        const int warp_cap = rb ? std::min(max_warps, 20) : max_warps;   // ring variant: 640 threads, 96 registers
        static const int stagers_env = getenv("B200S_STAGERS") ? atoi(getenv("B200S_STAGERS")) : 4;
        int nSw = std::max(1, std::min(stagers_env, warp_cap - (nVw + nHw + nWw)));   // staging sits on the critical path
        static const int min_stagers = getenv("B200S_MIN_STAGERS") ? atoi(getenv("B200S_MIN_STAGERS")) : 2;
        if (nSw < min_stagers && NCQ > 8) continue;   // a single stager warp costs ~15 % (measured); prefer a narrower tile
        if (nVw + nHw + nWw + nSw > warp_cap) continue;
        if (nSw < 2) continue;                               // one warp for the left rows, the others share the right rows
        const int ncols = 4 * NCQ;
        if (ncols > 32 * 5 || (ncols + nd) / 4 + 1 > 32 * 4 * (nSw - 1)) continue;   // register-prefetch limits of the stagers
        const int rowsS = NS * SWD, rowsC = std::max(ncols, rowsS + 2 * r + 4);
        P.CWb = nd * 2; P.SWb = nd * 2 + 16; P.NK4 = (NGH + 3) / 4; P.KWb = (P.NK4 * 4 + 4) * 4;
        P.RLW = (ncols + nd) / 4 + 1;
        int units = (4 * P.RLW + 15) / 16;
        while ((units & 3) != 2) ++units;
        P.CSB = units * 16;
        size_t o = 0;
        for (int s2 = 0; s2 < 2; ++s2) { P.oStage[s2] = (int)o; o += 2 * (size_t)ncols * 4 + 8 * (size_t)P.CSB; o = (o + 15) & ~(size_t)15; }
        P.oTc = (int)o; o += 9 * (size_t)ncols * 4; o = (o + 15) & ~(size_t)15;
        for (int s2 = 0; s2 < 2; ++s2) { P.oK[s2] = (int)o; o += (size_t)rowsS * P.KWb; }
        for (int s2 = 0; s2 < 2; ++s2) { P.oC[s2] = (int)o; o += (size_t)rowsC * P.CWb; }
        for (int s2 = 0; s2 < 2; ++s2) { P.oS[s2] = (int)o; o += (size_t)rowsS * P.SWb; }
        if (o > smem_max) continue;
        P.TW = TW; P.ncols = ncols; P.NCQ = NCQ; P.NK = NK; P.NGH = NGH; P.NS = NS; P.SWD = SWD;
        P.nVw = nVw; P.nHw = nHw; P.nWw = nWw; P.nSw = nSw; P.rowsS = rowsS; P.rowsC = rowsC;
        const int tilesX = (XB - X0base + TW - 1) / TW;
        const int max_bands = std::max(1, rows / (2 * r + 4));
        for (int bands = 1; bands <= max_bands; ++bands) {
            int BH = (rows + bands - 1) / bands;
            if (BH > 480 - 2 * r) continue;                      // the per-row bias of the column sums bounds the band
            int nb = tilesX * ((rows + BH - 1) / BH);
            int waves = (nb + n_sm - 1) / n_sm;
            double cost = (double)waves * (BH + 2 * r + 6) * (ncols + 16);   // +6 rows: pipeline fill/drain
            if (cost < best_cost) {
                best_cost = cost; best = P; best_bands = bands; smem = o;
                nt = 32 * (nVw + nHw + nWw + nSw);
                ok = true;
            }
        }
    }
    if (!ok || nt > 768) return 0;
    P = best;
    P.Lp = Lp; P.Rp = Rp; P.pitch = pitch; P.disp = disp; P.cost = cost;
    P.W = W; P.H = H; P.nd = nd; P.minD = cfg.minD; P.r = r; P.cap = cfg.cap;
    P.texThr = cfg.textureThreshold; P.uniq = cfg.uniquenessRatio; P.lofs = lofs;
    P.X0base = X0base; P.XA = XA; P.XB = XB; P.YA = YA; P.YB = YB;
    const int tilesX = (XB - X0base + P.TW - 1) / P.TW;
    P.BH = (rows + best_bands - 1) / best_bands;
    dim3 grid(tilesX, (rows + P.BH - 1) / P.BH);
    cudaError_t e;
    if (nd == 256) e = launch_ws<256>(P, rb, grid, nt, smem, st);
    else if (nd == 128) e = launch_ws<128>(P, rb, grid, nt, smem, st);
    else if (nd == 64) e = launch_ws<64>(P, rb, grid, nt, smem, st);
    else e = launch_ws<0>(P, rb, grid, nt, smem, st);
    return e == cudaSuccess ? 1 : -1;
}

}  // namespace b200s
