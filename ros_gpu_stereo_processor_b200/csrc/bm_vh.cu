// SAD block matcher v7 for sm_100a: both window sums of cv::StereoBM live in registers.
//
// The v4 kernel (bm_ws.cu) keeps the vertical column sums in registers and pushes them through shared memory for the
// horizontal sliding sum; more than half of its shared-memory wavefronts -- its binding resource -- are that traffic.
// Here one thread owns 16 adjacent output columns x 4 disparities and keeps the finished window sums S of those 64
// (pixel, disparity) pairs in 32 registers for the whole band:
//
//   per input row   E[e]  = |L-R|(entering row) - |L-R|(leaving row)            packed bytes, 4 disparities per word,
//                           for the 16 + 2r window columns of the thread (VABSDIFF4.U8; halo columns are recomputed
//                           instead of exchanged)
//                   R_0   = sum of the first 2r+1 E words (byte-pair sums, then widened to u16x2 lanes)
//                   R_i+1 = R_i +/- widen(128 +/- (E[i+2r+1] - E[i]))                 horizontal sliding, in registers
//                   S[i] += R_i                                                      vertical sliding, in registers
// The sign of the delta word alternates with i, so the +128 of its byte lanes cancels every second step instead of
// costing a subtraction per step: R (and with it S) of the odd columns carries a known bias of 128 per accumulated
// row, which the W role removes from the handful of values it finally uses (a common offset does not move the argmin).
//
// Everything is exact integer arithmetic: the staged bytes are clamped to 2*cap <= 62, so no byte lane of a biased
// pair sum or delta word can wrap, and the u16x2 registers are only ever combined linearly (a register is the integer
// lo + 65536 * hi, valid whenever the final lanes are inside [0, 65535]: the planner keeps
// 2*cap*block^2 + 128 * (band rows + 2r + 1) <= 65535).
//
// Roles (usually one block per SM, warp-specialised, mbarrier full/empty pairs, double-buffered shared memory):
//   stager warps   global rows -> stage[j & 1]  (left bytes pre-broadcast, right row as 4 word-shifted copies; the loads
//                  of row j + 1 are in flight while the buffer of row j is awaited)
//   VH warps       stage -> S registers -> Sbuf[o & 1] (all window sums of the row) + Kbuf[o & 1] (min per 4 disparities)
//   W warps        Sbuf/Kbuf -> disparity (argmin, uniqueness, texture, sub-pixel), one thread per pixel
#include "kernels.h"
#include "bm_common.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace b200s {

struct VhParams {
    const uint8_t* Lp;
    const uint8_t* Rp;
    size_t pitch;
    int16_t* disp;
    int16_t* cost;
    int W, H, nd, minD, r, cap, texThr, uniq, lofs;
    int X0base, XA, XB, YA, YB;
    int TW, BH, ncols, ncolsP;       // tile width (16 * NCB), band height, window columns (TW + 2r), padded stride (words)
    int NCB, G4;                     // column blocks, nd / 4
    int SWb, KWb, NK16, CSB, RLW;    // Sbuf / Kbuf row strides, LDS.128 per key row
    int nVw, nWw, nSw;
    int oStage[2];                   // per buffer: Lb [2][ncolsP] words, then Rc [2][4][CSB] bytes
    int oTc;                         // [8][ncols] words: texture column sums (ring over output rows)
    int oS[2], oK[2];
    int oMbar;                       // 8 mbarriers
    int oX;                          // halo exchange of the VH threads (HX): uint4 [2 buffers][2 sides][chunks][VH threads]
    size_t pre_stride, disp_stride;  // bytes between the frames of a batch (blockIdx.z = frame)
};

namespace vh {

// Hand-over between the roles: shared-memory mbarriers, one full/empty pair per buffer.  Unlike a named barrier
// (bar.sync makes every consumer wait for every other consumer), a consumer only waits for the producers, so the VH
// warps drift apart by up to one row and their ALU-heavy and IMAD-heavy phases overlap on the two integer pipes.
enum { MB_FULL_STAGE = 0, MB_EMPTY_STAGE = 2, MB_FULL_S = 4, MB_EMPTY_S = 6, MB_COUNT = 8 };

__device__ __forceinline__ void mbar_init(uint32_t a, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t a)
{
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MB_WAIT:\n"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MB_DONE;\n"
        "bra MB_WAIT;\n"
        "MB_DONE:\n"
        "}" ::"r"(a), "r"(parity) : "memory");
}

constexpr int NC = 16;               // output columns per VH thread

}  // namespace vh

#ifndef B200S_VH_MAXT
#define B200S_VH_MAXT 768     // register budget of the matcher = 65536 / this (rounded down to a multiple of 8)
#endif
// E words e in [LO, HI) of one VH thread for the current row pair: E = |L-R|(entering row) - |L-R|(leaving row) as one
// 32-bit integer, byte lanes in [-2cap, 2cap] with borrows between them (every later use adds a constant that makes all
// four lanes positive).  st: stage buffer; wn / wo: the thread's right-row words (entering / leaving row).
template <int LO, int HI, int NWORDS, int NEALL>
__device__ __forceinline__ void vh_e_words(const uint8_t* st, int loffs, int ncolsP, const uint32_t (&wn)[NWORDS],
                                           const uint32_t (&wo)[NWORDS], uint32_t (&E)[NEALL])
{
#pragma unroll
    for (int q = LO / 4; q < (HI + 3) / 4; ++q) {
        const uint4 ln4 = *(const uint4*)(st + loffs + 16 * q);
        const uint4 lo4 = *(const uint4*)(st + 4 * ncolsP + loffs + 16 * q);
        const uint32_t ln[4] = {ln4.x, ln4.y, ln4.z, ln4.w};
        const uint32_t lo[4] = {lo4.x, lo4.y, lo4.z, lo4.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int e = 4 * q + c;
            if (e >= LO && e < HI) {
                const uint32_t rn = c ? __funnelshift_r(wn[q], wn[q + 1], 8 * c) : wn[q];
                const uint32_t ro = c ? __funnelshift_r(wo[q], wo[q + 1], 8 * c) : wo[q];
                E[e] = __vabsdiffu4(ln[c], rn) - __vabsdiffu4(lo[c], ro);
            }
        }
    }
}

// HX (halo exchange) -- an experiment that is NOT in the default build (-DB200S_VH_HX_BUILD instantiates it,
// B200S_VH_HX=1 selects it): instead of recomputing the r window columns on either side of its 16 own columns (20 of 36 E
// words at block 21), a VH thread computes its own 16 E words, publishes the first and the last r of them in shared
// memory and reads its halos from the neighbouring column blocks after one named barrier of the VH warps.  Bit-identical
// (all parity tests pass with it) and 1.6 - 1.7 x SLOWER on C1 / C2 / C3 (profiles/r02_experiments.md section 6): the
// barrier puts the VH warps in lock-step every row, which is exactly what the mbarrier hand-over had removed -- the
// overlap of the VABSDIFF-heavy and the IMAD-heavy phases of different warps on the two integer pipes is worth more than
// the 12 - 20 % of instructions the exchange saves.
//
// WIDE: preFilterCap 32..63.  The staged bytes reach 126, so a byte lane can hold ONE biased E word (128 + e in [2, 254])
// but neither a pair sum nor a difference of two: every E word is widened on its own (the +128 of the entering and of
// the leaving word cancel, so no bias accumulates and the W role has nothing to subtract).  About a quarter more VH
// instructions than the narrow form, still well ahead of the bm_ws fallback these parameters used to take.
template <int R, int ND, bool WIDE, bool HX>
__global__ void __launch_bounds__(B200S_VH_MAXT, 1) bm_vh_kernel(const VhParams P)
{
    using namespace vh;
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int B = 2 * R + 1;
    constexpr int NE = NC + 2 * R;                 // window columns per VH thread
    constexpr int NLQ = (NE + 3) / 4;              // LDS.128 per left row
    constexpr int NWD = (NE + 2) / 4 + 1;          // right-row words a thread needs
    constexpr int NRQ = (NWD + 3) / 4;             // LDS.128 per right row
    const int nd = ND > 0 ? ND : P.nd;
    const int SWb = ND > 0 ? ND * 2 + 16 : P.SWb;
    const int KWb = ND > 0 ? (((ND / 4 + 3) / 4) * 4 + 4) * 4 : P.KWb;
    const int NK16 = ND > 0 ? (ND / 4 + 3) / 4 : P.NK16;
    const int G4 = ND > 0 ? ND / 4 : P.G4;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int NVt = P.nVw * 32, NWt = P.nWw * 32, NSt = P.nSw * 32;
    // frame of the batch this block works on
    const uint8_t* __restrict__ const gLp = P.Lp + blockIdx.z * P.pre_stride;
    const uint8_t* __restrict__ const gRp = P.Rp + blockIdx.z * P.pre_stride;
    int16_t* __restrict__ const gdisp = (int16_t*)((uint8_t*)P.disp + blockIdx.z * P.disp_stride);
    int16_t* __restrict__ const gcost = P.cost ? (int16_t*)((uint8_t*)P.cost + blockIdx.z * P.disp_stride) : nullptr;
    const int X0 = P.X0base + blockIdx.x * P.TW;
    const int yb0 = P.YA + blockIdx.y * P.BH;
    const int yb1 = min(yb0 + P.BH, P.YB);
    const int r = R, b = B;
    const int nIn = yb1 - yb0 + 2 * r;     // input rows consumed
    const int nOut = yb1 - yb0;            // output rows produced
    const int y_in0 = yb0 - r;             // image row of input index 0

    // pad keys stay "infinite" for the whole kernel
    for (int i = tid; i < (P.TW * KWb) / 4; i += blockDim.x) {
        ((uint32_t*)(smem + P.oK[0]))[i] = 0xFFFFFFFFu;
        ((uint32_t*)(smem + P.oK[1]))[i] = 0xFFFFFFFFu;
    }
    const uint32_t mb = (uint32_t)__cvta_generic_to_shared(smem + P.oMbar);
    if (tid == 0) {
        for (int s2 = 0; s2 < 2; ++s2) {
            mbar_init(mb + 8 * (MB_FULL_STAGE + s2), NSt);
            mbar_init(mb + 8 * (MB_EMPTY_STAGE + s2), NVt);
            mbar_init(mb + 8 * (MB_FULL_S + s2), NVt);
            mbar_init(mb + 8 * (MB_EMPTY_S + s2), NWt);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp < P.nVw) {
        // =============================== VH role ==============================================================
        const int cb = tid / G4;                            // the planner makes NCB * G4 a whole number of warps
        const int g4 = tid - cb * G4;
        const int jj = g4 & 3;
        const int loffs = 64 * cb;                                                   // left words of this block
        const int roffs = 2 * P.ncolsP * 4 + jj * P.CSB + (16 * cb + 4 * g4 - 4 * jj);   // right window, entering row
        const int soffs = (NC * cb) * SWb + 8 * g4;
        const int koffs = (NC * cb) * KWb + 4 * g4;
        uint32_t Se[NC], So[NC];
#pragma unroll
        for (int i = 0; i < NC; ++i) Se[i] = So[i] = 0;

        for (int j = 0; j < nIn; ++j) {
            const int sb = j & 1;
            mbar_wait(mb + 8 * (MB_FULL_STAGE + sb), (j >> 1) & 1);
            uint32_t E[NE];
            {
                const uint8_t* st = smem + P.oStage[sb];
                uint32_t wn[4 * NRQ], wo[4 * NRQ];
#pragma unroll
                for (int q = 0; q < NRQ; ++q) {
                    const uint4 a = *(const uint4*)(st + roffs + 16 * q);
                    const uint4 o4 = *(const uint4*)(st + roffs + 4 * P.CSB + 16 * q);
                    wn[4 * q] = a.x; wn[4 * q + 1] = a.y; wn[4 * q + 2] = a.z; wn[4 * q + 3] = a.w;
                    wo[4 * q] = o4.x; wo[4 * q + 1] = o4.y; wo[4 * q + 2] = o4.z; wo[4 * q + 3] = o4.w;
                }
                if (!HX) {
                    vh_e_words<0, NE>(st, loffs, P.ncolsP, wn, wo, E);
                } else {
                    vh_e_words<R, R + NC>(st, loffs, P.ncolsP, wn, wo, E);                 // the thread's own 16 columns
                    if (cb == 0) vh_e_words<0, R>(st, loffs, P.ncolsP, wn, wo, E);         // outer halos of the tile
                    if (cb == P.NCB - 1) vh_e_words<R + NC, NE>(st, loffs, P.ncolsP, wn, wo, E);
                }
            }
            mbar_arrive(mb + 8 * (MB_EMPTY_STAGE + sb));   // the staged rows are in registers now
            if (HX) {
                constexpr int CH = (R + 3) / 4;                    // 16-byte chunks per side
                uint4* xb = (uint4*)(smem + P.oX) + (size_t)(j & 1) * 2 * CH * NVt;
#pragma unroll
                for (int q = 0; q < CH; ++q) {
                    uint32_t a[4], b2[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int i = 4 * q + c;
                        a[c] = i < R ? E[R + i] : 0u;              // first r own words: the right halo of block cb - 1
                        b2[c] = i < R ? E[NC + i] : 0u;            // last r own words: the left halo of block cb + 1
                    }
                    xb[(0 * CH + q) * NVt + tid] = make_uint4(a[0], a[1], a[2], a[3]);
                    xb[(1 * CH + q) * NVt + tid] = make_uint4(b2[0], b2[1], b2[2], b2[3]);
                }
                asm volatile("bar.sync 1, %0;" ::"r"(NVt) : "memory");
                if (cb > 0) {
#pragma unroll
                    for (int q = 0; q < CH; ++q) {
                        const uint4 v = xb[(1 * CH + q) * NVt + tid - G4];
                        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (4 * q + c < R) E[4 * q + c] = w4[c];
                    }
                }
                if (cb < P.NCB - 1) {
#pragma unroll
                    for (int q = 0; q < CH; ++q) {
                        const uint4 v = xb[(0 * CH + q) * NVt + tid + G4];
                        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (4 * q + c < R) E[R + NC + 4 * q + c] = w4[c];
                    }
                }
            }

            // horizontal window sum of the first pixel: byte-pair sums (<= 252), then widened
            uint32_t Re, Ro;
            if (WIDE) {
                uint32_t se = 0, so = 0;
#pragma unroll
                for (int m = 0; m < B; ++m) {
                    const uint32_t p = E[m] + 0x80808080u;                        // bytes 128 + e in [2, 254]
                    se += p & 0x00ff00ffu;
                    so += __byte_perm(p, 0, 0x4341);
                }
                Re = se - (uint32_t)(128 * B) * 0x00010001u;
                Ro = so - (uint32_t)(128 * B) * 0x00010001u;
            } else {
                uint32_t se = 0, so = 0;
#pragma unroll
                for (int m = 0; m < R; ++m) {
                    const uint32_t p = E[2 * m] + E[2 * m + 1] + 0x80808080u;     // bytes 128 + pair in [4, 252]
                    se += p & 0x00ff00ffu;
                    so += __byte_perm(p, 0, 0x4341);
                }
                const uint32_t pl = E[2 * R] + 0x80808080u;
                se += pl & 0x00ff00ffu;
                so += __byte_perm(pl, 0, 0x4341);
                Re = se - (uint32_t)(128 * (R + 1)) * 0x00010001u;
                Ro = so - (uint32_t)(128 * (R + 1)) * 0x00010001u;
            }
            const bool emit = j >= 2 * r;
            const int cbuf = (j - 2 * r) & 1;
            if (emit) mbar_wait(mb + 8 * (MB_EMPTY_S + cbuf), (((j - 2 * r) >> 1) & 1) ^ 1);
            uint8_t* ps = smem + P.oS[cbuf] + soffs;
            uint8_t* pk = smem + P.oK[cbuf] + koffs;
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                Se[i] += Re;
                So[i] += Ro;
                if (emit) {
                    uint32_t m = __vminu2(Se[i], So[i]);
                    m = __vminu2(m, m << 16);                       // high half: min of the 4 disparities, low half: 0
                    *(uint2*)(ps + i * SWb) = make_uint2(Se[i], So[i]);
                    *(uint32_t*)(pk + i * KWb) = m + (uint32_t)g4;
                }
                if (i + 1 < NC && WIDE) {
                    const uint32_t Dn = E[i + B] + 0x80808080u, Dl = E[i] + 0x80808080u;     // entering / leaving column
                    const uint32_t Dno = __byte_perm(Dn, 0, 0x4341), Dlo = __byte_perm(Dl, 0, 0x4341);
                    Re += (Dn - (Dno << 8)) - (Dl - (Dlo << 8));
                    Ro += Dno - Dlo;
                } else if (i + 1 < NC) {
                    if (!(i & 1)) {
                        const uint32_t D = E[i + B] + 0x80808080u - E[i];      // bytes 128 + d in [4, 252]
                        const uint32_t Do = __byte_perm(D, 0, 0x4341);
                        Re += D - (Do << 8);                                    // even byte lanes = D & 0x00ff00ff
                        Ro += Do;
                    } else {
                        const uint32_t D = E[i] + 0x80808080u - E[i + B];      // bytes 128 - d
                        const uint32_t Do = __byte_perm(D, 0, 0x4341);
                        Re -= D - (Do << 8);
                        Ro -= Do;
                    }
                }
            }
            if (emit) mbar_arrive(mb + 8 * (MB_FULL_S + cbuf));
        }
    } else if (warp < P.nVw + P.nWw) {
        // =============================== W role ===============================================================
        const int px = tid - NVt;
        const bool wact = px < P.TW;
        const int16_t FILTERED = (int16_t)((P.minD - 1) * 16);
        const int X = X0 + px;
        const bool wout = wact && X >= P.XA && X < P.XB;
        for (int o = 0; o < nOut; ++o) {
            const int cb = o & 1;
            mbar_wait(mb + 8 * (MB_FULL_S + cb), (o >> 1) & 1);
            if (wact) {
                uint8_t* krow = smem + P.oK[cb] + px * KWb;
                uint8_t* srow = smem + P.oS[cb] + px * SWb;
                // odd columns of a VH thread carry 128 per accumulated input row on every window sum
                const int bias = (!WIDE && (px & 1)) ? 128 * (o + 2 * r + 1) : 0;
                uint32_t best = 0xFFFFFFFFu;
                if (ND > 0) {
#pragma unroll
                    for (int i = 0; i < NK16; ++i) {
                        const uint4 k4 = *(const uint4*)(krow + 16 * i);
                        best = min(min(best, k4.x), min(k4.y, min(k4.z, k4.w)));
                    }
                } else {
                    for (int i = 0; i < NK16; ++i) {
                        const uint4 k4 = *(const uint4*)(krow + 16 * i);
                        best = min(min(best, k4.x), min(k4.y, min(k4.z, k4.w)));
                    }
                }
                const int minsad_b = (int)(best >> 16), gs = (int)(best & 0xffffu);
                // the winner's group and its two neighbours (4 disparities each, u16 order k, k+2, k+1, k+3) give the exact
                // index, the sub-pixel neighbours and -- with mind-1, mind, mind+1 masked in registers -- their share of the
                // uniqueness test; Sbuf is never written by this role
                const int g0 = max(gs - 1, 0), g2 = min(gs + 1, G4 - 1);
                uint2 e0 = *(const uint2*)(srow + 8 * g0);
                uint2 e1 = *(const uint2*)(srow + 8 * gs);
                uint2 e2 = *(const uint2*)(srow + 8 * g2);
                const int v1[4] = {(int)(e1.x & 0xffffu), (int)(e1.y & 0xffffu), (int)(e1.x >> 16), (int)(e1.y >> 16)};
                int loc = 3;
                if (v1[2] == minsad_b) loc = 2;
                if (v1[1] == minsad_b) loc = 1;
                if (v1[0] == minsad_b) loc = 0;
                const int mind = 4 * gs + loc;
                const int minsad = minsad_b - bias;
                // S[mind + 1] (S[nd - 2] at the upper end) and S[mind - 1] (S[1] at the lower end)
                int pv = loc == 0 ? v1[1] : (loc == 1 ? v1[2] : v1[3]);
                if (loc == 3) pv = gs < G4 - 1 ? (int)(e2.x & 0xffffu) : v1[2];
                int nv = loc == 3 ? v1[2] : (loc == 2 ? v1[1] : v1[0]);
                if (loc == 0) nv = gs > 0 ? (int)(e0.y >> 16) : v1[1];
                pv -= bias;
                nv -= bias;
                bool filtered = false;
                if (P.uniq > 0) {
                    ((uint32_t*)krow)[g0] = 0xFFFFFFFFu;
                    ((uint32_t*)krow)[gs] = 0xFFFFFFFFu;
                    ((uint32_t*)krow)[g2] = 0xFFFFFFFFu;
                    uint32_t m2k = 0xFFFFFFFFu;
                    if (ND > 0) {
#pragma unroll
                        for (int i = 0; i < NK16; ++i) {
                            const uint4 k4 = *(const uint4*)(krow + 16 * i);
                            m2k = min(min(m2k, k4.x), min(k4.y, min(k4.z, k4.w)));
                        }
                    } else {
                        for (int i = 0; i < NK16; ++i) {
                            const uint4 k4 = *(const uint4*)(krow + 16 * i);
                            m2k = min(min(m2k, k4.x), min(k4.y, min(k4.z, k4.w)));
                        }
                    }
                    e1.x |= (loc <= 1 ? 0x0000FFFFu : 0u) | (loc >= 1 ? 0xFFFF0000u : 0u);    // k, k+2
                    e1.y |= (loc <= 2 ? 0x0000FFFFu : 0u) | (loc >= 2 ? 0xFFFF0000u : 0u);    // k+1, k+3
                    if (gs == 0) e0 = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);                    // clamped: the same group again
                    else if (loc == 0) e0.y |= 0xFFFF0000u;
                    if (gs == G4 - 1) e2 = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
                    else if (loc == 3) e2.x |= 0x0000FFFFu;
                    uint32_t acc = __vimin3_u16x2(e0.x, e0.y, e1.x);
                    acc = __vimin3_u16x2(acc, e1.y, e2.x);
                    acc = __vminu2(acc, e2.y);
                    const int m2 = (int)min(min(acc & 0xffffu, acc >> 16), m2k >> 16) - bias;
                    const int thr = minsad + (minsad * P.uniq / 100);
                    filtered = m2 <= thr;
                }
                if (wout) {
                    const int* tc = (const int*)(smem + P.oTc) + (o & 7) * P.ncols + px;
                    int tsum = 0;
#pragma unroll
                    for (int c = 0; c < B; ++c) tsum += tc[c];
                    int16_t out = FILTERED;
                    if (tsum >= P.texThr && !filtered) {
                        // cv::StereoBM's sub-pixel fit (bm_common.cuh subpixel_disp) with the truncating division done in fp32:
                        // |p - n| <= d, so the quotient is at most 256 and one correction step makes it exact
                        const int dd = pv + nv - 2 * minsad + abs(pv - nv);
                        const int num = (pv - nv) * 256, an = abs(num);
                        int q = 0;
                        if (dd != 0) {
                            q = (int)__fdividef((float)an, (float)dd);
                            const int rem = an - q * dd;
                            q += rem < 0 ? -1 : (rem >= dd ? 1 : 0);
                        }
                        out = (int16_t)(((nd - mind - 1 + P.minD) * 256 + (num < 0 ? -q : q) + 15) >> 4);
                    }
                    const int y = yb0 + o;
                    gdisp[(size_t)y * P.W + X] = out;
                    if (gcost) gcost[(size_t)y * P.W + X] = (int16_t)minsad;
                }
            }
            mbar_arrive(mb + 8 * (MB_EMPTY_S + cb));
        }
    } else {
        // =============================== stager role ===========================================================
        // stager 0: left rows (pre-broadcast, clamped) + running texture column sums; the others: right rows.
        // The global loads of row j + 1 are issued before the wait for the stage buffer of row j, so their latency
        // never sits on the hand-over path (the planner keeps ncolsP <= 32 * MAXL and RLW <= 32 * MAXR * (nSw - 1)).
        constexpr int MAXL = 6, MAXR = 4;
        const int s = warp - (P.nVw + P.nWw);
        const int lane = tid & 31;
        const int Xl0 = X0 - r;
        const int Xr0 = X0 - r - P.lofs;     // multiple of 4 by construction
        if (s == 0) {
            int an[MAXL], ao[MAXL], trun[MAXL];      // trun: running texture column sums of the lane's columns
#pragma unroll
            for (int m = 0; m < MAXL; ++m) {
                const int c = lane + 32 * m;
                an[m] = c < P.ncolsP ? (int)__ldg(gLp + (size_t)y_in0 * P.pitch + Xl0 + c) : 0;
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
                    const uint8_t* ln = gLp + (size_t)yi * P.pitch + Xl0;
                    const uint8_t* lo = gLp + (size_t)max(yi - b, 0) * P.pitch + Xl0;
#pragma unroll
                    for (int m = 0; m < MAXL; ++m) {
                        const int c = lane + 32 * m;
                        if (c < P.ncolsP) {
                            an[m] = (int)__ldg(ln + c);
                            ao[m] = has_old ? (int)__ldg(lo + c) : 0;
                        }
                    }
                }
                mbar_wait(mb + 8 * (MB_EMPTY_STAGE + sb), ((j >> 1) & 1) ^ 1);
                uint32_t* sLb = (uint32_t*)(smem + P.oStage[sb]);
                const bool publish = j >= 2 * r;
                const bool has_old = j >= b;
                int* tpub = (int*)(smem + P.oTc) + ((j - 2 * r) & 7) * P.ncols;
#pragma unroll
                for (int m = 0; m < MAXL; ++m) {
                    const int c = lane + 32 * m;
                    if (c < P.ncolsP) {
                        sLb[c] = (uint32_t)min(cn[m], 2 * P.cap) * 0x01010101u;
                        sLb[P.ncolsP + c] = (uint32_t)min(co[m], 2 * P.cap) * 0x01010101u;
                        if (c < P.ncols) {
                            trun[m] += abs(cn[m] - P.cap) - (has_old ? abs(co[m] - P.cap) : 0);
                            if (publish) tpub[c] = trun[m];   // texture column sums of output row j - 2r (ring of 8 rows)
                        }
                    }
                }
                mbar_arrive(mb + 8 * (MB_FULL_STAGE + sb));
            }
        } else {
            const uint32_t clampw = (uint32_t)(2 * P.cap) * 0x01010101u;
            const int nrw = P.nSw - 1, rw = s - 1;
            const int w0 = rw * 32 + lane, wstep = 32 * nrw;
            uint32_t vn[MAXR], vo[MAXR];
#pragma unroll
            for (int m = 0; m < MAXR; ++m) {
                const int wi = w0 + wstep * m;
                vn[m] = wi < P.RLW ? __ldg((const uint32_t*)(gRp + (size_t)y_in0 * P.pitch + Xr0) + wi) : 0u;
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
                    const uint32_t* rn = (const uint32_t*)(gRp + (size_t)yi * P.pitch + Xr0);
                    const uint32_t* ro = (const uint32_t*)(gRp + (size_t)max(yi - b, 0) * P.pitch + Xr0);
#pragma unroll
                    for (int m = 0; m < MAXR; ++m) {
                        const int wi = w0 + wstep * m;
                        if (wi < P.RLW) {
                            vn[m] = __ldg(rn + wi);
                            vo[m] = has_old ? __ldg(ro + wi) : 0u;
                        }
                    }
                }
                mbar_wait(mb + 8 * (MB_EMPTY_STAGE + sb), ((j >> 1) & 1) ^ 1);
                uint8_t* sRc = smem + P.oStage[sb] + 2 * P.ncolsP * 4;
#pragma unroll
                for (int m = 0; m < MAXR; ++m) {
                    const int wi = w0 + wstep * m;
                    if (wi < P.RLW) {
                        const uint32_t a = __vminu4(cn[m], clampw), o2 = __vminu4(co[m], clampw);
                        uint8_t* cp = sRc + 4 * wi;
                        // copy jj holds row[a + 4 jj] at byte a
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
                            if (wi >= jj) {
                                *(uint32_t*)(cp + jj * P.CSB - 4 * jj) = a;
                                *(uint32_t*)(cp + (4 + jj) * P.CSB - 4 * jj) = o2;
                            }
                    }
                }
                mbar_arrive(mb + 8 * (MB_FULL_STAGE + sb));
            }
        }
    }
}

template <int R, int ND, bool WIDE, bool HX>
static cudaError_t launch_vh2(const VhParams& P, dim3 grid, int nt, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(bm_vh_kernel<R, ND, WIDE, HX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // small tiles rely on several blocks per SM: ask for the whole shared-memory carve-out (no L1 use in this kernel)
    e = cudaFuncSetAttribute(bm_vh_kernel<R, ND, WIDE, HX>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    bm_vh_kernel<R, ND, WIDE, HX><<<grid, nt, smem, st>>>(P);
    return cudaGetLastError();
}

// registers per thread of an instantiation (decides how many blocks share an SM), asked from the runtime once
template <int R, int ND, bool WIDE, bool HX>
static int vh_regs2()
{
    static int regs = 0;
    if (!regs) {
        cudaFuncAttributes a;
        regs = cudaFuncGetAttributes(&a, bm_vh_kernel<R, ND, WIDE, HX>) == cudaSuccess && a.numRegs > 0 ? a.numRegs : 80;
    }
    return regs;
}
// the halo exchange experiment: instantiated only with -DB200S_VH_HX_BUILD, where its buffer can fit (64 / 128 disparities)
#ifdef B200S_VH_HX_BUILD
constexpr bool VH_HX_BUILT = true;
#else
constexpr bool VH_HX_BUILT = false;
#endif
static bool vh_hx_available(int nd, bool wide) { return VH_HX_BUILT && !wide && (nd == 64 || nd == 128); }
template <int R>
static int vh_regs1(int nd, bool wide)
{
    if (wide) return vh_regs2<R, 0, true, false>();       // the wide form is instantiated for the generic disparity loop only
    if (nd == 256) return vh_regs2<R, 256, false, false>();
    if (nd == 128) return vh_regs2<R, 128, false, VH_HX_BUILT>();
    if (nd == 64) return vh_regs2<R, 64, false, VH_HX_BUILT>();
    return vh_regs2<R, 0, false, false>();
}
static int vh_regs(int r, int nd, bool wide)
{
    switch (r) {
    case 2: return vh_regs1<2>(nd, wide);
    case 3: return vh_regs1<3>(nd, wide);
    case 4: return vh_regs1<4>(nd, wide);
    case 5: return vh_regs1<5>(nd, wide);
    case 6: return vh_regs1<6>(nd, wide);
    case 7: return vh_regs1<7>(nd, wide);
    case 8: return vh_regs1<8>(nd, wide);
    case 9: return vh_regs1<9>(nd, wide);
    case 10: return vh_regs1<10>(nd, wide);
    default: return 80;
    }
}

struct DeviceShape {       // what the planner needs to know about the GPU it launches on
    int sms = 148, regs_per_sm = 65536, smem_per_sm = 228 * 1024, smem_per_block = 227 * 1024, threads_per_sm = 2048;
};
static const DeviceShape& device_shape()
{
    static DeviceShape shapes[64];
    static bool known[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!known[dev]) {
        DeviceShape d;
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) d.sms = v;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxRegistersPerMultiprocessor, dev) == cudaSuccess && v > 0) d.regs_per_sm = v;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev) == cudaSuccess && v > 0) d.smem_per_sm = v;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess && v > 0) d.smem_per_block = v;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxThreadsPerMultiProcessor, dev) == cudaSuccess && v > 0) d.threads_per_sm = v;
        shapes[dev] = d;
        known[dev] = true;
    }
    return shapes[dev];
}

template <int R>
static cudaError_t launch_vh(const VhParams& P, dim3 grid, int nt, size_t smem, cudaStream_t st)
{
    if (P.cap > 31) return launch_vh2<R, 0, true, false>(P, grid, nt, smem, st);
    if (P.nd == 256) return launch_vh2<R, 256, false, false>(P, grid, nt, smem, st);
    if (P.nd == 128) return P.oX >= 0 ? launch_vh2<R, 128, false, VH_HX_BUILT>(P, grid, nt, smem, st) : launch_vh2<R, 128, false, false>(P, grid, nt, smem, st);
    if (P.nd == 64) return P.oX >= 0 ? launch_vh2<R, 64, false, VH_HX_BUILT>(P, grid, nt, smem, st) : launch_vh2<R, 64, false, false>(P, grid, nt, smem, st);
    return launch_vh2<R, 0, false, false>(P, grid, nt, smem, st);
}

// returns 1 when launched, 0 when this configuration is not handled (caller falls back), < 0 on CUDA errors
int launch_bm_vh(const uint8_t* Lp, const uint8_t* Rp, size_t pitch, int W, int H, const BMConfig& cfg, int r, int lofs,
                 int XA, int XB, int YA, int YB, int16_t* disp, int16_t* cost, cudaStream_t st, int nf, size_t pre_stride,
                 size_t disp_stride)
{
    using vh::NC;
    const int nd = cfg.nd;
    if (cfg.cap > 63 || r < 2 || r > 10 || (nd & 15)) return 0;   // R is a template parameter; cap > 31 takes the WIDE form
    const bool wide = cfg.cap > 31;
    const int G4 = nd / 4;
    static const int max_warps = getenv("B200S_VH_WARPS") ? atoi(getenv("B200S_VH_WARPS")) : 24;
    static const int stagers_env = getenv("B200S_STAGERS") ? atoi(getenv("B200S_STAGERS")) : 2;
    const DeviceShape& gpu = device_shape();
    static const int n_sm_env = getenv("B200S_WS_BLOCKS") ? atoi(getenv("B200S_WS_BLOCKS")) : 0;
    const int n_sm = n_sm_env > 0 ? n_sm_env : gpu.sms;
    const int regs = vh_regs(r, nd, wide);
    static const int ncb_env = getenv("B200S_VH_NCB") ? atoi(getenv("B200S_VH_NCB")) : 0;
    static const int bands_env = getenv("B200S_VH_BANDS") ? atoi(getenv("B200S_VH_BANDS")) : 0;
    static const int verbose = getenv("B200S_VH_VERBOSE") ? atoi(getenv("B200S_VH_VERBOSE")) : 0;
    static const int use_hx = getenv("B200S_VH_HX") ? atoi(getenv("B200S_VH_HX")) : 0;
    const size_t smem_max = (size_t)gpu.smem_per_block - 1024;
    const int X0base = XA - ((XA - r - lofs) & 3);
    const int need = XB - X0base;
    const int rows = YB - YA;
    VhParams P, best;
    size_t smem = 0;
    int nt = 0, best_bands = 1;
    bool ok = false;
    double best_cost = 1e300;
    for (int NCB = std::min(32, (need + NC - 1) / NC); NCB >= 1; --NCB) {
        if (ncb_env > 0 && NCB != ncb_env) continue;
        const int TW = NC * NCB;
        const int nVw = (NCB * G4 + 31) / 32, nWw = (TW + 31) / 32;
        int nSw = std::min(stagers_env, max_warps - (nVw + nWw));
        if (nSw < 2) continue;                               // one warp for the left rows, the others share the right rows
        if ((NCB * G4) & 31) continue;                       // whole VH warps only
        const int ncols = TW + 2 * r;
        P.ncolsP = ((ncols + 3) / 4) * 4 + 4;               // the VH threads read whole 16-byte groups of left words
        P.SWb = nd * 2 + 16; P.NK16 = (G4 + 3) / 4; P.KWb = (P.NK16 * 4 + 4) * 4;
        P.RLW = (std::max(ncols, TW + 12) + nd) / 4 + 4;    // right-row words: window bytes up to TW + nd + 4 * NRQ * 4
        if (P.ncolsP > 32 * 6 || P.RLW > 32 * 4 * (nSw - 1)) continue;   // register-prefetch limits of the stagers (MAXL, MAXR)
        int units = (4 * P.RLW + 15) / 16;
        while ((units & 3) != 2) ++units;
        P.CSB = units * 16;
        size_t o = 0;
        for (int s2 = 0; s2 < 2; ++s2) { P.oStage[s2] = (int)o; o += 2 * (size_t)P.ncolsP * 4 + 8 * (size_t)P.CSB; o = (o + 15) & ~(size_t)15; }
        P.oTc = (int)o; o += 8 * (size_t)ncols * 4; o = (o + 15) & ~(size_t)15;
        for (int s2 = 0; s2 < 2; ++s2) { P.oK[s2] = (int)o; o += (size_t)TW * P.KWb; }
        for (int s2 = 0; s2 < 2; ++s2) { P.oS[s2] = (int)o; o += (size_t)TW * P.SWb; }
        P.oMbar = (int)o; o += 8 * vh::MB_COUNT;
        P.oX = -1;
        if (use_hx && vh_hx_available(nd, wide) && NCB > 1) {
            o = (o + 15) & ~(size_t)15;
            const size_t xbytes = (size_t)2 * 2 * ((r + 3) / 4) * 16 * (size_t)(nVw * 32);
            if (o + xbytes <= smem_max) { P.oX = (int)o; o += xbytes; }
        }
        if (o > smem_max) continue;
        P.TW = TW; P.ncols = ncols; P.NCB = NCB; P.G4 = G4;
        P.nVw = nVw; P.nWw = nWw; P.nSw = nSw;
        const int tilesX = (need + TW - 1) / TW;
        const int max_bands = std::max(1, rows / (2 * r + 4));
        // Blocks that fit an SM together (registers of this instantiation, shared memory incl. the 1 KiB system reservation).
        const int nthr = 32 * (nVw + nWw + nSw);
        const int occ = std::max(1, std::min(std::min(gpu.regs_per_sm / (regs * nthr), (int)(gpu.smem_per_sm / (o + 1024))), gpu.threads_per_sm / nthr));
        for (int bands = 1; bands <= max_bands; ++bands) {
            if (bands_env > 0 && bands != std::min(bands_env, max_bands)) continue;
            const int BH = (rows + bands - 1) / bands;
            if ((wide ? 0 : 128 * (BH + 2 * r + 1)) + 2 * cfg.cap * (2 * r + 1) * (2 * r + 1) > 65535) continue;   // bias of the odd columns
            const int nb = tilesX * ((rows + BH - 1) / BH) * nf;
            const int per_sm = (nb + n_sm - 1) / n_sm;
            const int conc = std::min(occ, per_sm), waves = (per_sm + occ - 1) / occ;
            // Measured on B200 (tools/sweep_vh.sh, profiles/r01_v7_planner_sweep.md): a block row costs about
            // 0.35 + 0.33 * m us, m = VH warps per SM sub-partition (warps of co-resident blocks pile up on the same
            // sub-partitions), plus a little for every further warp and for the W warps; ~2 rows of pipeline fill.
            const int Veff = conc > 1 ? conc * 4 * ((nVw + 3) / 4) : nVw;
            const double krow = 0.35 + 0.33 * ((Veff + 3) / 4) + 0.03 * ((Veff + 3) % 4);
            const double cost = (double)waves * (BH + 2 * r + 2) * krow * (1.0 + 0.03 * conc * nWw) * (1.0 + 0.05 * (conc - 1));
            if (cost < best_cost) {
                best_cost = cost; best = P; best_bands = bands; smem = o;
                nt = nthr;
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
    P.pre_stride = pre_stride; P.disp_stride = disp_stride;
    const int tilesX = (need + P.TW - 1) / P.TW;
    P.BH = (rows + best_bands - 1) / best_bands;
    dim3 grid(tilesX, (rows + P.BH - 1) / P.BH, nf);
    if (verbose)
        fprintf(stderr, "bm_vh plan: NCB=%d TW=%d BH=%d grid=%dx%dx%d warps V/W/S=%d/%d/%d smem=%zu hx=%d\n", P.NCB, P.TW, P.BH, grid.x, grid.y,
                grid.z, P.nVw, P.nWw, P.nSw, smem, P.oX >= 0);
    cudaError_t e;
    switch (r) {
    case 2: e = launch_vh<2>(P, grid, nt, smem, st); break;
    case 3: e = launch_vh<3>(P, grid, nt, smem, st); break;
    case 4: e = launch_vh<4>(P, grid, nt, smem, st); break;
    case 5: e = launch_vh<5>(P, grid, nt, smem, st); break;
    case 6: e = launch_vh<6>(P, grid, nt, smem, st); break;
    case 7: e = launch_vh<7>(P, grid, nt, smem, st); break;
    case 8: e = launch_vh<8>(P, grid, nt, smem, st); break;
    case 9: e = launch_vh<9>(P, grid, nt, smem, st); break;
    case 10: e = launch_vh<10>(P, grid, nt, smem, st); break;
    default: return 0;
    }
    return e == cudaSuccess ? 1 : -1;
}

}  // namespace b200s
