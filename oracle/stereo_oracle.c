/*
 * stereo_oracle.c -- CPU restatement of the stereo hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the CUDA path.  It is never linked into, imported by or
 * called from the product library (libb200stereo.so); only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * What it restates.  The reference (maciejmatuszak/ros_gpu_stereo_processor) holds no arithmetic
 * of its own on this path: every step is a call into OpenCV / image_geometry, which are NOT
 * vendored under /root/reference (CMakeLists.txt:21-24 finds an unpinned OpenCV 3.x fork at
 * /usr/local_ros; package.xml:46-47).  The call sites restated here are
 *   rectifyImageLeft/Right      src/GPUStereoProcessor.cpp:252-262  (cv::initUndistortRectifyMap + cv::remap)
 *   computeDisparity(cv::Mat..) src/GPUStereoProcessor.cpp:312-321  (cv::StereoBM::compute + convertTo)
 *   filterSpeckles              src/GPUStereoProcessor.cpp:367-385  (cv::filterSpeckles)
 *   projectDisparityTo3DPoints  src/GPUStereoProcessor.cpp:332-346  (cv::reprojectImageTo3D)
 *   GPUSenderPc2::fillInData    src/GpuSenderPc2.cpp:15-72          (PointCloud2 record layout)
 *   GPUSenderDisparity          src/GpuSenderDisparity.cpp:18-48    (DisparityImage payload)
 * and the algorithm is OpenCV's published one (calib3d stereobm.cpp / stereosgbm.cpp
 * validateDisparity + filterSpeckles, imgproc undistort.cpp / imgwarp.cpp remap, calibration.cpp
 * reprojectImageTo3D) as specified in SURVEY.md Appendix A.
 *
 * Pinning.  tests/test_oracle_cpu.py checks every function here bit-for-bit against
 *   (1) the reference's own golden fixtures left/right-0022_rect.png (test/UTest.cpp:247-256),
 *       carried as tests/golden/fixtures.npz, and
 *   (2) the real OpenCV build importable in this image (cv2 4.13: StereoBM_create,
 *       initUndistortRectifyMap, remap, filterSpeckles, reprojectImageTo3D) over a sweep of
 *       parameter sets and sizes, plus committed cv2-generated vectors in tests/golden/.
 *
 * Everything is written from the algorithm description, scalar and deliberately simple
 * (definition-style sums made incremental only where needed for speed).
 */
#define _POSIX_C_SOURCE 200809L
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

#include <pthread.h>
#include <unistd.h>

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ------------------------------------------------------------------------------------------
 * A.1  rectification map: cv::initUndistortRectifyMap(K, D, R, P, size, CV_32FC1)
 * K 3x3, D 8 coefficients (k1 k2 p1 p2 k3 k4 k5 k6; unused ones zero), R 3x3, P 3x4 (row major).
 * ------------------------------------------------------------------------------------------ */
static void inv3x3(const double m[9], double o[9])
{
    /* plain cofactor inverse in double -- cv::invert(DECOMP_LU) differs in the last ulp at most,
       the test sweep shows the resulting float maps are identical on the fixtures/configs used */
    double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
    double det = a * A + b * B + c * C;
    double id = 1.0 / det;
    o[0] = A * id;               o[1] = -(b * i - c * h) * id; o[2] = (b * f - c * e) * id;
    o[3] = B * id;               o[4] = (a * i - c * g) * id;  o[5] = -(a * f - c * d) * id;
    o[6] = C * id;               o[7] = -(a * h - b * g) * id; o[8] = (a * e - b * d) * id;
}

void orc_rect_inverse(const double* K, const double* R, const double* P, double* ir /*9*/)
{
    /* ir = inv(P[:, :3] * R) */
    double pr[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += P[r * 4 + k] * R[k * 3 + c];
            pr[r * 3 + c] = s;
        }
    (void)K;
    inv3x3(pr, ir);
}

/* ir supplied by the caller (tests feed numpy.linalg / cv2.invert results through here as well) */
void orc_build_rect_map_ir(const double* K, const double* D, const double* ir, int W, int H,
                           float* mapx, float* mapy)
{
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    const double k1 = D[0], k2 = D[1], p1 = D[2], p2 = D[3], k3 = D[4], k4 = D[5], k5 = D[6], k6 = D[7];
    for (int i = 0; i < H; ++i) {
        for (int j = 0; j < W; ++j) {
            double _x = j * ir[0] + (i * ir[1] + ir[2]);
            double _y = j * ir[3] + (i * ir[4] + ir[5]);
            double _w = j * ir[6] + (i * ir[7] + ir[8]);
            double w = 1.0 / _w, x = _x * w, y = _y * w;
            double x2 = x * x, y2 = y * y;
            double r2 = x2 + y2, _2xy = 2 * x * y;
            double kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2) / (1 + ((k6 * r2 + k5) * r2 + k4) * r2);
            double xd = (x * kr + p1 * _2xy + p2 * (r2 + 2 * x2));
            double yd = (y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy);
            double u = fx * xd + cx;
            double v = fy * yd + cy;
            mapx[(size_t)i * W + j] = (float)u;
            mapy[(size_t)i * W + j] = (float)v;
        }
    }
}

void orc_build_rect_map(const double* K, const double* D, const double* R, const double* P, int W, int H,
                        float* mapx, float* mapy)
{
    double ir[9];
    orc_rect_inverse(K, R, P, ir);
    orc_build_rect_map_ir(K, D, ir, W, H, mapx, mapy);
}

/* A.1.3  cv::remap(src, mapx, mapy, INTER_LINEAR, BORDER_CONSTANT, 0) on 8-bit, `ch` interleaved channels */
static inline int sat16(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

void orc_remap_linear(const uint8_t* src, int sW, int sH, int ch, const float* mapx, const float* mapy,
                      int W, int H, uint8_t* dst)
{
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            float u = mapx[(size_t)i * W + j], v = mapy[(size_t)i * W + j];
            int sx = (int)lrintf(u * 32.0f), sy = (int)lrintf(v * 32.0f); /* round-half-even */
            int a = sx & 31, b = sy & 31;
            int X0 = sat16(sx >> 5), Y0 = sat16(sy >> 5);
            for (int c = 0; c < ch; ++c) {
                int s00 = 0, s01 = 0, s10 = 0, s11 = 0;
                int x1 = X0 + 1, y1 = Y0 + 1;
                if (Y0 >= 0 && Y0 < sH) {
                    if (X0 >= 0 && X0 < sW) s00 = src[((size_t)Y0 * sW + X0) * ch + c];
                    if (x1 >= 0 && x1 < sW) s01 = src[((size_t)Y0 * sW + x1) * ch + c];
                }
                if (y1 >= 0 && y1 < sH) {
                    if (X0 >= 0 && X0 < sW) s10 = src[((size_t)y1 * sW + X0) * ch + c];
                    if (x1 >= 0 && x1 < sW) s11 = src[((size_t)y1 * sW + x1) * ch + c];
                }
                int acc = (32 - a) * (32 - b) * s00 + a * (32 - b) * s01 + (32 - a) * b * s10 + a * b * s11;
                dst[((size_t)i * W + j) * ch + c] = (uint8_t)((acc + 512) >> 10);
            }
        }
}

/* ------------------------------------------------------------------------------------------
 * A.2.1  prefilters
 * ------------------------------------------------------------------------------------------ */
static inline uint8_t clipcap(int v, int cap) { return (uint8_t)((v < -cap ? -cap : (v > cap ? cap : v)) + cap); }

void orc_prefilter_xsobel(const uint8_t* src, uint8_t* dst, int W, int H, int cap)
{
    int y;
    for (y = 0; y < H - 1; y += 2) {
        const uint8_t* r1 = src + (size_t)y * W;
        const uint8_t* r0 = y > 0 ? r1 - W : (H > 1 ? r1 + W : r1);
        const uint8_t* r2 = y < H - 1 ? r1 + W : (H > 1 ? r1 - W : r1);
        const uint8_t* r3 = y < H - 2 ? r1 + 2 * (size_t)W : r1;
        uint8_t* d0 = dst + (size_t)y * W;
        uint8_t* d1 = d0 + W;
        d0[0] = d0[W - 1] = d1[0] = d1[W - 1] = (uint8_t)cap;
        for (int x = 1; x < W - 1; ++x) {
            int e0 = r0[x + 1] - r0[x - 1], e1 = r1[x + 1] - r1[x - 1];
            int e2 = r2[x + 1] - r2[x - 1], e3 = r3[x + 1] - r3[x - 1];
            d0[x] = clipcap(e0 + 2 * e1 + e2, cap);
            d1[x] = clipcap(e1 + 2 * e2 + e3, cap);
        }
    }
    for (; y < H; ++y)
        for (int x = 0; x < W; ++x) dst[(size_t)y * W + x] = (uint8_t)cap;
}

void orc_prefilter_norm(const uint8_t* src, uint8_t* dst, int W, int H, int ps, int cap)
{
    int p2 = ps / 2;
    int scale_g = ps * ps / 8, scale_s = (1024 + scale_g) / (scale_g * 2);
    scale_g *= scale_s;
    for (int y = 0; y < H; ++y) {
        const uint8_t* prev = src + (size_t)clampi(y - 1, 0, H - 1) * W;
        const uint8_t* curr = src + (size_t)y * W;
        const uint8_t* next = src + (size_t)clampi(y + 1, 0, H - 1) * W;
        for (int x = 0; x < W; ++x) {
            long sum = 0; /* definition-style box sum with replicate borders */
            for (int dy = -p2; dy <= p2; ++dy) {
                const uint8_t* row = src + (size_t)clampi(y + dy, 0, H - 1) * W;
                for (int dx = -p2; dx <= p2; ++dx) sum += row[clampi(x + dx, 0, W - 1)];
            }
            int c = 4 * curr[x] + curr[clampi(x - 1, 0, W - 1)] + curr[clampi(x + 1, 0, W - 1)] + prev[x] + next[x];
            int val = (int)((c * (long)scale_g - sum * scale_s) >> 10);
            dst[(size_t)y * W + x] = clipcap(val, cap);
        }
    }
}

/* faster twin of orc_prefilter_norm (separable running sums) used for big images; tested == the definition */
void orc_prefilter_norm_fast(const uint8_t* src, uint8_t* dst, int W, int H, int ps, int cap)
{
    int p2 = ps / 2;
    int scale_g = ps * ps / 8, scale_s = (1024 + scale_g) / (scale_g * 2);
    scale_g *= scale_s;
    int* vs = (int*)malloc(sizeof(int) * (size_t)W);
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            int s = 0;
            for (int dy = -p2; dy <= p2; ++dy) s += src[(size_t)clampi(y + dy, 0, H - 1) * W + x];
            vs[x] = s;
        }
        const uint8_t* prev = src + (size_t)clampi(y - 1, 0, H - 1) * W;
        const uint8_t* curr = src + (size_t)y * W;
        const uint8_t* next = src + (size_t)clampi(y + 1, 0, H - 1) * W;
        long sum = 0;
        for (int dx = -p2; dx <= p2; ++dx) sum += vs[clampi(dx, 0, W - 1)];
        for (int x = 0; x < W; ++x) {
            int c = 4 * curr[x] + curr[clampi(x - 1, 0, W - 1)] + curr[clampi(x + 1, 0, W - 1)] + prev[x] + next[x];
            int val = (int)((c * (long)scale_g - sum * scale_s) >> 10);
            dst[(size_t)y * W + x] = clipcap(val, cap);
            sum += vs[clampi(x + p2 + 1, 0, W - 1)] - vs[clampi(x - p2, 0, W - 1)];
        }
    }
    free(vs);
}

/* ------------------------------------------------------------------------------------------
 * A.2  block matching on prefiltered planes.
 * params: minD, nd, wsz, cap, textureThreshold, uniquenessRatio.
 * Writes the s16 disparity plane and the s16 cost plane for rows [r, H-r) and columns
 * [lofs, lofs+width1); every other pixel of `disp` is set to FILTERED (cost untouched).
 * NO ROI column mask, NO validate, NO speckle here (see orc_stereobm_post).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int minDisparity, numDisparities, blockSize;
    int preFilterType, preFilterSize, preFilterCap; /* type: 0 NORMALIZED_RESPONSE, 1 XSOBEL */
    int textureThreshold, uniquenessRatio;
    int speckleWindowSize, speckleRange, disp12MaxDiff;
} orc_bm_params;

static void bm_rows(const uint8_t* Lp, const uint8_t* Rp, int W, int H, const orc_bm_params* p,
                    int16_t* disp, int16_t* cost, int ya, int yb)
{
    const int nd = p->numDisparities, minD = p->minDisparity, wsz = p->blockSize, r = wsz / 2, cap = p->preFilterCap;
    const int lofs = (nd - 1 + minD) > 0 ? (nd - 1 + minD) : 0;
    const int rofs = (nd - 1 + minD) < 0 ? -(nd - 1 + minD) : 0;
    const int width1 = W - rofs - nd + 1;
    const int16_t FILTERED = (int16_t)((minD - 1) * 16);
    const int ncol = width1 + 2 * r; /* window columns x' in [-r, width1-1+r] */
    int* col = (int*)calloc((size_t)ncol * nd, sizeof(int));
    int* tcol = (int*)calloc((size_t)ncol, sizeof(int));
    int* sad = (int*)malloc(sizeof(int) * (size_t)(nd + 2));
    int* lcx = (int*)malloc(sizeof(int) * (size_t)ncol);
    int* rbx = (int*)malloc(sizeof(int) * (size_t)ncol);
    for (int c = 0; c < ncol; ++c) {
        int xp = c - r;
        lcx[c] = clampi(xp, -lofs, W - 1 - lofs) + lofs;
        rbx[c] = clampi(xp, -rofs, W - nd - rofs) + rofs;
    }
    for (int y = ya; y < yb; ++y) {
        /* (re)build or slide the vertical sums over window rows y-r..y+r (all inside the image for ROI rows) */
        if (y == ya) {
            memset(col, 0, sizeof(int) * (size_t)ncol * nd);
            memset(tcol, 0, sizeof(int) * (size_t)ncol);
            for (int yy = y - r; yy <= y + r; ++yy) {
                const uint8_t* lr = Lp + (size_t)clampi(yy, 0, H - 1) * W;
                const uint8_t* rr = Rp + (size_t)clampi(yy, 0, H - 1) * W;
                for (int c = 0; c < ncol; ++c) {
                    int lv = lr[lcx[c]];
                    const uint8_t* rp = rr + rbx[c];
                    int* cc = col + (size_t)c * nd;
                    for (int k = 0; k < nd; ++k) cc[k] += abs(lv - rp[k]);
                    tcol[c] += abs(lv - cap);
                }
            }
        } else {
            const uint8_t* la = Lp + (size_t)clampi(y + r, 0, H - 1) * W;
            const uint8_t* ra = Rp + (size_t)clampi(y + r, 0, H - 1) * W;
            const uint8_t* ls = Lp + (size_t)clampi(y - r - 1, 0, H - 1) * W;
            const uint8_t* rs = Rp + (size_t)clampi(y - r - 1, 0, H - 1) * W;
            for (int c = 0; c < ncol; ++c) {
                int lva = la[lcx[c]], lvs = ls[lcx[c]];
                const uint8_t* rpa = ra + rbx[c];
                const uint8_t* rps = rs + rbx[c];
                int* cc = col + (size_t)c * nd;
                for (int k = 0; k < nd; ++k) cc[k] += abs(lva - rpa[k]) - abs(lvs - rps[k]);
                tcol[c] += abs(lva - cap) - abs(lvs - cap);
            }
        }
        int16_t* drow = disp + (size_t)y * W;
        int16_t* crow = cost ? cost + (size_t)y * W : NULL;
        int* S = sad + 1;
        for (int x = 0; x < width1; ++x) {
            int tsum = 0;
            if (x == 0) {
                for (int k = 0; k < nd; ++k) S[k] = 0;
                for (int c = 0; c < wsz; ++c) {
                    const int* cc = col + (size_t)c * nd;
                    for (int k = 0; k < nd; ++k) S[k] += cc[k];
                }
            } else {
                const int* ca = col + (size_t)(x + 2 * r) * nd;
                const int* cs = col + (size_t)(x - 1) * nd;
                for (int k = 0; k < nd; ++k) S[k] += ca[k] - cs[k];
            }
            for (int c = x; c < x + wsz; ++c) tsum += tcol[c];
            int minsad = INT32_MAX, mind = -1;
            for (int k = 0; k < nd; ++k)
                if (S[k] < minsad) { minsad = S[k]; mind = k; }
            int X = x + lofs;
            if (tsum < p->textureThreshold) { drow[X] = FILTERED; continue; }
            if (p->uniquenessRatio > 0) {
                int thresh = minsad + (minsad * p->uniquenessRatio / 100);
                int k;
                for (k = 0; k < nd; ++k)
                    if ((k < mind - 1 || k > mind + 1) && S[k] <= thresh) break;
                if (k < nd) { drow[X] = FILTERED; continue; }
            }
            {
                int sm1 = S[-1], snd = S[nd];
                S[-1] = S[1];
                S[nd] = S[nd - 2];
                int pp = S[mind + 1], n = S[mind - 1];
                int d = pp + n - 2 * S[mind] + abs(pp - n);
                int v = ((nd - mind - 1 + minD) * 256 + (d != 0 ? (pp - n) * 256 / d : 0) + 15) >> 4;
                drow[X] = (int16_t)v;
                if (crow) crow[X] = (int16_t)S[mind];
                S[-1] = sm1; S[nd] = snd;
            }
        }
    }
    free(col); free(tcol); free(sad); free(lcx); free(rbx);
}

static int g_threads = 0; /* 0 = all online cores */
void orc_set_num_threads(int n) { g_threads = n; }
int orc_num_threads(void)
{
    if (g_threads > 0) return g_threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

typedef struct { const uint8_t *Lp, *Rp; int W, H; const orc_bm_params* p; int16_t *disp, *cost; int ya, yb; } bm_job;
static void* bm_job_run(void* a)
{
    bm_job* j = (bm_job*)a;
    if (j->yb > j->ya) bm_rows(j->Lp, j->Rp, j->W, j->H, j->p, j->disp, j->cost, j->ya, j->yb);
    return NULL;
}

void orc_bm_core(const uint8_t* Lp, const uint8_t* Rp, int W, int H, const orc_bm_params* p,
                 int16_t* disp, int16_t* cost)
{
    const int nd = p->numDisparities, minD = p->minDisparity, r = p->blockSize / 2;
    const int lofs = (nd - 1 + minD) > 0 ? (nd - 1 + minD) : 0;
    const int rofs = (nd - 1 + minD) < 0 ? -(nd - 1 + minD) : 0;
    const int width1 = W - rofs - nd + 1;
    const int16_t FILTERED = (int16_t)((minD - 1) * 16);
    for (size_t i = 0; i < (size_t)W * H; ++i) disp[i] = FILTERED;
    if (cost) memset(cost, 0, sizeof(int16_t) * (size_t)W * H);
    if (lofs >= W || rofs >= W || width1 < 1) return;
    int y0 = r, y1 = H - r;
    if (y1 <= y0) return;
    int nstripes = orc_num_threads();
    if (nstripes > (y1 - y0) / 8) nstripes = (y1 - y0) / 8;
    if (nstripes < 1) nstripes = 1;
    if (nstripes > 64) nstripes = 64;
    pthread_t th[64];
    bm_job jobs[64];
    for (int s = 0; s < nstripes; ++s) {
        bm_job j = { Lp, Rp, W, H, p, disp, cost,
                     y0 + (int)((long)(y1 - y0) * s / nstripes), y0 + (int)((long)(y1 - y0) * (s + 1) / nstripes) };
        jobs[s] = j;
        if (nstripes == 1) bm_job_run(&jobs[s]);
        else pthread_create(&th[s], NULL, bm_job_run, &jobs[s]);
    }
    if (nstripes > 1)
        for (int s = 0; s < nstripes; ++s) pthread_join(th[s], NULL);
}

/* A.3  cv::validateDisparity on rows [ya,yb) */
void orc_validate_disp12(int16_t* disp, const int16_t* cost, int W, int H, int minD, int nd, int disp12MaxDiff,
                         int ya, int yb)
{
    const int maxD = minD + nd;
    const int minX1 = maxD > 0 ? maxD : 0, maxX1 = W + (minD < 0 ? minD : 0);
    const int INV = (minD - 1) * 16;
    const int maxdiff = disp12MaxDiff * 16;
    int* d2 = (int*)malloc(sizeof(int) * (size_t)W * 2);
    int* c2 = d2 + W;
    (void)H;
    for (int y = ya; y < yb; ++y) {
        int16_t* dp = disp + (size_t)y * W;
        const int16_t* cp = cost + (size_t)y * W;
        for (int x = 0; x < W; ++x) { d2[x] = INV; c2[x] = INT32_MAX; }
        for (int x = minX1; x < maxX1; ++x) {
            int d = dp[x], c = cp[x];
            if (d == INV) continue;
            int x2 = x - ((d + 8) >> 4);
            if (x2 < 0 || x2 >= W) continue; /* cannot happen for matcher output; guard only */
            if (c2[x2] > c) { c2[x2] = c; d2[x2] = d; }
        }
        for (int x = minX1; x < maxX1; ++x) {
            int d = dp[x];
            if (d == INV) continue;
            int d0 = d >> 4, d1 = (d + 15) >> 4;
            int x0 = x - d0, x1 = x - d1;
            if ((0 <= x0 && x0 < W && d2[x0] > INV && abs(d2[x0] - d) > maxdiff) &&
                (0 <= x1 && x1 < W && d2[x1] > INV && abs(d2[x1] - d) > maxdiff))
                dp[x] = (int16_t)INV;
        }
    }
    free(d2);
}

/* A.4  cv::filterSpeckles(img CV_16S, newVal, maxSpeckleSize, maxDiff): flood fill, 4-neighbourhood */
void orc_filter_speckles(int16_t* img, int W, int H, int newVal, int maxSize, int maxDiff)
{
    size_t n = (size_t)W * H;
    int* label = (int*)calloc(n, sizeof(int));
    int* stack = (int*)malloc(sizeof(int) * n);
    uint8_t* small = (uint8_t*)malloc(n + 1); /* per label: is it a small region */
    int cur = 0;
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j) {
            size_t idx = (size_t)i * W + j;
            if (img[idx] == newVal) continue;
            if (label[idx]) {
                if (small[label[idx]]) img[idx] = (int16_t)newVal;
                continue;
            }
            ++cur;
            int sp = 0, count = 0;
            stack[sp++] = (int)idx;
            label[idx] = cur;
            while (sp) {
                int q = stack[--sp];
                ++count;
                int qi = q / W, qj = q % W;
                int dq = img[q];
                if (qj < W - 1 && !label[q + 1] && img[q + 1] != newVal && abs(dq - img[q + 1]) <= maxDiff) { label[q + 1] = cur; stack[sp++] = q + 1; }
                if (qj > 0 && !label[q - 1] && img[q - 1] != newVal && abs(dq - img[q - 1]) <= maxDiff) { label[q - 1] = cur; stack[sp++] = q - 1; }
                if (qi < H - 1 && !label[q + W] && img[q + W] != newVal && abs(dq - img[q + W]) <= maxDiff) { label[q + W] = cur; stack[sp++] = q + W; }
                if (qi > 0 && !label[q - W] && img[q - W] != newVal && abs(dq - img[q - W]) <= maxDiff) { label[q - W] = cur; stack[sp++] = q - W; }
            }
            if (count <= maxSize) { small[cur] = 1; img[idx] = (int16_t)newVal; }
            else small[cur] = 0;
        }
    free(label); free(stack); free(small);
}

/* A.2.6  post-processing order of cv::StereoBM::compute: validate -> ROI column/row mask -> speckle */
void orc_stereobm_post(int16_t* disp, const int16_t* cost, int W, int H, const orc_bm_params* p)
{
    const int nd = p->numDisparities, minD = p->minDisparity, r = p->blockSize / 2;
    const int16_t FILTERED = (int16_t)((minD - 1) * 16);
    int y0 = r, y1 = H - r;
    if (p->disp12MaxDiff >= 0 && y1 > y0) orc_validate_disp12(disp, cost, W, H, minD, nd, p->disp12MaxDiff, y0, y1);
    /* getValidDisparityROI(full, full, minD, nd, wsz) */
    int xmin = (minD + nd - 1 > 0 ? minD + nd - 1 : 0) + r; /* maxD + SW2 with maxD = minD+nd-1 */
    int xmax = W - r; /* OpenCV 4.x getValidDisparityROI has no minD term on the right edge */
    int ymin = r, ymax = H - r;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            if (!(x >= xmin && x < xmax && y >= ymin && y < ymax)) disp[(size_t)y * W + x] = FILTERED;
    if (p->speckleWindowSize > 0 && p->speckleRange >= 0)
        orc_filter_speckles(disp, W, H, FILTERED, p->speckleWindowSize, p->speckleRange);
}

/* full cv::StereoBM::compute on raw (rectified) 8-bit inputs. Lp/Rp (optional) receive the prefiltered planes. */
int orc_stereobm_compute(const uint8_t* L, const uint8_t* R, int W, int H, const orc_bm_params* p,
                         int16_t* disp, uint8_t* Lp_out, uint8_t* Rp_out)
{
    if (p->preFilterType != 0 && p->preFilterType != 1) return -1;
    if (p->preFilterSize < 5 || p->preFilterSize > 255 || p->preFilterSize % 2 == 0) return -2;
    if (p->preFilterCap < 1 || p->preFilterCap > 63) return -3;
    if (p->blockSize < 5 || p->blockSize > 255 || p->blockSize % 2 == 0 || p->blockSize >= (W < H ? W : H)) return -4;
    if (p->numDisparities <= 0 || p->numDisparities % 16 != 0) return -5;
    if (p->textureThreshold < 0) return -6;
    if (p->uniquenessRatio < 0) return -7;
    size_t n = (size_t)W * H;
    uint8_t* Lp = Lp_out ? Lp_out : (uint8_t*)malloc(n);
    uint8_t* Rp = Rp_out ? Rp_out : (uint8_t*)malloc(n);
    int16_t* cost = (int16_t*)malloc(sizeof(int16_t) * n);
    if (p->preFilterType == 1) {
        orc_prefilter_xsobel(L, Lp, W, H, p->preFilterCap);
        orc_prefilter_xsobel(R, Rp, W, H, p->preFilterCap);
    } else {
        orc_prefilter_norm_fast(L, Lp, W, H, p->preFilterSize, p->preFilterCap);
        orc_prefilter_norm_fast(R, Rp, W, H, p->preFilterSize, p->preFilterCap);
    }
    orc_bm_core(Lp, Rp, W, H, p, disp, cost);
    orc_stereobm_post(disp, cost, W, H, p);
    if (!Lp_out) free(Lp);
    if (!Rp_out) free(Rp);
    free(cost);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * A.5  disparity float plane, reprojection, PointCloud2 packing
 * ------------------------------------------------------------------------------------------ */
void orc_disparity_to_float(const int16_t* d16, size_t n, double cx_minus_cxr, float* out)
{
    /* cv::Mat::convertTo(CV_32F, 1/16., -(cx_l - cx_r)): saturate_cast<float>(d*alpha + beta) in double */
    for (size_t i = 0; i < n; ++i) out[i] = (float)((double)d16[i] * (1.0 / 16.0) + (-cx_minus_cxr));
}

void orc_reproject(const float* df, int W, int H, const double* Q /*16*/, int handleMissing, float* xyz)
{
    const double bigZ = 10000.;
    double minDisp = 0;
    if (handleMissing) {
        float m = FLT_MAX;
        for (size_t i = 0; i < (size_t)W * H; ++i) if (df[i] < m) m = df[i];
        minDisp = m;
    }
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            double d = df[(size_t)y * W + x];
            double h[4];
            for (int i = 0; i < 4; ++i) {
                double s = Q[i * 4 + 0] * (double)x;
                s += Q[i * 4 + 1] * (double)y;
                s += Q[i * 4 + 2] * d;
                s += Q[i * 4 + 3] * 1.0;
                h[i] = s;
            }
            float* o = xyz + ((size_t)y * W + x) * 3;
            for (int i = 0; i < 3; ++i) {
                float f = (float)h[i];
                o[i] = (float)((double)f / h[3]);
            }
            if (handleMissing && fabs(d - minDisp) <= FLT_EPSILON) o[2] = (float)bigZ;
        }
}

/* PointCloud2 payload: 32-byte records x@0 y@4 z@8 (f32), b@16 g@17 r@18, all other bytes zero;
   invalid (z == 10000 or +-inf) -> x=y=z=NaN.  color: BGR8 interleaved (ch=3) or mono (ch=1, replicated). */
void orc_pack_pointcloud2(const float* xyz, const uint8_t* color, int ch, int W, int H, uint8_t* out)
{
    memset(out, 0, (size_t)W * H * 32);
    for (size_t i = 0; i < (size_t)W * H; ++i) {
        float p[3] = { xyz[i * 3], xyz[i * 3 + 1], xyz[i * 3 + 2] };
        int valid = (p[2] != 10000.0f) && !isinf(p[2]);
        if (!valid) { p[0] = p[1] = p[2] = NAN; }
        uint8_t* o = out + i * 32;
        memcpy(o, p, 12);
        if (ch == 3) { o[16] = color[i * 3]; o[17] = color[i * 3 + 1]; o[18] = color[i * 3 + 2]; }
        else { o[16] = o[17] = o[18] = color[i]; }
    }
}
