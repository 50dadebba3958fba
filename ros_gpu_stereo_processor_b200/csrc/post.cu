// Post-filters of cv::StereoBM::compute for sm_100a: L/R consistency (cv::validateDisparity), valid-ROI mask,
// speckle filter (cv::filterSpeckles) as connected-component labelling.  SURVEY.md A.2.6, A.3, A.4.
// Reference call sites: block_matcher_cpu_->setDisp12MaxDiff (src/GPUStereoProcessor.cpp:23),
// filterSpeckles (src/GPUStereoProcessor.cpp:356-385, a host flood fill behind a stream sync in the reference).
#include "kernels.h"

#include <algorithm>

namespace b200s {

// ---- validateDisparity: one block per row ---------------------------------------------------------------
// pass 1 is a scatter-min: the sequential rule "c2[x2] > c (strict), x ascending" keeps the smallest cost and,
// among equal costs, the smallest x  ==  atomicMin over the key (cost + 32768) << 16 | x.
__global__ void __launch_bounds__(256) validate_kernel(int16_t* __restrict__ disp, const int16_t* __restrict__ cost,
                                                       int W, int minD, int nd, int maxdiff16, int ya, int roiX0, int roiX1)
{
    extern __shared__ uint32_t vsm[];
    uint32_t* key = vsm;                 // [W]
    int16_t* d2 = (int16_t*)(vsm + W);   // [W]
    const int y = ya + blockIdx.x;
    int16_t* dp = disp + (size_t)y * W;
    const int16_t* cp = cost + (size_t)y * W;
    const int INV = (minD - 1) * 16;
    const int maxD = minD + nd;
    const int minX1 = max(maxD, 0), maxX1 = W + min(minD, 0);
    for (int x = threadIdx.x; x < W; x += blockDim.x) key[x] = 0xFFFFFFFFu;
    __syncthreads();
    for (int x = minX1 + threadIdx.x; x < maxX1; x += blockDim.x) {
        int d = dp[x];
        if (d == INV) continue;
        int x2 = x - ((d + 8) >> 4);
        if ((unsigned)x2 >= (unsigned)W) continue;
        uint32_t k = ((uint32_t)((int)cp[x] + 32768) << 16) | (uint32_t)x;
        atomicMin(&key[x2], k);
    }
    __syncthreads();
    for (int x = threadIdx.x; x < W; x += blockDim.x) {
        uint32_t k = key[x];
        d2[x] = (k == 0xFFFFFFFFu) ? (int16_t)INV : dp[k & 0xFFFFu];
    }
    __syncthreads();
    for (int x = minX1 + threadIdx.x; x < maxX1; x += blockDim.x) {
        int d = dp[x];
        if (d == INV) continue;
        int d0 = d >> 4, d1 = (d + 15) >> 4;
        int x0 = x - d0, x1 = x - d1;
        bool b0 = (unsigned)x0 < (unsigned)W && d2[x0] > INV && abs((int)d2[x0] - d) > maxdiff16;
        bool b1 = (unsigned)x1 < (unsigned)W && d2[x1] > INV && abs((int)d2[x1] - d) > maxdiff16;
        if (b0 && b1) dp[x] = (int16_t)INV;
    }
    // valid-ROI mask of the row (cv::StereoBM applies it after the L/R check): the columns outside [roiX0, roiX1) were
    // only computed to feed the check; nobody reads another thread's dp[] any more at this point
    for (int x = threadIdx.x; x < W; x += blockDim.x)
        if (x < roiX0 || x >= roiX1) dp[x] = (int16_t)INV;
}

__global__ void __launch_bounds__(256) roi_mask_kernel(int16_t* __restrict__ disp, int W, int H, int x0, int x1, int y0,
                                                       int y1, int16_t filtered)
{
    int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    if (!(x >= x0 && x < x1 && y >= y0 && y < y1)) disp[(size_t)y * W + x] = filtered;
}

// ---- speckle filter: union-find connected components over the 4-neighbourhood ---------------------------
// nodes: pixels != newVal; edge between 4-neighbours when |a - b| <= maxDiff.  A component of size <= maxSize is
// set to newVal.  Component membership and sizes are order independent, so this equals OpenCV's flood fill.
__device__ __forceinline__ int uf_find(int* L, int i)
{
    int p = L[i];
    while (p != i) {
        int gp = L[p];
        if (gp != p) L[i] = gp;   // path halving (benign race: only ever shortens towards a root)
        i = p;
        p = gp;
    }
    return i;
}

__device__ __forceinline__ void uf_union(int* L, int a, int b)
{
    while (true) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }   // a > b: hook the larger root under the smaller
        int old = atomicCAS(&L[a], a, b);   // succeeds only while a is still a root
        if (old == a) return;
        a = old;   // somebody else hooked a first; retry from its new parent
    }
}

// ---- stage 1: connected components inside 64x16 tiles, entirely in shared memory -------------------------------
constexpr int CTX = 64, CTY = 16;

__device__ __forceinline__ int sm_find(volatile int* L, int i)
{
    int p = L[i];
    while (p != i) {
        int gp = L[p];
        if (gp != p) L[i] = gp;   // path halving (a stale write still points at an ancestor)
        i = p;
        p = gp;
    }
    return i;
}

__device__ __forceinline__ void sm_union(int* L, int a, int b)
{
    while (true) {
        a = sm_find(L, a);
        b = sm_find(L, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }
        int old = atomicCAS(&L[a], a, b);
        if (old == a) return;
        a = old;
    }
}

// Stage 1.  Rows first: a warp scans two rows and labels every pixel with the first pixel of its horizontal run
// (inclusive max-scan of run starts with shuffles), so only run pairs -- not pixels -- need a union in the vertical pass.
// Every run end then adds its run length to the tile-local root's pixel count (one shared-memory atomic per run, not per
// pixel).  Writes L[i] = global index of the tile-local root (or -1 for newVal pixels) and sz[i] = pixel count of the
// tile-local component at its root, 0 everywhere else.
// Batches: the frame index is the last used grid dimension of every ccl kernel; `ist` / `sst` are the byte distances
// between the image planes and between the scratch blocks (L, sz) of consecutive frames.
__global__ void __launch_bounds__(256) ccl_local_kernel(const int16_t* __restrict__ img, int* __restrict__ L,
                                                        int* __restrict__ sz, int W, int H, int newVal, int maxDiff,
                                                        size_t ist, size_t sst)
{
    img = (const int16_t*)((const uint8_t*)img + blockIdx.z * ist);
    L = (int*)((uint8_t*)L + blockIdx.z * sst);
    sz = (int*)((uint8_t*)sz + blockIdx.z * sst);
    __shared__ int16_t v[CTY * CTX];
    __shared__ int16_t rs[CTY * CTX];     // immutable run start (tile index) of every pixel, -1 for newVal
    __shared__ int lab[CTY * CTX];        // union-find parents over tile indices (only run starts ever get hooked)
    __shared__ int cnt[CTY * CTX];        // pixels per tile-local root
    const int x0 = blockIdx.x * CTX, y0 = blockIdx.y * CTY;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int idx = threadIdx.x + 256 * k;
        const int tx = idx & (CTX - 1), ty = idx >> 6;
        const int x = x0 + tx, y = y0 + ty;
        int val = newVal;
        if (x < W && y < H) val = img[(size_t)y * W + x];
        v[idx] = (int16_t)val;
        cnt[idx] = 0;
    }
    __syncthreads();
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int base = (2 * warp + rr) * CTX;
            const int a0 = v[base + 2 * lane], a1 = v[base + 2 * lane + 1];
            const int left = __shfl_up_sync(0xffffffffu, a1, 1);
            const bool val0 = a0 != newVal, val1 = a1 != newVal;
            const bool conn0 = val0 && lane > 0 && left != newVal && abs(a0 - left) <= maxDiff;
            const bool conn1 = val1 && val0 && abs(a1 - a0) <= maxDiff;
            const int s0 = conn0 ? -1 : 2 * lane, s1 = conn1 ? -1 : 2 * lane + 1;
            int e = max(s0, s1);
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, e, d);
                if (lane >= d) e = max(e, t);
            }
            int E = __shfl_up_sync(0xffffffffu, e, 1);
            if (lane == 0) E = -1;
            const int st0 = max(E, s0), st1 = s1 >= 0 ? s1 : st0;
            rs[base + 2 * lane] = (int16_t)(val0 ? base + st0 : -1);
            rs[base + 2 * lane + 1] = (int16_t)(val1 ? base + st1 : -1);
            lab[base + 2 * lane] = val0 ? base + st0 : -1;
            lab[base + 2 * lane + 1] = val1 ? base + st1 : -1;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int idx = threadIdx.x + 256 * k;
        const int tx = idx & (CTX - 1), ty = idx >> 6;
        if (ty + 1 >= CTY) continue;
        const int a = v[idx], u = v[idx + CTX];
        if (a == newVal || u == newVal || abs(u - a) > maxDiff) continue;
        // one union per pair of touching runs: only where the pair of runs starts to be connected
        bool first = tx == 0 || rs[idx] != rs[idx - 1] || rs[idx + CTX] != rs[idx + CTX - 1];
        if (!first) {
            const int al = v[idx - 1], ul = v[idx + CTX - 1];
            first = abs(ul - al) > maxDiff;     // both left pixels are valid here (they are in the same runs)
        }
        if (first) sm_union(lab, rs[idx], rs[idx + CTX]);
    }
    __syncthreads();
    int lroot[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int idx = threadIdx.x + 256 * k;
        const int tx = idx & (CTX - 1);
        lroot[k] = -1;
        if (rs[idx] >= 0) {
            lroot[k] = sm_find(lab, rs[idx]);
            if (tx == CTX - 1 || rs[idx + 1] != rs[idx]) atomicAdd(&cnt[lroot[k]], idx - rs[idx] + 1);   // run end
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int idx = threadIdx.x + 256 * k;
        const int tx = idx & (CTX - 1), ty = idx >> 6;
        const int x = x0 + tx, y = y0 + ty;
        if (x >= W || y >= H) continue;
        const int r = lroot[k];
        // raster order is preserved: root index <= own index
        L[(size_t)y * W + x] = r >= 0 ? (y0 + (r >> 6)) * W + x0 + (r & (CTX - 1)) : -1;
        sz[(size_t)y * W + x] = r == idx ? cnt[idx] : 0;
    }
}

// ---- stage 2: unions across tile borders (global union-find), one thread per border pixel ------------------------
__global__ void __launch_bounds__(256) ccl_border_kernel(const int16_t* __restrict__ img, int* __restrict__ L, int W,
                                                         int H, int newVal, int maxDiff, size_t ist, size_t sst)
{
    img = (const int16_t*)((const uint8_t*)img + blockIdx.y * ist);
    L = (int*)((uint8_t*)L + blockIdx.y * sst);
    const int ncx = (W - 1) / CTX;            // tile columns that have a right neighbour
    const int ncy = (H - 1) / CTY;            // tile rows that have a lower neighbour
    const long long nR = (long long)ncx * H, total = nR + (long long)ncy * W;
    long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    if (t >= total) return;
    int x, y, j, back;       // back: offset to the previous pixel pair along the border, 0 at the start of a tile
    if (t < nR) {
        y = (int)(t / ncx);
        x = ((int)(t - (long long)y * ncx) + 1) * CTX - 1;
        j = y * W + x + 1;
        back = (y % CTY) ? W : 0;
    } else {
        t -= nR;
        const int ry = (int)(t / W);
        x = (int)(t - (long long)ry * W);
        y = (ry + 1) * CTY - 1;
        j = (y + 1) * W + x;
        back = (x % CTX) ? 1 : 0;
    }
    const int i = y * W + x;
    const int a = img[i], u = img[j];
    if (a == newVal || u == newVal || abs(u - a) > maxDiff) return;
    if (back) {
        // the previous pair is joined by its own thread and both pixels are joined to their predecessors inside their
        // tiles (never across a tile corner, so there is no cycle of skipped edges): this union would be redundant
        const int ap = img[i - back], up = img[j - back];
        if (ap != newVal && up != newVal && abs(ap - a) <= maxDiff && abs(up - u) <= maxDiff && abs(up - ap) <= maxDiff) return;
    }
    uf_union(L, i, j);
}

// ---- stage 3: every tile-local root (sz > 0) that was hooked under another root adds its count to the global root
// and is pointed straight at it, so that afterwards L[L[i]] is the global root of any pixel i ----------------------
__global__ void __launch_bounds__(256) ccl_merge_counts_kernel(int* __restrict__ L, int* __restrict__ sz, int n, size_t sst)
{
    L = (int*)((uint8_t*)L + blockIdx.y * sst);
    sz = (int*)((uint8_t*)sz + blockIdx.y * sst);
    // four pixels per thread (one 16-byte load of the counts): roots are sparse, most threads stop here
    const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= n) return;
    int c4[4] = {0, 0, 0, 0};
    if (i0 + 4 <= n) {
        const int4 q = *(const int4*)(sz + i0);
        c4[0] = q.x; c4[1] = q.y; c4[2] = q.z; c4[3] = q.w;
    } else {
        for (int k = 0; i0 + k < n; ++k) c4[k] = sz[i0 + k];
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c4[k], i = i0 + k;
        if (c <= 0) continue;
        const int p = L[i];
        if (p == i) continue;                     // a global root keeps its own count
        int g = p;
        for (int q = L[g]; q != g; q = L[g]) g = q;
        atomicAdd(&sz[g], c);                     // only global roots ever receive additions
        L[i] = g;                                 // concurrent walkers see the old parent or the root: both are ancestors
    }
}

__global__ void __launch_bounds__(256) ccl_apply_kernel(int16_t* __restrict__ img, const int* __restrict__ L,
                                                        const int* __restrict__ sz, int n, int newVal, int maxSize,
                                                        size_t ist, size_t sst)
{
    img = (int16_t*)((uint8_t*)img + blockIdx.y * ist);
    L = (const int*)((const uint8_t*)L + blockIdx.y * sst);
    sz = (const int*)((const uint8_t*)sz + blockIdx.y * sst);
    const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= n) return;
    int p4[4] = {-1, -1, -1, -1};
    if (i0 + 4 <= n) {
        const int4 q = *(const int4*)(L + i0);    // tile-local roots (parents are always tile-local roots)
        p4[0] = q.x; p4[1] = q.y; p4[2] = q.z; p4[3] = q.w;
    } else {
        for (int k = 0; i0 + k < n; ++k) p4[k] = L[i0 + k];
    }
    int g4[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) g4[k] = p4[k] >= 0 ? L[p4[k]] : -1;          // global roots after stage 3
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (g4[k] >= 0 && sz[g4[k]] <= maxSize) img[i0 + k] = (int16_t)newVal;
}

static inline dim3 grid2d(int W, int H) { return dim3((W + 31) / 32, (H + 7) / 8); }

int launch_validate_disp12(int16_t* disp, const int16_t* cost, int W, int H, const BMConfig& cfg, cudaStream_t st)
{
    int r = cfg.wsz / 2;
    int y0 = r, y1 = H - r;
    if (y1 <= y0 || W > 65535) return y1 <= y0 ? 0 : -1;
    size_t smem = (size_t)W * 4 + (size_t)W * 2 + 16;
    if (smem > 48 * 1024) cudaFuncSetAttribute(validate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int x0 = std::max(cfg.minD + cfg.nd - 1, 0) + r, x1 = W - r;
    validate_kernel<<<y1 - y0, 256, smem, st>>>(disp, cost, W, cfg.minD, cfg.nd, cfg.disp12MaxDiff * 16, y0, x0, x1);
    return 1;
}

int launch_roi_mask(int16_t* disp, int W, int H, const BMConfig& cfg, cudaStream_t st)
{
    int r = cfg.wsz / 2;
    int x0 = std::max(cfg.minD + cfg.nd - 1, 0) + r, x1 = W - r, y0 = r, y1 = H - r;
    roi_mask_kernel<<<grid2d(W, H), 256, 0, st>>>(disp, W, H, x0, x1, y0, y1, (int16_t)((cfg.minD - 1) * 16));
    return 1;
}

int launch_filter_speckles(int16_t* img, int W, int H, int newVal, int maxSize, int maxDiff, int* scratch, cudaStream_t st,
                           int nf, size_t img_stride, size_t scratch_stride)
{
    int n = W * H;
    int* L = scratch;
    int* sz = scratch + (((size_t)n + 3) & ~(size_t)3);   // 16-byte aligned like L (the scratch holds 3n ints)
    ccl_local_kernel<<<dim3((W + CTX - 1) / CTX, (H + CTY - 1) / CTY, nf), 256, 0, st>>>(img, L, sz, W, H, newVal, maxDiff, img_stride,
                                                                                        scratch_stride);
    {
        const long long nbp = (long long)((W - 1) / CTX) * H + (long long)((H - 1) / CTY) * W;
        if (nbp > 0)
            ccl_border_kernel<<<dim3((unsigned)((nbp + 255) / 256), nf), 256, 0, st>>>(img, L, W, H, newVal, maxDiff, img_stride, scratch_stride);
    }
    const int nb4 = (n + 1023) / 1024;
    ccl_merge_counts_kernel<<<dim3(nb4, nf), 256, 0, st>>>(L, sz, n, scratch_stride);
    ccl_apply_kernel<<<dim3(nb4, nf), 256, 0, st>>>(img, L, sz, n, newVal, maxSize, img_stride, scratch_stride);
    return 4;
}

}  // namespace b200s
