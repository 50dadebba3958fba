// Prefilters of cv::StereoBM (x-Sobel and normalised response) and the colour conversions on the path.
// Reference call sites: prefilter runs inside StereoBM::compute (src/GPUStereoProcessor.cpp:283,319);
// colour conversion: convertColor (src/GPUStereoProcessor.cpp:119-172).  Algorithm: SURVEY.md A.2.1.
#include "kernels.h"

namespace b200s {

// x-Sobel with OpenCV's border rules: rows mirrored (reflect-101), columns 0/W-1 = cap, odd-height last row = cap.
__global__ void __launch_bounds__(256) xsobel_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                     size_t dpitch, int W, int H, int cap)
{
    int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    int out = cap;
    bool last_odd = (H & 1) && (y == H - 1);
    if (x > 0 && x < W - 1 && !last_odd && H > 1) {
        int yu = y == 0 ? 1 : y - 1, yd = y == H - 1 ? H - 2 : y + 1;
        const uint8_t* r0 = src + (size_t)yu * W + x;
        const uint8_t* r1 = src + (size_t)y * W + x;
        const uint8_t* r2 = src + (size_t)yd * W + x;
        int v = ((int)__ldg(r0 + 1) - (int)__ldg(r0 - 1)) + 2 * ((int)__ldg(r1 + 1) - (int)__ldg(r1 - 1)) +
                ((int)__ldg(r2 + 1) - (int)__ldg(r2 - 1));
        out = min(max(v, -cap), cap) + cap;
    }
    dst[(size_t)y * dpitch + x] = (uint8_t)out;
}

// normalised response: vertical box sums (replicate) into scratch, then horizontal box + centre term
__global__ void __launch_bounds__(256) norm_vsum_kernel(const uint8_t* __restrict__ src, int* __restrict__ vs,
                                                        int W, int H, int p2)
{
    int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    int s = 0;
    for (int dy = -p2; dy <= p2; ++dy) s += __ldg(src + (size_t)min(max(y + dy, 0), H - 1) * W + x);
    vs[(size_t)y * W + x] = s;
}

__global__ void __launch_bounds__(256) norm_final_kernel(const uint8_t* __restrict__ src, const int* __restrict__ vs,
                                                         uint8_t* __restrict__ dst, size_t dpitch, int W, int H, int p2,
                                                         int scale_g, int scale_s, int cap)
{
    int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    long long sum = 0;
    for (int dx = -p2; dx <= p2; ++dx) sum += __ldg(vs + (size_t)y * W + min(max(x + dx, 0), W - 1));
    const uint8_t* cur = src + (size_t)y * W;
    int c = 4 * cur[x] + cur[max(x - 1, 0)] + cur[min(x + 1, W - 1)] + src[(size_t)max(y - 1, 0) * W + x] +
            src[(size_t)min(y + 1, H - 1) * W + x];
    long long val = ((long long)c * scale_g - sum * scale_s) >> 10;
    int v = (int)max(-(long long)cap, min((long long)cap, val)) + cap;
    dst[(size_t)y * dpitch + x] = (uint8_t)v;
}

__global__ void bgr_to_gray_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int n, int rgb)
{
    // cv::cvtColor BGR2GRAY fixed point (OpenCV 4.x, 15 bit): (B*3735 + G*19235 + R*9798 + 16384) >> 15
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c0 = src[3 * (size_t)i], c1 = src[3 * (size_t)i + 1], c2 = src[3 * (size_t)i + 2];
    int b = rgb ? c2 : c0, r = rgb ? c0 : c2;
    dst[i] = (uint8_t)((b * 3735 + c1 * 19235 + r * 9798 + 16384) >> 15);
}

__global__ void gray_to_bgr_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t v = src[i];
    dst[3 * (size_t)i] = v; dst[3 * (size_t)i + 1] = v; dst[3 * (size_t)i + 2] = v;
}

__global__ void swap_rb_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t a = src[3 * (size_t)i], b = src[3 * (size_t)i + 1], c = src[3 * (size_t)i + 2];
    dst[3 * (size_t)i] = c; dst[3 * (size_t)i + 1] = b; dst[3 * (size_t)i + 2] = a;
}

// printStats: per-channel min / max / sum of a plane; one partial (min, max, sum) triple per block and channel
template <typename T>
__global__ void __launch_bounds__(256) mat_stats_kernel(const T* __restrict__ src, size_t npix, int ch, double* __restrict__ part)
{
    __shared__ double smn[256], smx[256], ssum[256];
    for (int c = 0; c < ch; ++c) {
        double mn = 1e300, mx = -1e300, sum = 0;
        for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < npix; i += (size_t)gridDim.x * 256) {
            const double v = (double)src[i * ch + c];
            mn = fmin(mn, v); mx = fmax(mx, v); sum += v;
        }
        smn[threadIdx.x] = mn; smx[threadIdx.x] = mx; ssum[threadIdx.x] = sum;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (threadIdx.x < s) {
                smn[threadIdx.x] = fmin(smn[threadIdx.x], smn[threadIdx.x + s]);
                smx[threadIdx.x] = fmax(smx[threadIdx.x], smx[threadIdx.x + s]);
                ssum[threadIdx.x] += ssum[threadIdx.x + s];
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            double* o = part + ((size_t)blockIdx.x * ch + c) * 3;
            o[0] = smn[0]; o[1] = smx[0]; o[2] = ssum[0];
        }
        __syncthreads();
    }
}

int launch_mat_stats(const void* src, int elem_kind, size_t npix, int ch, double* partials, int nblocks, cudaStream_t st)
{
    switch (elem_kind) {
    case 0: mat_stats_kernel<uint8_t><<<nblocks, 256, 0, st>>>((const uint8_t*)src, npix, ch, partials); break;
    case 1: mat_stats_kernel<int16_t><<<nblocks, 256, 0, st>>>((const int16_t*)src, npix, ch, partials); break;
    default: mat_stats_kernel<float><<<nblocks, 256, 0, st>>>((const float*)src, npix, ch, partials); break;
    }
    return 1;
}

static inline dim3 grid2d(int W, int H) { return dim3((W + 31) / 32, (H + 7) / 8); }

int launch_prefilter_xsobel(const uint8_t* src, uint8_t* dst, size_t dst_pitch, int W, int H, int cap, cudaStream_t st)
{
    xsobel_kernel<<<grid2d(W, H), 256, 0, st>>>(src, dst, dst_pitch, W, H, cap);
    return 1;
}

int launch_prefilter_norm(const uint8_t* src, uint8_t* dst, size_t dst_pitch, int W, int H, int ps, int cap, int* scratch, cudaStream_t st)
{
    int p2 = ps / 2;
    int scale_g = ps * ps / 8, scale_s = (1024 + scale_g) / (scale_g * 2);
    scale_g *= scale_s;
    norm_vsum_kernel<<<grid2d(W, H), 256, 0, st>>>(src, scratch, W, H, p2);
    norm_final_kernel<<<grid2d(W, H), 256, 0, st>>>(src, scratch, dst, dst_pitch, W, H, p2, scale_g, scale_s, cap);
    return 2;
}

int launch_bgr_to_gray(const uint8_t* src, uint8_t* dst, int n, int rgb_order, cudaStream_t st)
{
    bgr_to_gray_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, dst, n, rgb_order);
    return 1;
}
int launch_gray_to_bgr(const uint8_t* src, uint8_t* dst, int n, cudaStream_t st)
{
    gray_to_bgr_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, dst, n);
    return 1;
}
int launch_swap_rb(const uint8_t* src, uint8_t* dst, int n, cudaStream_t st)
{
    swap_rb_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, dst, n);
    return 1;
}

}  // namespace b200s
