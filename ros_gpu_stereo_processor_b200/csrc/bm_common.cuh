// Shared device helpers of the block-matching kernels (bm_sad.cu, bm_ws.cu).
#pragma once
#include <stdint.h>

namespace b200s {

// A.2.5 sub-pixel fit and fixed-point packing of cv::StereoBM
__device__ __forceinline__ int16_t subpixel_disp(int minsad, int mind, int p, int n, int nd, int minD)
{
    int d = p + n - 2 * minsad + abs(p - n);
    int v = ((nd - mind - 1 + minD) * 256 + (d != 0 ? (p - n) * 256 / d : 0) + 15) >> 4;
    return (int16_t)v;
}

// position of disparity index k inside a group of four u16 lanes (V-phase lane order is k, k+2, k+1, k+3)
__device__ __forceinline__ int kpos(int k) { return (k & ~3) | ((k & 1) << 1) | ((k >> 1) & 1); }

}  // namespace b200s
