// Integer-ALU issue-rate micro-benchmark: gives the measured INT roofline denominator that BASELINE.json asks
// for (MEASURED_PEAKS.json has no integer figure).  Each variant runs ITER x 8 independent dependent-chains of one
// SASS instruction per thread on every SM; lane-ops/s = threads * ITER * 8 / time.
#include "kernels.h"

namespace b200s {

constexpr int IP_ITERS = 4096;

template <int WHICH>
__global__ void __launch_bounds__(256) int_peak_kernel(uint32_t* out, uint32_t a, uint32_t b)
{
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 2654435761u + i * 40503u + a;
    uint32_t y = b + threadIdx.x, z = a ^ 0x5bd1e995u;
#pragma unroll 1
    for (int it = 0; it < IP_ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (WHICH == 0) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(x[i]) : "r"(y), "r"(z));
            if (WHICH == 1) asm volatile("vabsdiff4.u32.u32.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y), "r"(z));
            if (WHICH == 2) x[i] = __vadd2(x[i], y);
            if (WHICH == 3) x[i] = __vminu2(x[i] ^ z, y);
            if (WHICH == 4) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y), "r"(z));
            if (WHICH == 5) asm volatile("prmt.b32 %0, %0, %1, 0x4341;" : "+r"(x[i]) : "r"(y));
            if (WHICH == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y), "r"(z));
            if (WHICH == 8) x[i] = __shfl_down_sync(0xffffffffu, x[i], 1) + y;
            if (WHICH == 7) {
                if (i & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y), "r"(z));
                else asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(x[i]) : "r"(y), "r"(z));
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= x[i];
    if (s == 0x12345678u) out[threadIdx.x] = s;   // practically never; keeps the chains alive
}

template <int WHICH>
static float time_variant(uint32_t* dbuf, int blocks, cudaStream_t st)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    int_peak_kernel<WHICH><<<blocks, 256, 0, st>>>(dbuf, 1u, 2u);   // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0, st);
        int_peak_kernel<WHICH><<<blocks, 256, 0, st>>>(dbuf, 1u + rep, 2u);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

int run_int_peak(int which, double* lane_ops_per_s, double* sm_mhz, cudaStream_t st)
{
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    uint32_t* dbuf = nullptr;
    if (cudaMalloc(&dbuf, 256 * sizeof(uint32_t)) != cudaSuccess) return -1;
    int blocks = sms * 8;   // 2048 threads per SM
    float ms;
    switch (which) {
        case 0: ms = time_variant<0>(dbuf, blocks, st); break;
        case 1: ms = time_variant<1>(dbuf, blocks, st); break;
        case 2: ms = time_variant<2>(dbuf, blocks, st); break;
        case 3: ms = time_variant<3>(dbuf, blocks, st); break;
        case 4: ms = time_variant<4>(dbuf, blocks, st); break;
        case 5: ms = time_variant<5>(dbuf, blocks, st); break;
        case 6: ms = time_variant<6>(dbuf, blocks, st); break;
        case 7: ms = time_variant<7>(dbuf, blocks, st); break;
        case 8: ms = time_variant<8>(dbuf, blocks, st); break;
        default: cudaFree(dbuf); return -1;
    }
    cudaFree(dbuf);
    if (cudaGetLastError() != cudaSuccess) return -1;
    double ops = (double)blocks * 256.0 * IP_ITERS * 8.0;
    if (which == 3) ops *= 1.0;   // the xor feeding VIMNMX is a second instruction; reported rate is for the pair
    *lane_ops_per_s = ops / (ms * 1e-3);
    if (sm_mhz) *sm_mhz = khz / 1000.0;
    return 0;
}

}  // namespace b200s
