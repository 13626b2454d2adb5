// FP64 pipe microbenchmark: the measured DFMA issue rate of this GPU, the denominator of the FP64-pipe utilisation that
// bench.py prints beside the HBM roofline fraction (BASELINE.md §3 / SURVEY.md §7.2: measure the FP64 peak before quoting
// a utilisation).  Independent DFMA chains, no memory traffic; 64 resident warps per SM.
#include "../../include/hidenn_b200.h"
#include "common.cuh"

namespace hidenn {

constexpr int kPeakIters = 2048, kPeakChains = 8;

__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, double a, double b) {
    double x[kPeakChains];
#pragma unroll
    for (int i = 0; i < kPeakChains; ++i) x[i] = a + i + threadIdx.x;
#pragma unroll 1
    for (int it = 0; it < kPeakIters; ++it) {
#pragma unroll
        for (int i = 0; i < kPeakChains; ++i) x[i] = fma(x[i], b, a);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kPeakChains; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;      // never true: keeps the chains alive
}

}  // namespace hidenn

using namespace hidenn;

extern "C" int hidenn_fp64_peak(double* dfma_per_s, void* stream_v) {
    HIDENN_REQUIRE(dfma_per_s != nullptr, "fp64_peak: NULL");
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
    int dev = 0;
    HIDENN_CUDA_OK(cudaGetDevice(&dev));
    const int grid = sm_count(dev) * 8;
    double* d_out = nullptr;
    HIDENN_CUDA_OK(cudaMalloc(&d_out, sizeof(double)));
    cudaEvent_t e0, e1;
    HIDENN_CUDA_OK(cudaEventCreate(&e0));
    HIDENN_CUDA_OK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {      // first pass warms up
        cudaEventRecord(e0, stream);
        fp64_peak_kernel<<<grid, 256, 0, stream>>>(d_out, 1.0000001, 0.9999999);
        cudaEventRecord(e1, stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    HIDENN_CUDA_OK(cudaGetLastError());
    *dfma_per_s = (double)grid * 256.0 * kPeakIters * kPeakChains / ((double)best * 1e-3);
    return 0;
}
