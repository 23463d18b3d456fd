#include <cuda_runtime.h>
#include "dhts_api.h"
DHTS_EXPORT int dhts_version(void) { return 200; }

// FP64 issue-rate probe for the roofline report (bench.py `roofline_fp64`): every thread runs 8 independent
// DFMA chains, so the FP64 pipe, not dependency latency, is what limits it.  Not on the simulation path.
namespace {
__global__ void __launch_bounds__(256) fp64_probe_kernel(double* __restrict__ out, int iters) {
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * blockIdx.x;
    double x0 = 0.1, x1 = 0.2, x2 = 0.3, x3 = 0.4, x4 = 0.5, x5 = 0.6, x6 = 0.7, x7 = 0.8;
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
}  // namespace

DHTS_EXPORT long long dhts_fp64_probe(double* out, int blocks, int iters, void* stream) {
    if (!out || blocks < 1 || iters < 1) return -1;
    fp64_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, iters);
    if (cudaGetLastError() != cudaSuccess) return -1;
    return (long long)blocks * 256 * 8 * (long long)iters;     // DFMA thread-instructions enqueued
}
