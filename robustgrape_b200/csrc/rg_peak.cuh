// rg_peak.cuh -- FP64 peak microbenchmarks (roofline denominators for the FP64-bound kernels).
// MEASURED_PEAKS.json holds only HBM and bf16 figures; the "% of FP64 peak" denominator is measured
// here: a dependent-chain-free DFMA loop and an mma.sync m8n8k4 f64 (DMMA) loop.
#pragma once
#include <cuda_runtime.h>

static __global__ void __launch_bounds__(256) k_peak_dfma(double* out, int iters, double a, double b) {
    double r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = fma(r[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += r[i];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static __global__ void __launch_bounds__(256) k_peak_dmma(double* out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = threadIdx.x * 1e-3; c[i][1] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
