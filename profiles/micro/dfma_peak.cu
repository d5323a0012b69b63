// FP64 roofline denominator: sweep of the DFMA-chain microbenchmark over (independent chains per thread, warps per SM, CTAs per SM).
// The library's ca_measure_fp64_peak uses the best configuration found here.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dfma_peak dfma_peak.cu && ./dfma_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_peak(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) x[j] = threadIdx.x + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 256 / ILP; ++r) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += x[j];
    if (s == 12345.678) out[0] = s;
}

template <int ILP>
static void run(int sms, int threads, int ctas_per_sm, double* d) {
    const int iters = 4096;   // 4096 x 256 DFMA per thread
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k_peak<ILP><<<sms * ctas_per_sm, threads>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double flops = 2.0 * 256.0 * iters * (double)threads * sms * ctas_per_sm;
    printf("chains %2d warps/SM %2d (CTAs/SM %d x %3d threads)  %.3f ms  %.2f TFLOP/s\n", ILP, threads / 32 * ctas_per_sm, ctas_per_sm, threads, best, flops / best * 1e-9);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, max clock %d MHz: nominal %.2f TFLOP/s (SMs x 64 FMA/clk x 2 x clock)\n", p.name, p.multiProcessorCount, clk / 1000,
           p.multiProcessorCount * 128.0 * clk * 1e-9);
    double* d; cudaMalloc(&d, 64);
    const int sms = p.multiProcessorCount;
    for (int threads : {128, 256, 512, 1024}) {
        run<4>(sms, threads, 1, d); run<8>(sms, threads, 1, d); run<16>(sms, threads, 1, d);
    }
    run<8>(sms, 1024, 2, d); run<4>(sms, 1024, 2, d); run<2>(sms, 1024, 2, d); run<1>(sms, 1024, 2, d);
    return 0;
}
