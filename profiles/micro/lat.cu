// Micro-benchmarks behind the K3/K4 design notes: dependent-issue latency and throughput of the FP64 pipe, shared-memory
// round trips through __syncwarp, shuffles and the MUFU seeds, one warp per SM (latency) or many (throughput).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lat lat.cu && ./lat
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma_chain(double* out, long long* cyc, int n, double a, double b) {
    double x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = fma(x, a, b);
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (x == 1234.5) out[0] = x;
}
template <int ILP>
__global__ void k_dfma_ilp(double* out, long long* cyc, int n, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) x[j] = threadIdx.x + j;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], a, b);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += x[j];
    if (s == 1234.5) out[0] = s;
}
__global__ void k_smem_roundtrip(double* out, long long* cyc, int n) {
    __shared__ double sm[64];
    double x = threadIdx.x;
    const int lane = threadIdx.x & 31;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        sm[lane] = x;
        __syncwarp();
        x = sm[lane ^ 1] + 1.0;
        __syncwarp();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (x == 1234.5) out[0] = x;
}
__global__ void k_shfl(double* out, long long* cyc, int n) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (x == 1234.5) out[0] = x;
}
__global__ void k_rcp(double* out, long long* cyc, int n) {
    double x = threadIdx.x + 2.0;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + 1.5; }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (x == 1234.5) out[0] = x;
}
__global__ void k_dadd_chain(double* out, long long* cyc, int n, double b) {
    double x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = x + b;
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (x == 1234.5) out[0] = x;
}
__global__ void k_ffma_chain(float* out, long long* cyc, int n, float a, float b) {
    float x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) x = fmaf(x, a, b);
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (x == 1234.5f) out[0] = x;
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 64); cudaMalloc(&cyc, 8 * 1024);
    long long h[1024];
    const int n = 4096;
    auto report = [&](const char* name, double per) {
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 8 * 148, cudaMemcpyDeviceToHost);
        printf("%-44s %8.2f cycles per op (block 0: %lld cycles / %d)\n", name, (double)h[0] / n / per, h[0], n);
    };
    for (int warps = 1; warps <= 16; warps *= 2) {
        char nm[96];
        k_dfma_chain<<<148, 32 * warps>>>(out, cyc, n, 1.0000001, 1e-9); snprintf(nm, 96, "DFMA dependent chain, %d warp(s)/SM", warps); report(nm, 1);
    }
    k_dadd_chain<<<148, 32>>>(out, cyc, n, 1e-9); report("DADD dependent chain, 1 warp/SM", 1);
    k_ffma_chain<<<148, 32>>>((float*)out, cyc, n, 1.0000001f, 1e-9f); report("FFMA dependent chain, 1 warp/SM", 1);
    k_dfma_ilp<2><<<148, 32>>>(out, cyc, n, 1.0000001, 1e-9); report("DFMA ILP 2, 1 warp/SM (per DFMA)", 2);
    k_dfma_ilp<4><<<148, 32>>>(out, cyc, n, 1.0000001, 1e-9); report("DFMA ILP 4, 1 warp/SM (per DFMA)", 4);
    k_dfma_ilp<8><<<148, 32>>>(out, cyc, n, 1.0000001, 1e-9); report("DFMA ILP 8, 1 warp/SM (per DFMA)", 8);
    k_dfma_ilp<16><<<148, 32>>>(out, cyc, n, 1.0000001, 1e-9); report("DFMA ILP 16, 1 warp/SM (per DFMA)", 16);
    k_dfma_ilp<8><<<148, 256>>>(out, cyc, n, 1.0000001, 1e-9); report("DFMA ILP 8, 8 warps/SM (per DFMA per warp)", 8);
    k_dfma_ilp<8><<<148, 512>>>(out, cyc, n, 1.0000001, 1e-9); report("DFMA ILP 8, 16 warps/SM (per DFMA per warp)", 8);
    k_dfma_ilp<16><<<148, 1024>>>(out, cyc, n, 1.0000001, 1e-9); report("DFMA ILP 16, 32 warps/SM (per DFMA per warp)", 16);
    k_smem_roundtrip<<<148, 32>>>(out, cyc, n); report("STS -> syncwarp -> LDS -> DADD -> syncwarp", 1);
    k_shfl<<<148, 32>>>(out, cyc, n); report("SHFL.xor (fp64 = 2 SHFL) + DADD chain", 1);
    k_rcp<<<148, 32>>>(out, cyc, n); report("MUFU.RCP64H + DADD chain", 1);
    return 0;
}
