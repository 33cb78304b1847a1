// Microbenchmarks: dependent-issue latency of DFMA / DADD / LDS.64 / SHFL pair, and DFMA throughput per SM.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat_kernel(double* out, long long* cyc, int iters) {
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double a = 1.0 + threadIdx.x * 1e-9, b = 0.999999, c = 1e-12;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) a = fma(a, b, c);
    }
    long long t1 = clock64();
    double d = a;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) d = d + c;
    }
    long long t2 = clock64();
    int idx = threadIdx.x & 31;
    double e = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) { double v = sm[idx]; idx = (idx + (int)(v)) & 1023; e += v; }
    }
    long long t3 = clock64();
    double f = a;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) f += __shfl_xor_sync(0xffffffffu, f, 1 << (k % 5));
    }
    long long t4 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + d + e + f;
}
__global__ void thr_kernel(double* out, long long* cyc, int iters) {
    double a[8];
    for (int k = 0; k < 8; ++k) a[k] = 1.0 + threadIdx.x * 1e-9 + k;
    const double b = 0.999999, c = 1e-12;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = fma(a[k], b, c);
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    double s = 0; for (int k = 0; k < 8; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 1024 * 1024); cudaMallocManaged(&cyc, 64);
    int iters = 1000;
    lat_kernel<<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize();
    printf("latency (cycles per dependent op, 1 warp): DFMA %.1f  DADD %.1f  LDS.64+IADD chain %.1f  SHFL.64+DADD %.1f\n",
           cyc[0] / (16.0 * iters), cyc[1] / (16.0 * iters), cyc[2] / (16.0 * iters), cyc[3] / (16.0 * iters));
    for (int warps : {1, 2, 4, 8, 16, 32}) {
        thr_kernel<<<148, 32 * warps>>>(out, cyc, iters); cudaDeviceSynchronize();
        double fma_per_clk = 8.0 * iters * 32 * warps / (double)cyc[0];
        printf("throughput: %2d warps/SM x 8 independent chains: %.1f DFMA lanes/clk/SM\n", warps, fma_per_clk);
    }
    return 0;
}
