#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
template <int NM, int NF>
__global__ void k_mix(double* out, int iters) {
    double c[8][2], f[16];
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
    for (int i = 0; i < 16; ++i) f[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NM; ++i) dmma884(c[i], a, b);
#pragma unroll
        for (int i = 0; i < NF; ++i) f[i] = fma(f[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    for (int i = 0; i < 16; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    double* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double));
    const int iters = 20000, grid = 148 * 2, threads = 256;
    float m1 = timeit([&] { k_mix<8, 0><<<grid, threads>>>(out, iters); });
    float m2 = timeit([&] { k_mix<0, 16><<<grid, threads>>>(out, iters); });
    float m3 = timeit([&] { k_mix<8, 16><<<grid, threads>>>(out, iters); });
    float m4 = timeit([&] { k_mix<8, 8><<<grid, threads>>>(out, iters); });
    double w = (double)iters * grid * (threads / 32);
    printf("dmma only   : %.2f ms  %.2f TF\n", m1, 2.0 * 256 * 8 * w / m1 / 1e9);
    printf("dfma only   : %.2f ms  %.2f TF\n", m2, 2.0 * 32 * 16 * w / m2 / 1e9);
    printf("8 dmma+16 dfma: %.2f ms  (sum of separate %.2f)  combined %.2f TF\n", m3, m1 + m2, (2.0 * 256 * 8 + 2.0 * 32 * 16) * w / m3 / 1e9);
    printf("8 dmma+ 8 dfma: %.2f ms  combined %.2f TF\n", m4, (2.0 * 256 * 8 + 2.0 * 32 * 8) * w / m4 / 1e9);
    return 0;
}
