// Microbenchmark (development): FP64 tensor-core mma throughput vs DFMA on B200, plus dependent-issue latencies.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ILP>
__global__ void k_dmma884(double* out, int iters) {
    double c[ILP][2];
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma884(c[i], a, b);
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_dmma16816(double* out, int iters) {
    double c[ILP][4];
    for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0;
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < 4; ++i) b[i] = 1.0 + threadIdx.x * 1e-4 + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma16816(c[i], a, b);
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_dfma(double* out, int iters) {
    double c[ILP];
    for (int i = 0; i < ILP; ++i) c[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a, b);
    double s = 0;
    for (int i = 0; i < ILP; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_lat(double* out, long long* cyc, int iters) {
    double c = threadIdx.x, a = 1.0000001, b = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) c = fma(c, a, b);
    long long t1 = clock64();
    double s = c;
    for (int it = 0; it < iters; ++it) s += __shfl_xor_sync(0xffffffffu, s, 1);
    long long t2 = clock64();
    double m[2] = {c, s};
    for (int it = 0; it < iters; ++it) dmma884(m, a, b);
    long long t3 = clock64();
    out[threadIdx.x] = s + m[0] + m[1];
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; }
}

template <class F>
float timeit(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    double* out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double));
    long long* cyc; cudaMalloc(&cyc, 64);
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        int grid = 148 * 2, threads = warps * 32 / 2 < 32 ? 32 : warps * 32 / 2;  // warps per SM = 2 blocks x threads/32
        float ms = timeit([&] { k_dmma884<8><<<grid, threads>>>(out, iters); });
        double fl = 2.0 * 256 * 8 * (double)iters * grid * (threads / 32);
        printf("dmma m8n8k4   warps/SM %2d: %.2f TFLOP/s\n", 2 * threads / 32, fl / ms / 1e9);
        ms = timeit([&] { k_dmma16816<4><<<grid, threads>>>(out, iters); });
        fl = 2.0 * 2048 * 4 * (double)iters * grid * (threads / 32);
        printf("dmma m16n8k16 warps/SM %2d: %.2f TFLOP/s\n", 2 * threads / 32, fl / ms / 1e9);
        ms = timeit([&] { k_dfma<16><<<grid, threads>>>(out, iters); });
        fl = 2.0 * 32 * 16 * (double)iters * grid * (threads / 32);
        printf("dfma          warps/SM %2d: %.2f TFLOP/s\n", 2 * threads / 32, fl / ms / 1e9);
    }
    k_lat<<<1, 32>>>(out, cyc, 4096);
    long long h[3]; cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost);
    printf("dependent latency (cycles): DFMA %.1f  SHFL+DADD(f64) %.1f  DMMA884 %.1f\n", h[0] / 4096.0, h[1] / 4096.0, h[2] / 4096.0);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
