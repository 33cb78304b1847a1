// Do a "fat" kernel (1024 threads, 227 KB smem, few long CTAs) and a "thin" kernel (96 threads, many CTAs)
// overlap when launched on two streams?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(1024) fat(long long cycles, int nreal, double* out) {
    extern __shared__ double sm[];
    if (blockIdx.x >= nreal) return;
    long long t0 = clock64();
    double a = threadIdx.x;
    while (clock64() - t0 < cycles) a = a * 1.0000001 + 1e-9;
    sm[threadIdx.x] = a;
    if (threadIdx.x == 0) out[blockIdx.x] = sm[0];
}
__global__ void thin(long long cycles, double* out) {
    long long t0 = clock64();
    double a = threadIdx.x;
    while (clock64() - t0 < cycles) a = a * 1.0000001 + 1e-9;
    if (threadIdx.x == 0) out[blockIdx.x] = a;
}
int main() {
    double* out; cudaMalloc(&out, 1 << 20);
    cudaStream_t s1, s2; cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    cudaFuncSetAttribute(fat, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaEvent_t e0, e1, ef, ej; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreateWithFlags(&ef, cudaEventDisableTiming); cudaEventCreateWithFlags(&ej, cudaEventDisableTiming);
    const long long fat_cyc = 6000000, thin_cyc = 250000;  // ~3 ms per fat CTA; thin: 3348 CTAs x 21/SM -> ~ 1 wave... 
    for (int mode = 0; mode < 3; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaDeviceSynchronize();
            cudaEventRecord(e0, s1);
            if (mode == 0) {  // serial on one stream
                fat<<<418, 1024, 227 * 1024, s1>>>(fat_cyc, 28, out);
                thin<<<3348 * 8, 96, 8192, s1>>>(thin_cyc, out);
            } else if (mode == 1) {  // fork/join: fat first
                cudaEventRecord(ef, s1); cudaStreamWaitEvent(s2, ef, 0);
                fat<<<418, 1024, 227 * 1024, s2>>>(fat_cyc, 28, out);
                cudaEventRecord(ej, s2);
                thin<<<3348 * 8, 96, 8192, s1>>>(thin_cyc, out);
                cudaStreamWaitEvent(s1, ej, 0);
            } else {  // fork/join: thin first
                cudaEventRecord(ef, s1); cudaStreamWaitEvent(s2, ef, 0);
                thin<<<3348 * 8, 96, 8192, s1>>>(thin_cyc, out);
                fat<<<418, 1024, 227 * 1024, s2>>>(fat_cyc, 28, out);
                cudaEventRecord(ej, s2);
                cudaStreamWaitEvent(s1, ej, 0);
            }
            cudaEventRecord(e1, s1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("mode %d rep %d: %.3f ms\n", mode, rep, ms);
        }
    }
    return 0;
}
