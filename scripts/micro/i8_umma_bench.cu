// i8_umma_bench.cu -- measured int8 tensor-core peak of this GPU (tcgen05.mma kind::i8), the roofline denominator of
// stage 1 (gram_i8.cuh).  MEASURED_PEAKS.json has bf16 only; SURVEY 8d asks for a directly measured figure.
//
// One CTA per SM, one elected thread issues back-to-back tcgen05.mma (operands in shared memory, SWIZZLE_64B K-major,
// contents irrelevant) into four rotating TMEM accumulators and commits to an mbarrier at the end.  Two shapes: the
// production shape of the Gram kernel (M = 128, N = 64, K = 32) and the widest one (N = 256).
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o i8_umma_bench i8_umma_bench.cu && ./i8_umma_bench
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned long long desc_sw64(unsigned smem_addr) {
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr >> 4) & 0x3fff);
    d |= (unsigned long long)1 << 16;
    d |= (unsigned long long)(512 >> 4) << 32;
    d |= (unsigned long long)1 << 46;
    d |= (unsigned long long)4 << 61;
    return d;
}

template <int N>
__global__ void __launch_bounds__(128, 1) umma_i8_kernel(int iters) {
    extern __shared__ unsigned char raw[];
    unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ unsigned long long bar;
    __shared__ unsigned tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + N) * 64 / 4; i += 128) reinterpret_cast<unsigned*>(tiles)[i] = 0x01010101u * (i & 3);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes of the tiles -> async proxy (MMA)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = tmem_slot;
    if (tid == 0) {
        const unsigned idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
        const unsigned long long ad = desc_sw64(smem_u32(tiles)), bd = desc_sw64(smem_u32(tiles) + 128 * 64);
        constexpr int NACC = 512 / N < 4 ? 512 / N : 4;
        for (int it = 0; it < iters; ++it) {
            const unsigned acc = it >= NACC ? 1u : 0u;
            const unsigned d = tmem + (it % NACC) * N;
            const unsigned long long a = ad + 2 * (it & 1), b = bd + 2 * (it & 1);  // the two K = 32 halves of the 64-byte rows
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d),
                "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        unsigned done = 0;
        for (long long spin = 0; !done; ++spin) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done)
                         : "r"(smem_u32(&bar)), "r"(0u)
                         : "memory");
            if (spin > (1LL << 28)) __trap();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

template <int N>
static double run(int sms, int iters) {
    const size_t smem = (128 + N) * 64 + 1024;
    cudaFuncSetAttribute(umma_i8_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    umma_i8_kernel<N><<<sms, 128, smem>>>(iters);  // warm-up
    cudaDeviceSynchronize();
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        umma_i8_kernel<N><<<sms, 128, smem>>>(iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tops = 2.0 * 128 * N * 32 * (double)iters * sms / (ms * 1e-3) / 1e12;
        if (tops > best) best = tops;
    }
    return best;
}

int main() {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) {
        std::printf("no CUDA device\n");
        return 1;
    }
    const int sms = prop.multiProcessorCount;
    const int iters = 200000;
    std::printf("i8 umma m128n64k32: %.1f TOPS (%d SMs, %d MMAs per CTA)\n", run<64>(sms, iters), sms, iters);
    std::printf("i8 umma m128n256k32: %.1f TOPS (%d SMs, %d MMAs per CTA)\n", run<256>(sms, iters / 4), sms, iters / 4);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        std::printf("CUDA error: %s\n", cudaGetErrorString(e));
        return 1;
    }
    return 0;
}
