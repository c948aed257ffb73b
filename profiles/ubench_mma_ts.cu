// Micro-benchmark: what does one tcgen05.mma (M = 128, K = 16, bf16) cost as a function of N and of where the A operand lives
// (SS form: A from shared memory, TS form: A from tensor memory)?  The attention kernel's P.V product is a TS-form MMA with
// N = 48 / 80; its S product an SS-form MMA with N = 64.  One CTA per SM, one thread issues 16 MMAs per
// iteration back to back on garbage operands; time = kernel duration (the tensor pipe drains in order).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench_mma_ts ubench_mma_ts.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t db, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(db), "r"(idesc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t desc_k(uint32_t a) {
    return static_cast<uint64_t>((a & 0x3FFFF) >> 4) | (1ull << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t a, uint32_t lbo) {
    return static_cast<uint64_t>((a & 0x3FFFF) >> 4) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t idesc(int M, int N, bool bmn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((bmn ? 1u : 0u) << 16) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// TS: A from TMEM.  BMN: B operand MN-major (the V tile of attention) or K-major (the K tile).  ALT: alternate between two accumulators.
template <int N, bool TS, bool BMN>
__global__ void __launch_bounds__(128, 1) k(int iters) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    const uint32_t a = smem_u32(smem), b = a + 32768;
    constexpr uint32_t ID = idesc(128, N, BMN);
    if (warp == 1 && elect_one()) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                // D: columns [0, N); the TS form's A: 8 columns per K step behind the accumulator
                const uint64_t db = BMN ? desc_mn(b + (i & 3) * 2048, 8192) : desc_k(b + (i & 3) * 32);
                if (TS) umma_ts(tm, tm + 256 + (i & 7) * 8, db, ID);
                else umma_ss(tm, desc_k(a + (i & 3) * 32), db, ID);
            }
        }
        commit(&bar);
        while (!mbar_try_wait(&bar, 0)) {}
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
    }
}

template <int N, bool TS, bool BMN>
int run(double ghz) {
    const int iters = 2000;
    CK(cudaFuncSetAttribute(k<N, TS, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 << 10));
    k<N, TS, BMN><<<148, 128, 128 << 10>>>(100);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<N, TS, BMN><<<148, 128, 128 << 10>>>(iters);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ns = ms * 1e6 / (iters * 16.0);
    printf("M=128 N=%3d K=16  A from %s, B %s-major: %6.1f ns = %6.1f cycles at %.2f GHz per MMA   (math alone: %5.1f cycles at 8192 flop/clk/SM)\n", N, TS ? "TMEM" : "smem",
           BMN ? "MN" : "K ", ns, ns * ghz, ghz, 2.0 * 128 * N * 16 / 8192.0);
    return 0;
}

int main() {
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;   // nominal maximum; the measured kernels run near it when only the tensor pipe is busy
    if (run<48, false, true>(ghz) || run<64, false, true>(ghz) || run<128, false, true>(ghz) || run<256, false, true>(ghz)) return 1;
    if (run<48, true, true>(ghz) || run<64, true, true>(ghz) || run<80, true, true>(ghz) || run<128, true, true>(ghz) || run<256, true, true>(ghz)) return 1;
    if (run<64, false, false>(ghz) || run<128, false, false>(ghz) || run<256, false, false>(ghz)) return 1;
    if (run<64, true, false>(ghz) || run<128, true, false>(ghz) || run<256, true, false>(ghz)) return 1;
    return 0;
}
