// Micro-benchmark behind the attention kernel's MMA-warp design: what does ONE thread pay per key tile for
// issuing tcgen05.mma / tcgen05.commit and polling mbarriers?  (Round 1 measured ~935 cycles per 128x64 tile for the single
// MMA-issuing thread of attn_kernel; the softmax itself needs 512 MUFU cycles per tile, profiles/ubench_tmem.cu.)
// One CTA per SM; one elected thread runs the loop on garbage operands (timing only); nobody waits on the committed barriers.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench_mma_issue ubench_mma_issue.cu
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
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
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

// MODE 0: 3 S-MMAs (128x64x16) + 4 PV-MMAs (128x48x16) per iteration, no commit
//      1: + 2 commits      2: + 4 commits      3: + 4 commits + 3 polls of completed barriers + 3 tcgen05 fences  (the round-1 loop)
//      4: 2 commits + 2 polls (merged barriers)
//      5: 128-key tile: 3 S-MMAs (128x128x16) + 8 PV-MMAs, 2 commits + 2 polls           (cycles per 128 keys!)
//      6: polls only (3)    7: commits only (4)
//      8: like 3 but the S part and the P V part are issued by two different warps concurrently (cycles = max of the two)
template <int MODE>
__global__ void __launch_bounds__(128, 1) k(int iters, long long* cyc) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[8];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    const uint32_t q = smem_u32(smem), kk = q + 16384, v = q + 32768, p = q + 49152;
    constexpr uint32_t IS = idesc(128, 64, false), IO = idesc(128, 48, true), IS128 = idesc(128, 128, false);
    long long t0 = 0, t1 = 0;
    const bool two = MODE == 8;
    if ((warp == 1 || (two && warp == 2)) && elect_one()) {
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const bool do_s = !two || warp == 1, do_pv = !two || warp == 2;
            if (MODE == 6) {
                for (int i = 0; i < 3; ++i) { while (!mbar_try_wait(&bars[4 + i], 1)) {} }
                continue;
            }
            if (MODE == 7) { for (int i = 0; i < 4; ++i) commit(&bars[i]); continue; }
            if (do_s) {
                if (MODE == 3 || MODE == 8) { while (!mbar_try_wait(&bars[4], 1)) {} asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
                if (MODE == 3 || MODE == 4 || MODE == 5 || MODE == 8) { while (!mbar_try_wait(&bars[5], 1)) {} asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
#pragma unroll
                for (int ks = 0; ks < 3; ++ks) umma(tm + (it & 1) * 128, desc_k(q + ks * 32), desc_k(kk + ks * 32), MODE == 5 ? IS128 : IS, ks != 0);
                if (MODE == 2 || MODE == 3 || MODE == 8) commit(&bars[0]);
                if (MODE >= 1 && MODE <= 5 || MODE == 8) commit(&bars[1]);
            }
            if (do_pv) {
                if (MODE == 3 || MODE == 4 || MODE == 5 || MODE == 8) { while (!mbar_try_wait(&bars[6], 1)) {} asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
#pragma unroll
                for (int ks = 0; ks < (MODE == 5 ? 8 : 4); ++ks) umma(tm + 256 + (it & 1) * 48, desc_k(p + (ks & 3) * 32 + (ks >> 2) * 16384), desc_mn(v + ks * 2048, 8192), IO, 1);
                if (MODE == 2 || MODE == 3 || MODE == 8) commit(&bars[2]);
                if (MODE >= 1 && MODE <= 5 || MODE == 8) commit(&bars[3]);
            }
        }
        // drain: wait until everything issued has completed (commit + wait on a fresh barrier phase)
        t1 = clock64();
        if (cyc && blockIdx.x == 0) cyc[warp - 1] = t1 - t0;
    }
    __syncthreads();
    // let the tensor pipe finish before freeing TMEM
    if (warp == 1 && elect_one()) { commit(&bars[7]); while (!mbar_try_wait(&bars[7], 0)) {} }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
    }
}

template <int MODE>
int run(const char* name, long long* cyc) {
    const int iters = 4000;
    CK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 << 10));
    k<MODE><<<148, 128, 96 << 10>>>(100, nullptr);
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(cyc, 0, 16));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<148, 128, 96 << 10>>>(iters, cyc);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c[2];
    CK(cudaMemcpy(c, cyc, 16, cudaMemcpyDeviceToHost));
    printf("%-64s issue loop %7.1f cycles/iter (2nd warp %7.1f), kernel %7.1f ns/iter\n", name, (double)c[0] / iters, (double)c[1] / iters, ms * 1e6 / iters);
    return 0;
}

int main() {
    long long* cyc;
    CK(cudaMalloc(&cyc, 16));
    if (run<0>("7 MMAs (3x 128x64x16 + 4x 128x48x16), no commit", cyc)) return 1;
    if (run<1>("7 MMAs + 2 commits", cyc)) return 1;
    if (run<2>("7 MMAs + 4 commits", cyc)) return 1;
    if (run<3>("7 MMAs + 4 commits + 3 polls + fences (round-1 loop)", cyc)) return 1;
    if (run<4>("7 MMAs + 2 commits + 2 polls (merged barriers)", cyc)) return 1;
    if (run<5>("128-key tile: 3 + 8 MMAs + 2 commits + 2 polls (per 128 keys)", cyc)) return 1;
    if (run<6>("3 polls of completed barriers only", cyc)) return 1;
    if (run<7>("4 commits only", cyc)) return 1;
    if (run<8>("round-1 loop split over two warps (S | P V)", cyc)) return 1;
    return 0;
}
