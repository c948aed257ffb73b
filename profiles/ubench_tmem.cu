// Micro-benchmark behind the attention softmax design: how fast can the softmax warps of one SM pull S out of
// TMEM (tcgen05.ld) and exponentiate it (MUFU.EX2), alone and together, and do the two overlap
//   (i)  inside one warp when the next chunk's tcgen05.ld is issued before the exponentials of the current chunk, and
//   (ii) between different warps of one sub-partition?
// One CTA per SM, 512 TMEM columns, W warps per sub-partition; every warp repeats "tile" rounds over its own 32 TMEM lanes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench_tmem ubench_tmem.cu      Run: ./ubench_tmem
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ float ex2v(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

#define LD32(taddr, r) asm volatile( \
    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, " \
    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), \
      "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), \
      "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), \
      "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory")
#define LD16(taddr, r) asm volatile( \
    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), \
      "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory")
#define WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")
// 16 packed registers (32 bf16) -> 16 TMEM columns
#define ST16(taddr, r) asm volatile( \
    "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" \
    :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), \
       "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory")

// 2^x on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + f, f in [-0.5, 0.5], degree-3 minimax polynomial for 2^f,
// exponent added to the bit pattern.  ~1e-4 relative error (P is rounded to bf16, 4e-3).
__device__ __forceinline__ float exp2_poly(float x) {
    x = fmaxf(x, -125.0f);
    const float xf = x + 12582912.0f;                 // 1.5 * 2^23: the low mantissa bits now hold round(x)
    const float f = x - (xf - 12582912.0f);
    float p = fmaf(f, 0.0555041086f, 0.2402265069f);
    p = fmaf(p, f, 0.6931471806f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(xf) << 23));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}

// MODE 0: tcgen05.ld only (x32, wait after each)          1: ex2 only (64 per round on register data)
//      2: ld x32 -> wait -> 32 x (fma, ex2, max) , twice per round (the current softmax inner loop, no stores)
//      3: like 2, software-pipelined: the next chunk's ld is in flight under the current chunk's exponentials
//      4: warps with (warp/4) even do MODE 0, odd do MODE 1  (do ld and MUFU of DIFFERENT warps overlap?)
//      5: like 2 plus P packed to bf16 and written back to TMEM with tcgen05.st (P-in-TMEM cost)
//      8 / 9 / 10: like 5 (ld -> exp -> tcgen05.st P) with every 4th / 3rd / 2nd exponential on the FMA pipe (exp2_poly)
//      11: like 2 with every 4th exponential on the FMA pipe, no P store
//      6: like 2 with x16 loads (4 per round)
//      7: like 3 with x16 chunks
template <int MODE, int W>
__global__ void __launch_bounds__(W * 128, 1) k(int rounds, float* out, long long* cyc) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >> 2) * 64) % 448;
    float acc = 0.f, m = 0.f;
    const float c = 0.25f;
    uint32_t a[32], b[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { a[i] = 0x3c000000u + threadIdx.x + i; b[i] = 0x3c100000u + threadIdx.x * 3 + i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        const int mode = MODE == 4 ? ((warp >> 2) & 1) : MODE;
        if (mode == 0) {
            LD32(base, a); WAIT_LD();
            LD32(base + 32, b); WAIT_LD();
            acc += __uint_as_float(a[0]) + __uint_as_float(b[31]);
        } else if (mode == 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) { acc += ex2v(fmaf(__uint_as_float(a[i]), c, -m)); }
#pragma unroll
            for (int i = 0; i < 32; ++i) { acc += ex2v(fmaf(__uint_as_float(b[i]), c, -m)); }
            m += 1e-9f;
        } else if (mode == 2 || mode == 5) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                LD32(base + h * 32, a); WAIT_LD();
                float mx = m;
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float p0 = ex2v(fmaf(__uint_as_float(a[i]), c, -m)), p1 = ex2v(fmaf(__uint_as_float(a[i + 1]), c, -m));
                    mx = fmaxf(mx, fmaxf(__uint_as_float(a[i]), __uint_as_float(a[i + 1])));
                    pk[i >> 1] = pack2(p0, p1);
                }
                if (mode == 5) { ST16(base + h * 16, pk); }
                else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc += __uint_as_float(pk[i]);
                }
                m = fminf(mx, 1e-9f);
            }
        } else if (mode >= 8 && mode <= 11) {
            constexpr int EVERY = MODE == 8 ? 4 : MODE == 9 ? 3 : MODE == 10 ? 2 : 4;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                LD32(base + h * 32, a); WAIT_LD();
                float mx = m;
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float x0 = fmaf(__uint_as_float(a[i]), c, -m), x1 = fmaf(__uint_as_float(a[i + 1]), c, -m);
                    const float p0 = (i % EVERY) == EVERY - 1 ? exp2_poly(x0) : ex2v(x0);
                    const float p1 = ((i + 1) % EVERY) == EVERY - 1 ? exp2_poly(x1) : ex2v(x1);
                    mx = fmaxf(mx, fmaxf(__uint_as_float(a[i]), __uint_as_float(a[i + 1])));
                    pk[i >> 1] = pack2(p0, p1);
                }
                if (mode != 11) { ST16(base + h * 16, pk); }
                else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc += __uint_as_float(pk[i]);
                }
                m = fminf(mx, 1e-9f);
            }
        } else if (mode == 3) {
            // prologue of the round: a in flight
            if (r == 0) { LD32(base, a); }
            WAIT_LD();
            LD32(base + 32, b);
            {
                float mx = m;
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float p0 = ex2v(fmaf(__uint_as_float(a[i]), c, -m)), p1 = ex2v(fmaf(__uint_as_float(a[i + 1]), c, -m));
                    mx = fmaxf(mx, fmaxf(__uint_as_float(a[i]), __uint_as_float(a[i + 1])));
                    acc += __uint_as_float(pack2(p0, p1));
                }
                m = fminf(mx, 1e-9f);
            }
            WAIT_LD();
            LD32(base, a);
            {
                float mx = m;
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float p0 = ex2v(fmaf(__uint_as_float(b[i]), c, -m)), p1 = ex2v(fmaf(__uint_as_float(b[i + 1]), c, -m));
                    mx = fmaxf(mx, fmaxf(__uint_as_float(b[i]), __uint_as_float(b[i + 1])));
                    acc += __uint_as_float(pack2(p0, p1));
                }
                m = fminf(mx, 1e-9f);
            }
        } else if (mode == 6) {
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                uint32_t s[16];
                LD16(base + h * 16, s); WAIT_LD();
                float mx = m;
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const float p0 = ex2v(fmaf(__uint_as_float(s[i]), c, -m)), p1 = ex2v(fmaf(__uint_as_float(s[i + 1]), c, -m));
                    mx = fmaxf(mx, fmaxf(__uint_as_float(s[i]), __uint_as_float(s[i + 1])));
                    acc += __uint_as_float(pack2(p0, p1));
                }
                m = fminf(mx, 1e-9f);
            }
        } else if (mode == 7) {
            uint32_t (&s0)[32] = a;   // use halves of a / b as 16-register chunks
            uint32_t s1[16];
            if (r == 0) { LD16(base, s0); }
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                WAIT_LD();
                if (h & 1) { LD16(base + ((h + 1) & 3) * 16, s0); } else { LD16(base + ((h + 1) & 3) * 16, s1); }
                float mx = m;
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const uint32_t u0 = (h & 1) ? s1[i] : s0[i], u1 = (h & 1) ? s1[i + 1] : s0[i + 1];
                    const float p0 = ex2v(fmaf(__uint_as_float(u0), c, -m)), p1 = ex2v(fmaf(__uint_as_float(u1), c, -m));
                    mx = fmaxf(mx, fmaxf(__uint_as_float(u0), __uint_as_float(u1)));
                    acc += __uint_as_float(pack2(p0, p1));
                }
                m = fminf(mx, 1e-9f);
            }
        }
    }
    WAIT_LD();
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    const long long t1 = clock64();
    if (acc == 123.456f) out[threadIdx.x] = acc + m;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
    }
}

template <int MODE, int W>
int run(const char* name, int rounds, float* out, long long* cyc) {
    const int threads = W * 4 * 32, warps_per_sp = W;
    k<MODE, W><<<148, threads>>>(10, out, cyc);
    CK(cudaDeviceSynchronize());
    k<MODE, W><<<148, threads>>>(rounds, out, cyc);
    CK(cudaDeviceSynchronize());
    long long c;
    CK(cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost));
    // one "tile" = 128 rows x 64 columns = one round of 4 warps (one per sub-partition)
    const double tiles = (double)rounds * warps_per_sp * (MODE == 4 ? 0.5 : 1.0);
    printf("%-34s warps/SMSP %d: %8.1f cycles per 128x64 tile per SM  (%lld cycles, %d rounds)\n", name, warps_per_sp, c / tiles, c, rounds);
    return 0;
}

int main() {
    float* out; long long* cyc;
    CK(cudaMalloc(&out, 4096 * sizeof(float)));
    CK(cudaMalloc(&cyc, sizeof(long long)));
    const int R = 2000;
#define ALL(W) \
    if (run<0, W>("ld only (x32 + wait)", R, out, cyc)) return 1; \
    if (run<1, W>("ex2 only", R, out, cyc)) return 1; \
    if (run<2, W>("ld -> wait -> exp (current loop)", R, out, cyc)) return 1; \
    if (run<3, W>("ld pipelined under exp (x32)", R, out, cyc)) return 1; \
    if (W >= 2 && run<4, W>("half the warps ld, half ex2", R, out, cyc)) return 1; \
    if (run<5, W>("ld -> exp -> tcgen05.st P", R, out, cyc)) return 1; \
    if (run<6, W>("ld x16 -> wait -> exp", R, out, cyc)) return 1; \
    if (run<7, W>("ld x16 pipelined under exp", R, out, cyc)) return 1; \
    if (run<8, W>("ld -> exp (1/4 poly) -> tcgen05.st P", R, out, cyc)) return 1; \
    if (run<9, W>("ld -> exp (1/3 poly) -> tcgen05.st P", R, out, cyc)) return 1; \
    if (run<10, W>("ld -> exp (1/2 poly) -> tcgen05.st P", R, out, cyc)) return 1; \
    if (run<11, W>("ld -> exp (1/4 poly), no P store", R, out, cyc)) return 1;
    ALL(1) ALL(2) ALL(4) ALL(8)
    return 0;
}
