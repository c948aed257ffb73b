// How fast can a GroupNorm-apply-shaped pass stream a bf16 NHWC tensor (read 16 B, 8 FMAs + SiLU, write 16 B)?  Variants of
// loads in flight per thread and CTA shape; 42 MB tensors rotated over 8 buffers (HBM-resident).  Reference: cudaMemcpyAsync D2D.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
__device__ __forceinline__ float bl(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bh(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pk(float a, float b) { __nv_bfloat162 v = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ float silu(float x) { float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x)); float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e)); return x * r; }
__device__ __forceinline__ uint4 xf(uint4 v, float sc, float sh, int act) {
    float f[8] = {bl(v.x), bh(v.x), bl(v.y), bh(v.y), bl(v.z), bh(v.z), bl(v.w), bh(v.w)};
#pragma unroll
    for (int k = 0; k < 8; ++k) { float y = fmaf(f[k], sc, sh); f[k] = act ? silu(y) : y; }
    return make_uint4(pk(f[0], f[1]), pk(f[2], f[3]), pk(f[4], f[5]), pk(f[6], f[7]));
}
// flat grid-stride, U independent loads in flight per thread
template <int U>
__global__ void __launch_bounds__(256) flat(const uint4* __restrict__ x, uint4* __restrict__ y, int64_t n, float sc, float sh, int act) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldg(x + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) y[i + u * stride] = xf(v[u], sc, sh, act);
    }
    for (; i < n; i += stride) y[i] = xf(__ldg(x + i), sc, sh, act);
}
// one contiguous chunk per CTA (like a per-sample slab), U loads in flight
template <int U>
__global__ void __launch_bounds__(256) chunked(const uint4* __restrict__ x, uint4* __restrict__ y, int64_t n, int64_t per_cta, float sc, float sh, int act) {
    const int64_t b0 = (int64_t)blockIdx.x * per_cta, b1 = b0 + per_cta < n ? b0 + per_cta : n;
    int64_t i = b0 + threadIdx.x;
    for (; i + (U - 1) * 256 < b1; i += U * 256) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = __ldg(x + i + u * 256);
#pragma unroll
        for (int u = 0; u < U; ++u) y[i + u * 256] = xf(v[u], sc, sh, act);
    }
    for (; i < b1; i += 256) y[i] = xf(__ldg(x + i), sc, sh, act);
}
int main() {
    const int64_t n = 16LL * 4096 * 320 / 8;   // 16-byte vectors of one (16, 64, 64, 320) bf16 tensor
    const int NB = 8;
    uint4 *x[NB], *y[NB];
    for (int i = 0; i < NB; ++i) { CK(cudaMalloc(&x[i], n * 16)); CK(cudaMalloc(&y[i], n * 16)); CK(cudaMemset(x[i], 0x3c, n * 16)); }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](const char* name, auto launch) {
        for (int i = 0; i < 4; ++i) launch(i % NB);
        cudaEventRecord(e0);
        for (int i = 0; i < 40; ++i) launch(i % NB);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-44s %7.2f us  %5.2f TB/s (read + write)\n", name, ms * 1000 / 40, 2.0 * n * 16 / (ms / 40 * 1e-3) / 1e12);
    };
    time("cudaMemcpyAsync D2D", [&](int b) { cudaMemcpyAsync(y[b], x[b], n * 16, cudaMemcpyDeviceToDevice); });
    for (int act = 0; act < 2; ++act) {
        printf("-- %s\n", act ? "with SiLU" : "scale/shift only");
        time("flat U=4 grid 148*8", [&](int b) { flat<4><<<148 * 8, 256>>>(x[b], y[b], n, 1.1f, 0.1f, act); });
        time("flat U=8 grid 148*8", [&](int b) { flat<8><<<148 * 8, 256>>>(x[b], y[b], n, 1.1f, 0.1f, act); });
        time("flat U=8 grid 148*4", [&](int b) { flat<8><<<148 * 4, 256>>>(x[b], y[b], n, 1.1f, 0.1f, act); });
        time("flat U=16 grid 148*4", [&](int b) { flat<16><<<148 * 4, 256>>>(x[b], y[b], n, 1.1f, 0.1f, act); });
        time("flat U=4 one vector-quad per thread", [&](int b) { flat<4><<<(unsigned)((n / 4 + 255) / 256), 256>>>(x[b], y[b], n, 1.1f, 0.1f, act); });
        time("chunked U=4 1280 CTAs", [&](int b) { chunked<4><<<1280, 256>>>(x[b], y[b], n, (n + 1279) / 1280, 1.1f, 0.1f, act); });
        time("chunked U=8 1280 CTAs", [&](int b) { chunked<8><<<1280, 256>>>(x[b], y[b], n, (n + 1279) / 1280, 1.1f, 0.1f, act); });
        time("chunked U=8 592 CTAs", [&](int b) { chunked<8><<<592, 256>>>(x[b], y[b], n, (n + 591) / 592, 1.1f, 0.1f, act); });
        time("chunked U=8 5120 CTAs", [&](int b) { chunked<8><<<5120, 256>>>(x[b], y[b], n, (n + 5119) / 5120, 1.1f, 0.1f, act); });
    }
    return 0;
}
