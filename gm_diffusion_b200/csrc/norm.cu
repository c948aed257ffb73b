// Memory-bound companions of the tensor-core kernels: GroupNorm(+SiLU) over NHWC (two-source aware, so the
// up-block skip concat is materialised only once, already normalised), LayerNorm over token rows, row softmax,
// sinusoidal timestep embedding, SiLU.  16-byte vectorised; fp32 statistics; inputs bf16 or fp32 (the fp32
// variants read the tensors that never feed an MMA directly — conv1 outputs and the transformer token stream —
// so those are not rounded to bf16 on their way into a normalisation).
// Reductions are performed in a FIXED order (no atomics): results are bit-reproducible run to run and
// independent of the batch size, which is what makes image sharding across GPUs exact.
#include "common.cuh"
#include "../../include/gmd_b200.h"

namespace gmd {
void count_launch(int n);
namespace {

__device__ __forceinline__ void unpack8(uint4 v, float (&f)[8]) {
    f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
    f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
// 8 consecutive channels starting at element offset `off`
template <typename T> __device__ __forceinline__ void load8(const T* p, int64_t off, float (&f)[8]);
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, int64_t off, float (&f)[8]) {
    unpack8(__ldg(reinterpret_cast<const uint4*>(p + off)), f);
}
template <> __device__ __forceinline__ void load8<float>(const float* p, int64_t off, float (&f)[8]) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p + off)), b = __ldg(reinterpret_cast<const float4*>(p + off) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// ---------------------------------------------------------------------------------------------
// GroupNorm.  Thread t owns channel vector cv = t % (C/8) (8 consecutive channels) and pixel lane t / (C/8).
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void gn_load(const T* x0, int C0, const __nv_bfloat16* x1, int C1, int64_t px, int c, float (&f)[8]) {
    if (c < C0) load8<T>(x0, px * C0 + c, f);
    else load8<__nv_bfloat16>(x1, px * C1 + (c - C0), f);
}

template <typename T>
__global__ void gn_stats_kernel(const T* __restrict__ x0, int C0, const __nv_bfloat16* __restrict__ x1, int C1,
                                float* __restrict__ stats, float* __restrict__ finals, int* __restrict__ counters, int HW, int groups,
                                int px_per_cta, float eps) {
    extern __shared__ float s_part[];  // [blockDim][16]: per-thread per-channel sum / sum of squares
    const int C = C0 + C1, nvec = C / 8, cg = C / groups;
    const int n = blockIdx.y;
    const int cv = threadIdx.x % nvec, pl = threadIdx.x / nvec, P = blockDim.x / nvec;
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (pl < P) {
        const int p0 = blockIdx.x * px_per_cta;
        const int p1 = min(p0 + px_per_cta, HW);
        // 4 pixels per iteration: four independent 16-byte loads in flight per thread (the one-load-at-a-time loop
        // was latency-bound at ~2 TB/s on L2-resident tensors)
        int p = p0 + pl;
        for (; p + 3 * P < p1; p += 4 * P) {
            float f[4][8];
#pragma unroll
            for (int u = 0; u < 4; ++u) gn_load<T>(x0, C0, x1, C1, (int64_t)n * HW + p + u * P, cv * 8, f[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int k = 0; k < 8; ++k) { s[k] += f[u][k]; q[k] += f[u][k] * f[u][k]; }
        }
        for (; p < p1; p += P) {
            float f[8];
            gn_load<T>(x0, C0, x1, C1, (int64_t)n * HW + p, cv * 8, f);
#pragma unroll
            for (int k = 0; k < 8; ++k) { s[k] += f[k]; q[k] += f[k] * f[k]; }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { s_part[threadIdx.x * 16 + k] = s[k]; s_part[threadIdx.x * 16 + 8 + k] = q[k]; }
    __syncthreads();
    // thread g < groups folds its group's channels over all pixel lanes in a fixed order
    if ((int)threadIdx.x < groups) {
        const int g = threadIdx.x;
        float ss = 0.0f, qq = 0.0f;
        for (int l = 0; l < P; ++l)
            for (int c = g * cg; c < (g + 1) * cg; ++c) {
                const int t = l * nvec + (c >> 3), k = c & 7;
                ss += s_part[t * 16 + k];
                qq += s_part[t * 16 + 8 + k];
            }
        float* dst = stats + (((int64_t)n * gridDim.x + blockIdx.x) * groups + g) * 2;
        dst[0] = ss; dst[1] = qq;
    }
    // The LAST CTA of sample n to finish folds the per-chunk partials, always in chunk order (deterministic), into
    // (mean, rstd) per group, so the apply kernel starts streaming after two loads instead of a 32-step reduction per thread.
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        int ticket = atomicAdd(counters + n, 1);
        s_last = (ticket == (int)gridDim.x - 1);
        if (s_last) counters[n] = 0;   // self-resetting for the next launch
    }
    __syncthreads();
    if (s_last && (int)threadIdx.x < groups) {
        __threadfence();
        const int g = threadIdx.x;
        float s1 = 0.0f, s2 = 0.0f;
        const volatile float* src = stats + (int64_t)n * gridDim.x * groups * 2 + g * 2;
        for (int ch = 0; ch < (int)gridDim.x; ++ch) { s1 += src[(int64_t)ch * groups * 2]; s2 += src[(int64_t)ch * groups * 2 + 1]; }
        const float inv_cnt = 1.0f / ((float)cg * (float)HW);
        const float mean = s1 * inv_cnt;
        const float var = fmaxf(s2 * inv_cnt - mean * mean, 0.0f);
        finals[((int64_t)n * groups + g) * 2] = mean;
        finals[((int64_t)n * groups + g) * 2 + 1] = rsqrtf(var + eps);
    }
}

template <typename T>
__global__ void __launch_bounds__(1024, 1) gn_apply_kernel(const T* __restrict__ x0, int C0, const __nv_bfloat16* __restrict__ x1, int C1,
                                const float* __restrict__ finals, const float* __restrict__ gamma, const float* __restrict__ beta,
                                __nv_bfloat16* __restrict__ out, int HW, int groups, int apply_silu, int N, int blocks_per_sample,
                                int px_per_block) {
    // persistent: the grid is sized to what is resident (4 CTAs per SM) and strides over (sample, pixel block) items, so there is
    // no partial last wave (the one-CTA-per-chunk launch ran 512 CTAs on 444 slots: 1.15 waves)
    const int C = C0 + C1, nvec = C / 8, cg = C / groups;
    const int cv = threadIdx.x % nvec, pl = threadIdx.x / nvec, P = blockDim.x / nvec;
    if (pl >= P) return;
    float sc[8], sh[8];
    int cur_n = -1;
    const int items = N * blocks_per_sample;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int n = item / blocks_per_sample, blk = item - n * blocks_per_sample;
        if (n != cur_n) {
            cur_n = n;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                int c = cv * 8 + k, g = c / cg;
                const float mean = finals[((int64_t)n * groups + g) * 2];
                const float rstd = finals[((int64_t)n * groups + g) * 2 + 1];
                float ga = gamma[c], be = beta[c];
                sc[k] = rstd * ga; sh[k] = be - mean * rstd * ga;
            }
        }
        const int p0 = blk * px_per_block;
        const int p1 = min(p0 + px_per_block, HW);
        int p = p0 + pl;
        for (; p + 3 * P < p1; p += 4 * P) {
            float f[4][8];
#pragma unroll
            for (int u = 0; u < 4; ++u) gn_load<T>(x0, C0, x1, C1, (int64_t)n * HW + p + u * P, cv * 8, f[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float y = f[u][k] * sc[k] + sh[k];
                    f[u][k] = apply_silu ? silu_f(y) : y;
                }
                *reinterpret_cast<uint4*>(out + ((int64_t)n * HW + p + u * P) * C + cv * 8) = pack8(f[u]);
            }
        }
        for (; p < p1; p += P) {
            float f[8];
            const int64_t px = (int64_t)n * HW + p;
            gn_load<T>(x0, C0, x1, C1, px, cv * 8, f);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float y = f[k] * sc[k] + sh[k];
                f[k] = apply_silu ? silu_f(y) : y;
            }
            *reinterpret_cast<uint4*>(out + px * C + cv * 8) = pack8(f);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// GroupNorm with statistics from the PRODUCER's epilogue (north_star (b), include/gmd_b200.h "GroupNorm fused into the producing
// convolution / GEMM"): the producer's epilogue accumulates per-(sample, channel pair) sums as 64-bit fixed-point integers;
// gn_apply_sums_kernel is one streaming pass (read, scale / shift, SiLU, write) whose statistics are formed from the channel-pair
// sums of its one or two sources.  No statistics pass over the activation exists any more.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(320) gn_apply_sums_kernel(const T* __restrict__ x0, int C0, const long long* __restrict__ sums0,
                                                            const __nv_bfloat16* __restrict__ x1, int C1, const long long* __restrict__ sums1,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            __nv_bfloat16* __restrict__ out, int HW, int groups, int apply_silu, int px_per_cta, float eps) {
    __shared__ float s_stat[128];           // [groups][2] mean, rstd
    const int C = C0 + C1, nvec = C / 8, cg = C / groups;
    const int n = blockIdx.y;
    const int P = blockDim.x / nvec;        // blockDim.x == nvec * P: every thread owns a channel vector for its whole life
    const int cv = threadIdx.x % nvec, pl = threadIdx.x / nvec;
    const int p0 = blockIdx.x * px_per_cta, p1 = min(p0 + px_per_cta, HW);
    constexpr int U = 4;
    // the first batch of activations is requested BEFORE the statistics are formed (their loads hide the prologue's round trips)
    int p = p0 + pl;
    float f[U][8];
    const bool first_full = p + (U - 1) * P < p1;
    if (first_full) {
#pragma unroll
        for (int u = 0; u < U; ++u) gn_load<T>(x0, C0, x1, C1, (int64_t)n * HW + p + u * P, cv * 8, f[u]);
    }
    if ((int)threadIdx.x < groups) {
        // group g = channels [g*cg, (g+1)*cg) = channel pairs [g*cg/2, (g+1)*cg/2) of the concatenation (cg is even); a pair lives
        // entirely in one source (C0 is even)
        const int g = threadIdx.x;
        long long i1 = 0, i2 = 0;       // exact integer sums of the fixed-point accumulators
        for (int pr = g * cg / 2; pr < (g + 1) * cg / 2; ++pr) {
            const longlong2 v = pr < C0 / 2 ? __ldcg(reinterpret_cast<const longlong2*>(sums0) + (int64_t)n * (C0 / 2) + pr)
                                            : __ldcg(reinterpret_cast<const longlong2*>(sums1) + (int64_t)n * (C1 / 2) + (pr - C0 / 2));
            i1 += v.x; i2 += v.y;
        }
        const double inv_cnt = 1.0 / ((double)cg * (double)HW * 16777216.0);
        const double dm = (double)i1 * inv_cnt;
        const float mean = (float)dm;
        const float var = fmaxf((float)((double)i2 * inv_cnt - dm * dm), 0.0f);
        s_stat[g * 2] = mean;
        s_stat[g * 2 + 1] = rsqrtf(var + eps);
    }
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + cv * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + cv * 8) + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + cv * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + cv * 8) + 1);
    __syncthreads();
    float sc[8], sh[8];
    {
        const float ga[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int g = (cv * 8 + k) / cg;
            const float mean = s_stat[g * 2], rstd = s_stat[g * 2 + 1];
            sc[k] = rstd * ga[k]; sh[k] = fmaf(-mean, rstd * ga[k], be[k]);
        }
    }
    if (first_full) {
        for (;;) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float y = fmaf(f[u][k], sc[k], sh[k]);
                    f[u][k] = apply_silu ? silu_f(y) : y;
                }
                *reinterpret_cast<uint4*>(out + ((int64_t)n * HW + p + u * P) * C + cv * 8) = pack8(f[u]);
            }
            p += U * P;
            if (!(p + (U - 1) * P < p1)) break;
#pragma unroll
            for (int u = 0; u < U; ++u) gn_load<T>(x0, C0, x1, C1, (int64_t)n * HW + p + u * P, cv * 8, f[u]);
        }
    }
    for (; p < p1; p += P) {
        float g[8];
        const int64_t px = (int64_t)n * HW + p;
        gn_load<T>(x0, C0, x1, C1, px, cv * 8, g);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float y = fmaf(g[k], sc[k], sh[k]);
            g[k] = apply_silu ? silu_f(y) : y;
        }
        *reinterpret_cast<uint4*>(out + px * C + cv * 8) = pack8(g);
    }
}

// ---------------------------------------------------------------------------------------------
// One-pass GroupNorm for everything that fits on chip (all UNet shapes up to 1024^2 images): a thread-block CLUSTER owns
// (sample, slab of whole groups) and splits the pixels between its CTAs; every CTA parks its part of the input in shared
// memory, the per-group sums are exchanged through distributed shared memory in rank order (fixed order: deterministic
// and independent of the batch), and the normalised + SiLU'd bf16 output is produced from the parked copy.  The tensor is
// read once and written once by ONE launch — the two-kernel path below reads it twice and was wave-quantised / latency-bound
// (41.7 + 18.0 us for a 42 MB tensor, profiles/ncu_gn_r01_summary.txt).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cl_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cl_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cl_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem(const float* local, uint32_t rank) {
    uint32_t a = smem_u32(local), r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(r) : "memory");
    return v;
}

template <typename T> struct ParkVec;            // how 8 channels are parked in shared memory
template <> struct ParkVec<__nv_bfloat16> { using type = uint4; };
template <> struct ParkVec<float> { struct alignas(16) type { float4 a, b; }; };

template <typename T>
__device__ __forceinline__ void park_load(const T* x0, int C0, const __nv_bfloat16* x1, int C1, int64_t px, int c, typename ParkVec<T>::type& raw, float (&f)[8]);
template <>
__device__ __forceinline__ void park_load<__nv_bfloat16>(const __nv_bfloat16* x0, int C0, const __nv_bfloat16* x1, int C1, int64_t px, int c, uint4& raw, float (&f)[8]) {
    raw = (c < C0) ? __ldg(reinterpret_cast<const uint4*>(x0 + px * C0 + c)) : __ldg(reinterpret_cast<const uint4*>(x1 + px * C1 + (c - C0)));
    unpack8(raw, f);
}
template <>
__device__ __forceinline__ void park_load<float>(const float* x0, int C0, const __nv_bfloat16* x1, int C1, int64_t px, int c, ParkVec<float>::type& raw, float (&f)[8]) {
    gn_load<float>(x0, C0, x1, C1, px, c, f);
    raw.a = make_float4(f[0], f[1], f[2], f[3]); raw.b = make_float4(f[4], f[5], f[6], f[7]);
}
__device__ __forceinline__ void park_read(const uint4& raw, float (&f)[8]) { unpack8(raw, f); }
__device__ __forceinline__ void park_read(const ParkVec<float>::type& raw, float (&f)[8]) {
    f[0] = raw.a.x; f[1] = raw.a.y; f[2] = raw.a.z; f[3] = raw.a.w; f[4] = raw.b.x; f[5] = raw.b.y; f[6] = raw.b.z; f[7] = raw.b.w;
}

template <typename T>
__global__ void __launch_bounds__(256, 5) gn_fused_kernel(const T* __restrict__ x0, int C0, const __nv_bfloat16* __restrict__ x1, int C1,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       __nv_bfloat16* __restrict__ out, int HW, int groups, int apply_silu, float eps,
                                                       int V, int px_per_cta) {
    using PV = typename ParkVec<T>::type;
    extern __shared__ __align__(16) uint8_t gn_smem[];
    float* s_warp = reinterpret_cast<float*>(gn_smem);          // [8 warps][V][16] per-warp channel sums
    float* s_ch = s_warp + 8 * 16 * 16;                         // [V*8][2]  per-channel sums of this CTA
    float* s_grp = s_ch + 128 * 2;                              // [G][2]    per-group sums of this CTA (read by the cluster)
    float* s_stat = s_grp + 16 * 2;                             // [G][2]    mean, rstd
    PV* park = reinterpret_cast<PV*>(s_stat + 16 * 2);
    const int C = C0 + C1, cg = C / groups, G = V * 8 / cg;
    const uint32_t CLN = cl_size(), rank = cl_rank();
    const int slab = blockIdx.x / CLN, n = blockIdx.y;
    const int cv = threadIdx.x % V, pl = threadIdx.x / V, P = blockDim.x / V;
    const int c0 = (slab * V + cv) * 8;                         // first channel of this thread's vector
    const int p0 = rank * px_per_cta;
    const int npx = max(0, min(px_per_cta, HW - p0));
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float ga[8], be[8];
    if (pl < P) {
        // affine parameters: requested now, needed only in the apply phase (their latency hides behind everything else)
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0) + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c0) + 1);
        ga[0] = g0.x; ga[1] = g0.y; ga[2] = g0.z; ga[3] = g0.w; ga[4] = g1.x; ga[5] = g1.y; ga[6] = g1.z; ga[7] = g1.w;
        be[0] = b0.x; be[1] = b0.y; be[2] = b0.z; be[3] = b0.w; be[4] = b1.x; be[5] = b1.y; be[6] = b1.z; be[7] = b1.w;
        // Park this thread's vectors with cp.async: ALL of them are in flight at once (a register-staged loop paid one memory
        // round trip per 4 vectors, ~5 in a row per CTA, and that chain — not bandwidth — set the kernel time).
        const bool src1 = c0 >= C0;
        const char* base = src1 ? reinterpret_cast<const char*>(x1 + (int64_t)(n * (int64_t)HW + p0) * C1 + (c0 - C0))
                                : reinterpret_cast<const char*>(x0 + (int64_t)(n * (int64_t)HW + p0) * C0 + c0);
        const int64_t pitch = src1 ? (int64_t)C1 * 2 : (int64_t)C0 * (int64_t)sizeof(T);
        const bool wide = !src1 && sizeof(T) == 4;                      // fp32 source: 32 bytes per vector
        if (sizeof(T) == 4 && src1) {
            // bf16 skip tensor behind an fp32 first source: widen through registers (not on the UNet path)
            for (int p = pl; p < npx; p += P) {
                PV raw; float f[8];
                park_load<T>(x0, C0, x1, C1, (int64_t)n * HW + p0 + p, c0, raw, f);
                park[p * V + cv] = raw;
            }
        } else {
            for (int p = pl; p < npx; p += P) {
                const uint32_t dst = smem_u32(park + p * V + cv);
                const char* src = base + p * pitch;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
                if (wide) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16), "l"(src + 16) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        // each thread reads back exactly what it copied: no barrier needed
        for (int p = pl; p < npx; p += P) {
            float f[8];
            park_read(park[p * V + cv], f);
#pragma unroll
            for (int k = 0; k < 8; ++k) { s[k] += f[k]; q[k] += f[k] * f[k]; }
        }
    }
    // lanes L, L+V, L+2V, ... of a warp own the same channel vector: fold them with a shuffle tree (fixed order), so only V lanes
    // per warp go through shared memory
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int dmax = V;
    while (dmax * 2 < 32) dmax *= 2;
    for (int d = dmax; d >= V; d >>= 1) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float a = __shfl_down_sync(0xffffffffu, s[k], d), b = __shfl_down_sync(0xffffffffu, q[k], d);
            if (lane + d < 32) { s[k] += a; q[k] += b; }
        }
    }
    if (lane < V) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { s_warp[(warp * 16 + cv) * 16 + k] = s[k]; s_warp[(warp * 16 + cv) * 16 + 8 + k] = q[k]; }
    }
    __syncthreads();
    // per channel (and per moment): fold the 8 warps in warp order
    if ((int)threadIdx.x < V * 16) {
        const int ch = threadIdx.x >> 1, m = threadIdx.x & 1;        // channel within the slab, 0 = sum / 1 = sum of squares
        const int v = ch >> 3, k = ch & 7;
        float acc = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) acc += s_warp[(w * 16 + v) * 16 + m * 8 + k];
        s_ch[ch * 2 + m] = acc;
    }
    __syncthreads();
    if ((int)threadIdx.x < G * 2) {
        const int g = threadIdx.x >> 1, m = threadIdx.x & 1;
        float acc = 0.0f;
        for (int c = g * cg; c < (g + 1) * cg; ++c) acc += s_ch[c * 2 + m];
        s_grp[g * 2 + m] = acc;
    }
    cl_sync();                                                       // every CTA's s_grp is complete and visible cluster-wide
    if ((int)threadIdx.x < G) {
        const int g = threadIdx.x;
        float s1 = 0.0f, s2 = 0.0f;
        for (uint32_t r = 0; r < CLN; ++r) { s1 += ld_dsmem(s_grp + g * 2, r); s2 += ld_dsmem(s_grp + g * 2 + 1, r); }
        const float inv_cnt = 1.0f / ((float)cg * (float)HW);
        const float mean = s1 * inv_cnt;
        const float var = fmaxf(s2 * inv_cnt - mean * mean, 0.0f);
        s_stat[g * 2] = mean;
        s_stat[g * 2 + 1] = rsqrtf(var + eps);
    }
    // second cluster barrier, split: arrive now (this CTA's remote reads are done), wait only before exiting, so that no CTA
    // retires while a neighbour may still read its s_grp — the apply phase runs in between
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    __syncthreads();                                                 // s_stat visible to the CTA
    if (pl < P) {
        float sc[8], sh[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int cl = cv * 8 + k, g = cl / cg;
            const float mean = s_stat[g * 2], rstd = s_stat[g * 2 + 1];
            sc[k] = rstd * ga[k]; sh[k] = be[k] - mean * rstd * ga[k];
        }
        for (int p = pl; p < npx; p += P) {
            float f[8];
            park_read(park[p * V + cv], f);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float y = f[k] * sc[k] + sh[k];
                f[k] = apply_silu ? silu_f(y) : y;
            }
            *reinterpret_cast<uint4*>(out + ((int64_t)n * HW + p0 + p) * C + c0) = pack8(f);
        }
    }
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per token row; the row lives in registers between the two passes.
// ---------------------------------------------------------------------------------------------
template <typename T, int MAXV>  // vectors (of 8) per lane
__global__ void __launch_bounds__(256) layernorm_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, __nv_bfloat16* __restrict__ out,
                                                        int64_t M, int C, float eps) {
    const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int lane = threadIdx.x & 31, nvec = C / 8;
    float v[MAXV][8];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        int vi = lane + i * 32;
        if (vi < nvec) {
            load8<T>(x, row * C + vi * 8, v[i]);
#pragma unroll
            for (int k = 0; k < 8; ++k) s += v[i][k];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(~0u, s, o);
    const float mean = s / (float)C;
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        int vi = lane + i * 32;
        if (vi < nvec) {
#pragma unroll
            for (int k = 0; k < 8; ++k) { float d = v[i][k] - mean; q += d * d; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(~0u, q, o);
    const float rstd = rsqrtf(q / (float)C + eps);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        int vi = lane + i * 32;
        if (vi < nvec) {
            float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + vi * 2), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + vi * 2 + 1);
            float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + vi * 2), b1 = __ldg(reinterpret_cast<const float4*>(beta) + vi * 2 + 1);
            float ga[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            float f[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = (v[i][k] - mean) * rstd * ga[k] + be[k];
            reinterpret_cast<uint4*>(out + row * C)[vi] = pack8(f);
        }
    }
}

// fp32 token stream, C = LPR * 4 * V4 exactly (the UNet's 320 / 640 / 1280): LPR lanes per row, V4 float4 per lane, so every lane of
// every load is busy (the generic kernel's 8-element vectors leave 24 of 32 lanes idle in the second round at C = 320) and a warp
// at C = 320 keeps two rows in flight.  Same arithmetic order per row for any M: batch-independent.
template <int LPR, int V4>
__global__ void __launch_bounds__(256) layernorm_f32_exact_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, __nv_bfloat16* __restrict__ out,
                                                                  int64_t M, float eps) {
    constexpr int C = LPR * 4 * V4, RPW = 32 / LPR;   // rows per warp
    const int lane = threadIdx.x & 31, sub = lane % LPR;
    const int64_t row = (blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LPR;
    const bool live = row < M;
    const float4* xr = reinterpret_cast<const float4*>(x + (live ? row : 0) * C);
    float4 v[V4];
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
        v[i] = live ? __ldg(xr + sub + i * LPR) : make_float4(0.f, 0.f, 0.f, 0.f);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(~0u, s, o);
    const float mean = s * (1.0f / (float)C);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < V4; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(~0u, q, o);
    const float rstd = rsqrtf(q * (1.0f / (float)C) + eps);
    if (!live) return;
    uint2* orow = reinterpret_cast<uint2*>(out + row * C);
#pragma unroll
    for (int i = 0; i < V4; ++i) {
        const int vi = sub + i * LPR;
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + vi), b = __ldg(reinterpret_cast<const float4*>(beta) + vi);
        orow[vi] = make_uint2(pack_bf16x2((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y),
                              pack_bf16x2((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w));
    }
}

// ---------------------------------------------------------------------------------------------
// Row softmax (bf16 in/out, fp32 math): one CTA per row, three passes (the row stays in L1/L2).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_rows_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                           int64_t N, float scale, int n_valid, int causal_period) {
    // columns >= n_valid (padding of the key dimension) and, with causal_period > 0, columns > (row % causal_period) (CLIP's causal
    // text attention: one causal_period x causal_period matrix per head) get probability 0
    __shared__ float red[32];
    const __nv_bfloat16* xr = x + blockIdx.x * N;
    __nv_bfloat16* orow = out + blockIdx.x * N;
    const int nvec = (int)(N / 8);
    const int lim = causal_period > 0 ? min(n_valid, (int)(blockIdx.x % causal_period) + 1) : n_valid;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
        float f[8]; unpack8(__ldg(reinterpret_cast<const uint4*>(xr) + i), f);
#pragma unroll
        for (int k = 0; k < 8; ++k) if (i * 8 + k < lim) mx = fmaxf(mx, f[k]);
    }
    auto block_reduce = [&](float v, bool is_max) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { float t = __shfl_xor_sync(~0u, v, o); v = is_max ? fmaxf(v, t) : v + t; }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        __syncthreads();
        float r = red[0];
        for (int i = 1; i < (blockDim.x >> 5); ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
        return r;
    };
    mx = block_reduce(mx, true);
    const float sl2 = scale * 1.4426950408889634f;
    float sum = 0.0f;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
        float f[8]; unpack8(__ldg(reinterpret_cast<const uint4*>(xr) + i), f);
#pragma unroll
        for (int k = 0; k < 8; ++k) if (i * 8 + k < lim) sum += exp2f((f[k] - mx) * sl2);
    }
    sum = block_reduce(sum, false);
    const float inv = 1.0f / sum;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
        float f[8]; unpack8(__ldg(reinterpret_cast<const uint4*>(xr) + i), f);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = i * 8 + k < lim ? exp2f((f[k] - mx) * sl2) * inv : 0.0f;
        reinterpret_cast<uint4*>(orow)[i] = pack8(f);
    }
}

// diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0): [cos(t f_k) | sin(t f_k)], f_k = exp(-ln(1e4) k / half)
__global__ void timestep_embedding_kernel(float t, __nv_bfloat16* __restrict__ out, int B, int dim) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * dim) return;
    int c = i % dim, half = dim / 2;
    int k = c < half ? c : c - half;
    float freq = expf(-9.210340371976184f * (float)k / (float)half);
    float a = t * freq;
    out[i] = __float2bfloat16(c < half ? cosf(a) : sinf(a));
}

__global__ void silu_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = __float2bfloat16(silu_f(__bfloat162float(x[i])));
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// GMD_GN_TWO_PASS=1 forces the stats + apply pair (A/B measurements; the VAE's >= 256x256 planes always use it)
const bool g_gn_two_pass = [] { const char* e = getenv("GMD_GN_TWO_PASS"); return e && e[0] == '1'; }();

template <typename T>
int gn_launch(const void* x0, int C0, const void* x1, int C1, const float* gamma, const float* beta, void* out, int N, int HW, int groups,
              float eps, int apply_silu, float* stats_ws, cudaStream_t st) {
    const int nvec = (C0 + C1) / 8;
    {
        // one-pass cluster kernel whenever a sample's slab fits the shared memory of <= 8 CTAs.  The plan depends on (HW, C, groups,
        // dtype) only — never on N — so a sample's result does not depend on the batch it is in.
        const int cgc = (C0 + C1) / groups;
        int a = 8, b = cgc;
        while (b) { int t = a % b; a = b; b = t; }
        int V = cgc / a;                                          // lcm(8, cg) / 8 vectors hold whole groups
        while (V < 4 && nvec % (V * 2) == 0) V *= 2;              // >= 64 contiguous bytes per pixel where the shape allows
        const int G = V * 8 / cgc;
        const size_t vec_bytes = sizeof(typename ParkVec<T>::type), fixed = (8 * 16 * 16 + 128 * 2 + 16 * 2 + 16 * 2) * sizeof(float);
        if (V <= 16 && G <= 16 && nvec % V == 0 && !g_gn_two_pass) {
            // parked bytes per CTA <= 36 KB (+ 9.5 KB of reduction scratch): >= 4 CTAs per SM, whose load / reduce / apply phases overlap
            int cl = 1;
            while (cl < 8 && (size_t)((HW + cl - 1) / cl) * V * vec_bytes > (36u << 10)) cl *= 2;
            const int px_per_cta = (HW + cl - 1) / cl;
            const size_t smem = fixed + (size_t)px_per_cta * V * vec_bytes;
            // beyond 42 KB per CTA at the largest cluster the tensor cannot be parked in about two waves: the stats + apply pair
            // (second read from L2) streams better
            if ((size_t)px_per_cta * V * vec_bytes <= (42u << 10)) {
                static bool attr_done[kMaxDevices] = {};   // (one array per instantiation of this template)
                const int dev = device_ordinal();
                if (!attr_done[dev]) {
                    cudaFuncSetAttribute(gn_fused_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 56 << 10);
                    attr_done[dev] = true;
                }
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3((unsigned)(cl * (nvec / V)), (unsigned)N, 1);
                cfg.blockDim = dim3(256, 1, 1);
                cfg.dynamicSmemBytes = smem;
                cfg.stream = st;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr; cfg.numAttrs = 1;
                cudaError_t e = cudaLaunchKernelEx(&cfg, gn_fused_kernel<T>, static_cast<const T*>(x0), C0, static_cast<const __nv_bfloat16*>(x1), C1, gamma, beta,
                                                   static_cast<__nv_bfloat16*>(out), HW, groups, apply_silu, eps, V, px_per_cta);
                if (e != cudaSuccess) { set_last_error("gn_fused launch: %s", cudaGetErrorString(e)); return kErrCuda; }
                count_launch(1);
                return check_launch("groupnorm(fused)");
            }
        }
    }
    int P = 256 / nvec; if (P < 1) P = 1;
    int threads = (nvec * P + 31) / 32 * 32;
    // chunking depends on (HW, C) only, never on N: a sample's statistics are bit-identical for any batch size / sharding
    int px_per_cta = (HW + 31) / 32;
    if (px_per_cta < P * 4) px_per_cta = P * 4;
    int chunks = (HW + px_per_cta - 1) / px_per_cta;  // <= 32
    dim3 grid(chunks, N);
    // workspace: [1024] arrival counters at a FIXED offset (zero before first use, self-resetting; calls with different N share
    // the workspace) | [N][32 chunks][groups][2] partials | [N][groups][2] finals
    int* counters = reinterpret_cast<int*>(stats_ws);
    float* partials = stats_ws + 1024;
    float* finals = partials + (int64_t)N * 32 * groups * 2;
    gn_stats_kernel<T><<<grid, threads, threads * 16 * sizeof(float), st>>>(static_cast<const T*>(x0), C0, static_cast<const __nv_bfloat16*>(x1), C1,
                                                                            partials, finals, counters, HW, groups, px_per_cta, eps);
    const int px_per_block = 8 * P;
    const int blocks_per_sample = (HW + px_per_block - 1) / px_per_block;
    int64_t agrid = (int64_t)N * blocks_per_sample;
    if (agrid > 4 * 148) agrid = 4 * 148;
    gn_apply_kernel<T><<<(unsigned)agrid, threads, 0, st>>>(static_cast<const T*>(x0), C0, static_cast<const __nv_bfloat16*>(x1), C1, finals, gamma, beta,
                                                            static_cast<__nv_bfloat16*>(out), HW, groups, apply_silu, N, blocks_per_sample, px_per_block);
    count_launch(2);
    return check_launch("groupnorm");
}

}  // namespace
}  // namespace gmd

extern "C" int gmd_groupnorm_silu(const void* x0, int32_t C0, const void* x1, int32_t C1, const float* gamma, const float* beta,
                                  void* out, int32_t N, int32_t HW, int32_t groups, float eps, int32_t apply_silu, int32_t in_dtype,
                                  float* stats_ws, void* stream) {
    using namespace gmd;
    if (!x0 || !gamma || !beta || !out || !stats_ws) { set_last_error("gmd_groupnorm_silu: null pointer"); return kErrInvalid; }
    if (!x1) C1 = 0;
    const int C = C0 + C1;
    if (C0 % 8 || C1 % 8 || groups <= 0 || C % groups || N <= 0 || HW <= 0) { set_last_error("gmd_groupnorm_silu: bad shape C0=%d C1=%d groups=%d", C0, C1, groups); return kErrInvalid; }
    if (!al16(x0) || !al16(x1) || !al16(out) || !al16(gamma) || !al16(beta)) { set_last_error("gmd_groupnorm_silu: pointers must be 16-byte aligned"); return kErrInvalid; }
    if (C / 8 > 1024 || groups > 256 || N > 1024) { set_last_error("gmd_groupnorm_silu: C=%d groups=%d too large", C, groups); return kErrUnsupported; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (in_dtype == GMD_BF16) return gn_launch<__nv_bfloat16>(x0, C0, x1, C1, gamma, beta, out, N, HW, groups, eps, apply_silu, stats_ws, st);
    if (in_dtype == GMD_F32) return gn_launch<float>(x0, C0, x1, C1, gamma, beta, out, N, HW, groups, eps, apply_silu, stats_ws, st);
    set_last_error("gmd_groupnorm_silu: unknown in_dtype %d", in_dtype);
    return kErrInvalid;
}

extern "C" int gmd_groupnorm_apply(const void* x0, int32_t C0, const void* sums0, const void* x1, int32_t C1, const void* sums1,
                                   const float* gamma, const float* beta, void* out, int32_t N, int32_t HW, int32_t groups, float eps,
                                   int32_t apply_silu, int32_t in_dtype, void* stream) {
    using namespace gmd;
    if (!x0 || !sums0 || !gamma || !beta || !out) { set_last_error("gmd_groupnorm_apply: null pointer"); return kErrInvalid; }
    if (!x1) C1 = 0;
    if (x1 && !sums1) { set_last_error("gmd_groupnorm_apply: second source without statistics"); return kErrInvalid; }
    const int C = C0 + C1;
    if (C0 % 8 || C1 % 8 || groups <= 0 || groups > 64 || C % groups || ((C / groups) % 2) || N <= 0 || HW <= 0 || C / 8 > 320) {
        set_last_error("gmd_groupnorm_apply: bad shape C0=%d C1=%d groups=%d", C0, C1, groups); return kErrInvalid;
    }
    if (!al16(x0) || !al16(x1) || !al16(out) || !al16(gamma) || !al16(beta) || !al16(sums0) || !al16(sums1)) {
        set_last_error("gmd_groupnorm_apply: pointers must be 16-byte aligned"); return kErrInvalid;
    }
    const int nvec = C / 8;
    int P = 256 / nvec; if (P < 1) P = 1;
    const int threads = nvec * P;
    if (threads < groups) { set_last_error("gmd_groupnorm_apply: C too small for %d groups", groups); return kErrUnsupported; }
    const int ppc = 16 * P;                                   // four batches of four loads per thread
    const dim3 grid((HW + ppc - 1) / ppc, N);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (in_dtype == GMD_BF16)
        gn_apply_sums_kernel<__nv_bfloat16><<<grid, threads, 0, st>>>(static_cast<const __nv_bfloat16*>(x0), C0, static_cast<const long long*>(sums0), static_cast<const __nv_bfloat16*>(x1), C1, static_cast<const long long*>(sums1),
                                                                      gamma, beta, static_cast<__nv_bfloat16*>(out), HW, groups, apply_silu, ppc, eps);
    else if (in_dtype == GMD_F32)
        gn_apply_sums_kernel<float><<<grid, threads, 0, st>>>(static_cast<const float*>(x0), C0, static_cast<const long long*>(sums0), static_cast<const __nv_bfloat16*>(x1), C1, static_cast<const long long*>(sums1),
                                                              gamma, beta, static_cast<__nv_bfloat16*>(out), HW, groups, apply_silu, ppc, eps);
    else { set_last_error("gmd_groupnorm_apply: unknown in_dtype %d", in_dtype); return kErrInvalid; }
    count_launch(1);
    return check_launch("groupnorm_apply");
}

extern "C" int gmd_layernorm(const void* x, const float* gamma, const float* beta, void* out, int64_t M, int32_t C, float eps,
                             int32_t in_dtype, void* stream) {
    using namespace gmd;
    if (!x || !gamma || !beta || !out) { set_last_error("gmd_layernorm: null pointer"); return kErrInvalid; }
    if (C % 8 || C <= 0 || C > 8 * 32 * 8) { set_last_error("gmd_layernorm: C=%d unsupported (multiple of 8, <= 2048)", C); return kErrInvalid; }
    if (!al16(x) || !al16(out) || !al16(gamma) || !al16(beta)) { set_last_error("gmd_layernorm: pointers must be 16-byte aligned"); return kErrInvalid; }
    if (in_dtype != GMD_BF16 && in_dtype != GMD_F32) { set_last_error("gmd_layernorm: unknown in_dtype %d", in_dtype); return kErrInvalid; }
    if (M == 0) return kOk;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned grid = (unsigned)((M + 7) / 8);
    const int nvec = C / 8;
    auto* op = static_cast<__nv_bfloat16*>(out);
#define GMD_LN(T, V) layernorm_kernel<T, V><<<grid, 256, 0, st>>>(static_cast<const T*>(x), gamma, beta, op, M, C, eps)
    if (in_dtype == GMD_BF16) {
        if (nvec <= 64) GMD_LN(__nv_bfloat16, 2); else if (nvec <= 160) GMD_LN(__nv_bfloat16, 5); else GMD_LN(__nv_bfloat16, 8);
    } else {
        const float* xf = static_cast<const float*>(x);
        if (C == 320) layernorm_f32_exact_kernel<16, 5><<<(unsigned)((M + 15) / 16), 256, 0, st>>>(xf, gamma, beta, op, M, eps);
        else if (C == 640) layernorm_f32_exact_kernel<32, 5><<<grid, 256, 0, st>>>(xf, gamma, beta, op, M, eps);
        else if (C == 1280) layernorm_f32_exact_kernel<32, 10><<<grid, 256, 0, st>>>(xf, gamma, beta, op, M, eps);
        else if (nvec <= 64) GMD_LN(float, 2); else if (nvec <= 160) GMD_LN(float, 5); else GMD_LN(float, 8);
    }
#undef GMD_LN
    count_launch(1);
    return check_launch("layernorm");
}

extern "C" int gmd_softmax_rows(const void* x, void* out, int64_t M, int64_t N, float scale, void* stream) {
    using namespace gmd;
    if (!x || !out || N % 8 || !al16(x) || !al16(out)) { set_last_error("gmd_softmax_rows: bad arguments"); return kErrInvalid; }
    if (M == 0 || N == 0) return kOk;
    softmax_rows_kernel<<<(unsigned)M, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), N, scale, (int)N, 0);
    count_launch(1);
    return check_launch("softmax_rows");
}

extern "C" int gmd_softmax_rows_masked(const void* x, void* out, int64_t M, int64_t N, float scale, int32_t n_valid, int32_t causal_period, void* stream) {
    using namespace gmd;
    if (!x || !out || N % 8 || !al16(x) || !al16(out) || n_valid <= 0 || n_valid > N || causal_period < 0) { set_last_error("gmd_softmax_rows_masked: bad arguments"); return kErrInvalid; }
    if (M == 0 || N == 0) return kOk;
    softmax_rows_kernel<<<(unsigned)M, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), N, scale, n_valid, causal_period);
    count_launch(1);
    return check_launch("softmax_rows_masked");
}

extern "C" int gmd_timestep_embedding(float t, void* out, int32_t B, int32_t dim, void* stream) {
    using namespace gmd;
    if (!out || B <= 0 || dim <= 0 || dim % 2) { set_last_error("gmd_timestep_embedding: bad arguments"); return kErrInvalid; }
    timestep_embedding_kernel<<<(B * dim + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(t, static_cast<__nv_bfloat16*>(out), B, dim);
    count_launch(1);
    return check_launch("timestep_embedding");
}

extern "C" int gmd_silu(const void* x, void* out, int64_t n, void* stream) {
    using namespace gmd;
    if (!x || !out) { set_last_error("gmd_silu: null pointer"); return kErrInvalid; }
    if (n == 0) return kOk;
    silu_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), n);
    count_launch(1);
    return check_launch("silu");
}
