// Kernel (d): SDR + gain map -> HDR (Eq. 1) -> tone-mapping operator -> BT.2020->709 gamut, with min/max
// statistics, as ONE coalesced streaming pass (HBM-bound: 2 x 12 B read + 12 B written per pixel).
// Arithmetic follows gm_diffusion/stage1/tone_mapping.py:14-90 of the reference op for op.
#include "common.cuh"
#include "../../include/gmd_b200.h"

namespace gmd {
void count_launch(int n);

namespace {

struct HdrConsts {
    float qmax, eps, hi;          // hi = qmax + 1
    float inv_hi;                 // 1 / (qmax + 1)
    float mu, inv_log1p_mu;
    float log2_hi;                // for the optional exponential gain
    float rgbe_div;               // RGBE encodes hdr / rgbe_div (save_hdr_image's division by qmax + 1)
    float rgbe_min;               // smallest float >= 1e-32 (the C comparison `v < 1e-32` is done in double)
    int flags, tmo;
};

struct ByteOuts {
    uint8_t* rgbe;     // [pixels, 4] R,G,B,E (3-channel layouts only)
    uint8_t* sdr_u8;   // same element order as the input
    uint8_t* gm_u8;
};

// Radiance shared-exponent pixel, as OpenCV's HdrEncoder quantises it (Ward's float2rgbe): v = max(r,g,b); v < 1e-32 -> 0;
// else (m, e) = frexp(v); scale = m * 256 / v; bytes = trunc(c * scale), e + 128.  m * 256 / v is exactly 2^(8-e), so the scale
// is built from the exponent bits and the products are exact: the result is bit-identical to the C code.
__device__ __forceinline__ uint32_t rgbe_pack(float r, float g, float b, float vmin) {
    float v = fmaxf(r, fmaxf(g, b));
    if (!(v >= vmin)) return 0u;
    int e = (int)((__float_as_uint(v) >> 23) & 0xffu) - 126;
    float scale = __uint_as_float((uint32_t)(127 + 8 - e) << 23);
    int R = min(max((int)(r * scale), 0), 255), G = min(max((int)(g * scale), 0), 255), B = min(max((int)(b * scale), 0), 255);
    return (uint32_t)R | ((uint32_t)G << 8) | ((uint32_t)B << 16) | ((uint32_t)((e + 128) & 0xff) << 24);
}

// (x * 255).astype(uint8) of generate_hdr.py:243-244 (x in [0,1]; truncation)
__device__ __forceinline__ uint32_t to_u8(float x) { return (uint32_t)min(max((int)__fmul_rn(x, 255.0f), 0), 255); }

// x^2.2 on [0,1] as x*x * 2^(0.2*log2 x): the MUFU approximations only see the 0.2 exponent, so their
// absolute error (2^-22 on lg2) is damped 11x compared with a direct 2^(2.2*log2 x): <= ~4e-7 relative.
__device__ __forceinline__ float pow22_unit(float x) {
    float x2 = x * x;
    float r = exp2f(0.2f * __log2f(x));   // x = 0 -> log2 = -inf -> exp2 = 0
    return x2 * r;
}

__device__ __forceinline__ float denorm(float v, int flags) {
    return (flags & GMD_HDR_DENORM) ? fminf(fmaxf(v * 0.5f + 0.5f, 0.0f), 1.0f) : v;
}

// Eq.(1), tone_mapping.py:69-71
__device__ __forceinline__ float eq1(float sdr, float gm, const HdrConsts& c) {
    float s = fminf(fmaxf(sdr, 0.0f), 1.0f);
    float lin = pow22_unit(s);
    float gain = (c.flags & GMD_HDR_EXP_GAIN) ? exp2f(gm * c.log2_hi) : (1.0f + gm * c.qmax);
    float hdr = (lin + c.eps) * gain - c.eps;
    if (c.flags & GMD_HDR_CLAMP_OUT) hdr = fminf(fmaxf(hdr, 0.0f), c.hi);
    return hdr;
}

// log1p(z) for z >= 0 as ln(1 + z) through MUFU.LG2: absolute error <= ~2e-7 (2^-22 on lg2, 2^-24 relative on 1 + z), i.e.
// <= 4e-8 on the tone-mapped value after the division by log1p(mu) >= 6.2 — far inside the 1e-6 + 1e-5*|ref| gate, and it
// takes 3 instructions instead of log1pf's ~25 (the standalone TMO variants were compute-bound at 0.59 of HBM peak).
__device__ __forceinline__ float log1p_pos(float z) { return __logf(1.0f + z); }

// tone_mapping.py:14-47
__device__ __forceinline__ float tmo_apply(float x, const HdrConsts& c) {
    switch (c.tmo) {
        case GMD_TMO_LINEAR: return x * c.inv_hi;
        case GMD_TMO_HARD_CLIP: return fminf(fmaxf(x, 0.0f), 1.0f);
        case GMD_TMO_MULOG: {
            float y = x * c.inv_hi;
            const float m = c.mu * y;
            float t = log1p_pos(fmaxf(m, 0.0f)) * c.inv_log1p_mu;       // (-1, 0): a negative logarithm, clamped to 0 below like the reference's
            t = fminf(fmaxf(t, 0.0f), 1.0f);
            return m >= -1.0f ? t : __int_as_float(0x7fc00000);         // NaN input or the log of a negative number: NaN, as torch.log gives (tone_mapping.py:33-40)
        }
        case GMD_TMO_CUDA: {
            float y = fminf(fmaxf(x * 0.1f, 0.0f), 1.0f);
            return log1p_pos(c.mu * y) * c.inv_log1p_mu;
        }
        default: return x;
    }
}

// tone_mapping.py:78-90 (rows of M applied to the pixel's rgb column vector, then clamp)
__device__ __forceinline__ void gamut709(float& r, float& g, float& b) {
    float r2 = 1.660491f * r + -0.587641f * g + -0.072850f * b;
    float g2 = -0.124550f * r + 1.132900f * g + -0.008349f * b;
    float b2 = -0.018151f * r + -0.100579f * g + 1.118730f * b;
    r = fminf(fmaxf(r2, 0.0f), 1.0f);
    g = fminf(fmaxf(g2, 0.0f), 1.0f);
    b = fminf(fmaxf(b2, 0.0f), 1.0f);
}

__device__ __forceinline__ int32_t ordered_from_float(float f) {
    int32_t i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}

struct MinMax {
    float lo = INFINITY, hi = -INFINITY;
    bool nan = false;               // a NaN was seen: the maximum is reported as NaN (ordered code above +inf), as torch.max would
    __device__ __forceinline__ void add(float v) {
        lo = fminf(lo, v); hi = fmaxf(hi, v);
        nan |= v != v;              // (tmo_cuda's range check, tone_mapping.py:43-45, fails exactly for NaN inputs: +-inf are clamped first)
    }
};

__device__ __forceinline__ void minmax_commit(MinMax mm, int32_t* out) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mm.lo = fminf(mm.lo, __shfl_xor_sync(0xffffffffu, mm.lo, o));
        mm.hi = fmaxf(mm.hi, __shfl_xor_sync(0xffffffffu, mm.hi, o));
    }
    __shared__ float s_lo[32], s_hi[32];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) { s_lo[warp] = mm.lo; s_hi[warp] = mm.hi; }
    const int any_nan = __syncthreads_or(mm.nan ? 1 : 0);
    if (warp == 0) {
        float lo = lane < nw ? s_lo[lane] : __int_as_float(0x7f800000);
        float hi = lane < nw ? s_hi[lane] : __int_as_float(0xff800000);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) {
            atomicMin(out, ordered_from_float(lo));
            atomicMax(out + 1, any_nan ? 0x7fc00000 : ordered_from_float(hi));
        }
    }
}

template <typename T> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ void load(const float* p, int64_t i, float* v) {
        float4 t = __ldcs(reinterpret_cast<const float4*>(p + i));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ float load1(const float* p, int64_t i) { return __ldcs(p + i); }
};
template <> struct Vec4<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, int64_t i, float* v) {
        uint2 t = __ldcs(reinterpret_cast<const uint2*>(p + i));
        v[0] = bf16_lo(t.x); v[1] = bf16_hi(t.x); v[2] = bf16_lo(t.y); v[3] = bf16_hi(t.y);
    }
    static __device__ __forceinline__ float load1(const __nv_bfloat16* p, int64_t i) {
        return __bfloat162float(p[i]);
    }
};

__device__ __forceinline__ void store4(float* p, int64_t i, const float* v) {
    __stcs(reinterpret_cast<float4*>(p + i), make_float4(v[0], v[1], v[2], v[3]));
}

// One "item" = VEC consecutive pixels of one image; CH channel planes are processed together so the gamut
// matrix sees r,g,b of the same pixel.  CH = 1 is the flat elementwise case.
//   plane_stride: elements between channel planes (H*W) ; img_stride = CH * plane_stride
//   INTERLEAVED (CH == 3 only): pixel-major [n_px, 3]; VEC pixels = 3*VEC consecutive floats.
template <typename T, int CH, int VEC, bool INTERLEAVED, bool BYTES>
__global__ void __launch_bounds__(256) hdr_kernel(const T* __restrict__ sdr, const T* __restrict__ gm,
                                                  float* __restrict__ hdr_out, float* __restrict__ tmo_out,
                                                  int32_t* __restrict__ minmax, ByteOuts bo, int64_t items_per_img, int64_t n_items,
                                                  int64_t plane_stride, HdrConsts c) {
    MinMax mm;
    const bool do_eq1 = c.flags & GMD_HDR_EQ1;
    const bool do_gamut = (c.flags & GMD_HDR_GAMUT) && CH == 3;
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < n_items;
         it += (int64_t)gridDim.x * blockDim.x) {
        float s[CH][VEC], g[CH][VEC], h[CH][VEC];
        int64_t off[CH];
        int64_t px0;  // index of this item's first pixel over the whole batch
        if (INTERLEAVED) {
            // 3*VEC contiguous floats, VEC == 4 -> three 16-byte vectors
            int64_t base = it * (3 * VEC);
            float sv[3 * VEC], gv[3 * VEC];
            if (VEC == 4) {
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    float t4[4];
                    Vec4<T>::load(sdr, base + 4 * q, t4);
#pragma unroll
                    for (int k = 0; k < 4; ++k) sv[4 * q + k] = t4[k];
                    if (do_eq1) {
                        Vec4<T>::load(gm, base + 4 * q, t4);
#pragma unroll
                        for (int k = 0; k < 4; ++k) gv[4 * q + k] = t4[k];
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < 3 * VEC; ++k) {
                    sv[k] = Vec4<T>::load1(sdr, base + k);
                    gv[k] = do_eq1 ? Vec4<T>::load1(gm, base + k) : 0.0f;
                }
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v)
#pragma unroll
                for (int ch = 0; ch < CH; ++ch) { s[ch][v] = sv[3 * v + ch]; g[ch][v] = gv[3 * v + ch]; }
            off[0] = base;
            px0 = it * VEC;
        } else {
            int64_t img = it / items_per_img;
            int64_t px = (it - img * items_per_img) * VEC;
            px0 = img * plane_stride + px;
#pragma unroll
            for (int ch = 0; ch < CH; ++ch) off[ch] = (img * CH + ch) * plane_stride + px;
#pragma unroll
            for (int ch = 0; ch < CH; ++ch) {
                if (VEC == 4) {
                    Vec4<T>::load(sdr, off[ch], s[ch]);
                    if (do_eq1) Vec4<T>::load(gm, off[ch], g[ch]);
                } else {
                    s[ch][0] = Vec4<T>::load1(sdr, off[ch]);
                    if (do_eq1) g[ch][0] = Vec4<T>::load1(gm, off[ch]);
                }
            }
        }
        // Eq.(1)
#pragma unroll
        for (int ch = 0; ch < CH; ++ch)
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                float x = s[ch][v];
                if (do_eq1) x = eq1(denorm(x, c.flags), denorm(g[ch][v], c.flags), c);
                h[ch][v] = x;
                if (minmax) mm.add(x);
            }
        // byte outputs for the host tail (generate_hdr.py:27-30,243-244): PNG-ready uint8 SDR / GM and Radiance RGBE
        if (BYTES && (bo.sdr_u8 || bo.gm_u8)) {
#pragma unroll
            for (int which = 0; which < 2; ++which) {
                uint8_t* dst = which ? bo.gm_u8 : bo.sdr_u8;
                if (!dst) continue;
                if (INTERLEAVED) {
                    uint32_t w[(3 * VEC + 3) / 4] = {};
#pragma unroll
                    for (int v = 0; v < VEC; ++v)
#pragma unroll
                        for (int ch = 0; ch < CH; ++ch) {
                            int k = 3 * v + ch;
                            w[k / 4] |= to_u8(denorm(which ? g[ch][v] : s[ch][v], c.flags)) << (8 * (k % 4));
                        }
                    if (VEC == 4) {
#pragma unroll
                        for (int q = 0; q < 3; ++q) reinterpret_cast<uint32_t*>(dst + off[0])[q] = w[q];
                    } else {
#pragma unroll
                        for (int k = 0; k < 3 * VEC; ++k) dst[off[0] + k] = (uint8_t)(w[k / 4] >> (8 * (k % 4)));
                    }
                } else {
#pragma unroll
                    for (int ch = 0; ch < CH; ++ch) {
                        uint32_t w = 0;
#pragma unroll
                        for (int v = 0; v < VEC; ++v) w |= to_u8(denorm(which ? g[ch][v] : s[ch][v], c.flags)) << (8 * v);
                        if (VEC == 4) *reinterpret_cast<uint32_t*>(dst + off[ch]) = w;
                        else dst[off[ch]] = (uint8_t)w;
                    }
                }
            }
        }
        if (BYTES && CH == 3 && bo.rgbe) {
            uint32_t w[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                w[v] = rgbe_pack(__fdiv_rn(h[0][v], c.rgbe_div), __fdiv_rn(h[CH > 1 ? 1 : 0][v], c.rgbe_div),
                                 __fdiv_rn(h[CH > 2 ? 2 : 0][v], c.rgbe_div), c.rgbe_min);
            if (VEC == 4) __stcs(reinterpret_cast<uint4*>(bo.rgbe + 4 * px0), make_uint4(w[0], w[1], w[2], w[3]));
            else reinterpret_cast<uint32_t*>(bo.rgbe)[px0] = w[0];
        }
        if (hdr_out) {
            if (INTERLEAVED) {
                float ov[3 * VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v)
#pragma unroll
                    for (int ch = 0; ch < CH; ++ch) ov[3 * v + ch] = h[ch][v];
                if (VEC == 4) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        float t4[4] = {ov[4 * q], ov[4 * q + 1], ov[4 * q + 2], ov[4 * q + 3]};
                        store4(hdr_out, off[0] + 4 * q, t4);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 3 * VEC; ++k) hdr_out[off[0] + k] = ov[k];
                }
            } else {
#pragma unroll
                for (int ch = 0; ch < CH; ++ch) {
                    if (VEC == 4) store4(hdr_out, off[ch], h[ch]);
                    else hdr_out[off[ch]] = h[ch][0];
                }
            }
        }
        if (tmo_out) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
#pragma unroll
                for (int ch = 0; ch < CH; ++ch) h[ch][v] = tmo_apply(h[ch][v], c);
                if (do_gamut) gamut709(h[0][v], h[CH > 1 ? 1 : 0][v], h[CH > 2 ? 2 : 0][v]);
            }
            if (INTERLEAVED) {
                float ov[3 * VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v)
#pragma unroll
                    for (int ch = 0; ch < CH; ++ch) ov[3 * v + ch] = h[ch][v];
                if (VEC == 4) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        float t4[4] = {ov[4 * q], ov[4 * q + 1], ov[4 * q + 2], ov[4 * q + 3]};
                        store4(tmo_out, off[0] + 4 * q, t4);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 3 * VEC; ++k) tmo_out[off[0] + k] = ov[k];
                }
            } else {
#pragma unroll
                for (int ch = 0; ch < CH; ++ch) {
                    if (VEC == 4) store4(tmo_out, off[ch], h[ch]);
                    else tmo_out[off[ch]] = h[ch][0];
                }
            }
        }
    }
    if (minmax) minmax_commit(mm, minmax);
}

__global__ void minmax_init_kernel(int32_t* mm) {
    mm[0] = 0x7f800000;                      // +inf (ordered encoding of a non-negative float is itself)
    mm[1] = (int32_t)0xff800000 ^ 0x7fffffff;  // -inf
}

template <typename T, int CH, int VEC, bool INTER>
int launch(const gmd_hdr_params* p, const HdrConsts& c, int64_t items_per_img, int64_t n_items, int64_t plane_stride,
           cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int64_t want = (n_items + 255) / 256;
    int64_t cap = (int64_t)sms * 8;   // 8 resident CTAs of 256 threads per SM; grid-stride over the rest
    int grid = (int)(want < cap ? (want > 0 ? want : 1) : cap);
    // the byte outputs cost ~16-35 registers: only the launches that ask for them pay (BYTES instantiation)
    ByteOuts bo{p->rgbe_out, p->sdr_u8_out, p->gm_u8_out};
    auto* sdr = static_cast<const T*>(p->sdr);
    auto* gm = static_cast<const T*>(p->gm);
    if (bo.rgbe || bo.sdr_u8 || bo.gm_u8)
        hdr_kernel<T, CH, VEC, INTER, true><<<grid, 256, 0, st>>>(sdr, gm, p->hdr_out, p->tmo_out, p->minmax, bo, items_per_img, n_items, plane_stride, c);
    else
        hdr_kernel<T, CH, VEC, INTER, false><<<grid, 256, 0, st>>>(sdr, gm, p->hdr_out, p->tmo_out, p->minmax, bo, items_per_img, n_items, plane_stride, c);
    count_launch(1);
    return check_launch("hdr_kernel");
}

template <typename T>
int dispatch(const gmd_hdr_params* p, const HdrConsts& c, cudaStream_t st) {
    auto aligned16 = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const size_t in_align = sizeof(T) == 4 ? 15 : 7;
    bool ptr_ok = (reinterpret_cast<uintptr_t>(p->sdr) & in_align) == 0 &&
                  (p->gm == nullptr || (reinterpret_cast<uintptr_t>(p->gm) & in_align) == 0) && aligned16(p->hdr_out) &&
                  aligned16(p->tmo_out) && aligned16(p->rgbe_out) && aligned16(p->sdr_u8_out) && aligned16(p->gm_u8_out);
    if (p->layout == GMD_LAYOUT_PLANAR3) {
        if (ptr_ok && p->n_px % 4 == 0)
            return launch<T, 3, 4, false>(p, c, p->n_px / 4, p->batch * (p->n_px / 4), p->n_px, st);
        return launch<T, 3, 1, false>(p, c, p->n_px, p->batch * p->n_px, p->n_px, st);
    }
    if (p->layout == GMD_LAYOUT_INTERLEAVED3) {
        if (ptr_ok && p->n_px % 4 == 0) return launch<T, 3, 4, true>(p, c, 0, p->n_px / 4, 0, st);
        return launch<T, 3, 1, true>(p, c, 0, p->n_px, 0, st);
    }
    // flat
    int64_t n = p->n_px;
    if (ptr_ok && n % 4 == 0) return launch<T, 1, 4, false>(p, c, n / 4, n / 4, n, st);
    return launch<T, 1, 1, false>(p, c, n, n, n, st);
}


// ---------------------------------------------------------------------------------------------
// Backward of the same chain (stage-1 training: scripts/stage1/train_vqgan_lora.py:1134-1141 differentiates
// apply_gm_to_sdr -> TMO -> gamut_compress through torch autograd).  One pass, everything recomputed from the inputs;
// clamp gradients follow torch.clamp (passed where min <= x <= max), pow follows torch.pow (2.2 * s^1.2).
// ---------------------------------------------------------------------------------------------
struct BwdArgs {
    const float* sdr; const float* gm; const float* grad_out; float* grad_sdr; float* grad_gm;
    int64_t n_px, batch, ch_stride, px_stride, img_stride;   // element (img, ch, px) at img*img_stride + ch*ch_stride + px*px_stride
    int ch;             // 1 (flat) or 3
    int wrt_tmo;        // grad_out is w.r.t. the TMO(+gamut) output (1) or w.r.t. the Eq.(1) output (0)
};

__device__ __forceinline__ float clamp_mask(float x, float lo, float hi) { return (x >= lo && x <= hi) ? 1.0f : 0.0f; }

__global__ void __launch_bounds__(256) hdr_bwd_kernel(BwdArgs a, HdrConsts c) {
    const bool do_eq1 = c.flags & GMD_HDR_EQ1;
    const bool do_gamut = (c.flags & GMD_HDR_GAMUT) && a.ch == 3 && a.wrt_tmo;
    const int64_t total = a.batch * a.n_px;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t img = i / a.n_px, px = i - img * a.n_px;
        const int64_t base = img * a.img_stride + px * a.px_stride;
        float g[3] = {0, 0, 0}, x[3], dlin_ds[3], gain[3], lin[3], m_hdr[3], m_s[3];
        for (int k = 0; k < a.ch; ++k) g[k] = a.grad_out[base + k * a.ch_stride];
        // ---- forward recompute up to the TMO input ----
        for (int k = 0; k < a.ch; ++k) {
            const float s_raw = a.sdr[base + k * a.ch_stride];
            if (do_eq1) {
                const float gmv = a.gm[base + k * a.ch_stride];
                const float s = fminf(fmaxf(s_raw, 0.0f), 1.0f);
                m_s[k] = clamp_mask(s_raw, 0.0f, 1.0f);
                lin[k] = s > 0.0f ? powf(s, 2.2f) : 0.0f;
                dlin_ds[k] = s > 0.0f ? 2.2f * powf(s, 1.2f) : 0.0f;
                gain[k] = 1.0f + gmv * c.qmax;
                float h = (lin[k] + c.eps) * gain[k] - c.eps;
                m_hdr[k] = 1.0f;
                if (c.flags & GMD_HDR_CLAMP_OUT) { m_hdr[k] = clamp_mask(h, 0.0f, c.hi); h = fminf(fmaxf(h, 0.0f), c.hi); }
                x[k] = h;
            } else {
                x[k] = s_raw; m_hdr[k] = 1.0f;
            }
        }
        // ---- backward through gamut and TMO ----
        if (a.wrt_tmo) {
            float t[3], dt[3];
            for (int k = 0; k < a.ch; ++k) {
                switch (c.tmo) {
                    case GMD_TMO_LINEAR: t[k] = x[k] * c.inv_hi; dt[k] = c.inv_hi; break;
                    case GMD_TMO_HARD_CLIP: t[k] = fminf(fmaxf(x[k], 0.0f), 1.0f); dt[k] = clamp_mask(x[k], 0.0f, 1.0f); break;
                    case GMD_TMO_MULOG: {
                        const float y = x[k] * c.inv_hi;
                        const float tt = log1pf(c.mu * y) * c.inv_log1p_mu;
                        t[k] = fminf(fmaxf(tt, 0.0f), 1.0f);
                        dt[k] = clamp_mask(tt, 0.0f, 1.0f) * c.mu * c.inv_hi * c.inv_log1p_mu / (1.0f + c.mu * y);
                        break;
                    }
                    case GMD_TMO_CUDA: {
                        const float y0 = x[k] * 0.1f, y = fminf(fmaxf(y0, 0.0f), 1.0f);
                        t[k] = log1pf(c.mu * y) * c.inv_log1p_mu;
                        dt[k] = clamp_mask(y0, 0.0f, 1.0f) * 0.1f * c.mu * c.inv_log1p_mu / (1.0f + c.mu * y);
                        break;
                    }
                    default: t[k] = x[k]; dt[k] = 1.0f; break;
                }
            }
            if (do_gamut) {
                const float M[3][3] = {{1.660491f, -0.587641f, -0.072850f}, {-0.124550f, 1.132900f, -0.008349f}, {-0.018151f, -0.100579f, 1.118730f}};
                float go[3];
                for (int r = 0; r < 3; ++r) {
                    const float o = M[r][0] * t[0] + M[r][1] * t[1] + M[r][2] * t[2];
                    go[r] = g[r] * clamp_mask(o, 0.0f, 1.0f);
                }
                for (int k = 0; k < 3; ++k) g[k] = M[0][k] * go[0] + M[1][k] * go[1] + M[2][k] * go[2];
            }
            for (int k = 0; k < a.ch; ++k) g[k] *= dt[k];
        }
        // ---- backward through Eq.(1) ----
        for (int k = 0; k < a.ch; ++k) {
            const int64_t off = base + k * a.ch_stride;
            if (do_eq1) {
                const float gh = g[k] * m_hdr[k];
                if (a.grad_gm) a.grad_gm[off] = gh * (lin[k] + c.eps) * c.qmax;
                if (a.grad_sdr) a.grad_sdr[off] = gh * gain[k] * dlin_ds[k] * m_s[k];
            } else if (a.grad_sdr) {
                a.grad_sdr[off] = g[k];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// RandomExposureAdjust (gm_diffusion/stage1/augmentations.py:24-73) as one elementwise pass:
//   inverse camera curve ((sigma*y) / (1 + sigma - y + 1e-8))^(1/n)  ->  uint16 discretisation (clamp(x*65535, 0, 65535).round() / 65535)
//   ->  exposure + display gamma  clamp(x*exposure, 0, 1)^(1/gamma).   Stage flags select sub-chains (the class's helper methods).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) exposure_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n, int stages, float inv_n, float sigma,
                                                       float exposure, float inv_gamma) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float x = __ldcs(src + i);
        if (stages & 1) x = powf(__fdiv_rn(__fmul_rn(sigma, x), __fadd_rn(__fsub_rn(__fadd_rn(1.0f, sigma), x), 1e-8f)), inv_n);
        if (stages & 2) x = __fdiv_rn(rintf(fminf(fmaxf(__fmul_rn(x, 65535.0f), 0.0f), 65535.0f)), 65535.0f);
        if (stages & 4) x = powf(fminf(fmaxf(__fmul_rn(x, exposure), 0.0f), 1.0f), inv_gamma);
        __stcs(dst + i, x);
    }
}

}  // namespace
}  // namespace gmd

extern "C" int gmd_hdr_reconstruct(const gmd_hdr_params* p, void* stream) {
    using namespace gmd;
    if (!p || !p->sdr) { set_last_error("gmd_hdr_reconstruct: null input"); return kErrInvalid; }
    if ((p->flags & GMD_HDR_EQ1) && !p->gm) { set_last_error("gmd_hdr_reconstruct: Eq.(1) needs a gain map"); return kErrInvalid; }
    if (!p->hdr_out && !p->tmo_out && !p->minmax && !p->rgbe_out && !p->sdr_u8_out && !p->gm_u8_out) {
        set_last_error("gmd_hdr_reconstruct: no output requested"); return kErrInvalid;
    }
    if (p->rgbe_out && p->layout == GMD_LAYOUT_FLAT) { set_last_error("gmd_hdr_reconstruct: RGBE output needs a 3-channel layout"); return kErrInvalid; }
    if (p->gm_u8_out && !p->gm) { set_last_error("gmd_hdr_reconstruct: gm_u8_out without a gain map"); return kErrInvalid; }
    if (p->n_px < 0 || p->batch < 0) { set_last_error("gmd_hdr_reconstruct: negative size"); return kErrInvalid; }
    if ((p->flags & GMD_HDR_GAMUT) && p->layout == GMD_LAYOUT_FLAT) {
        set_last_error("gmd_hdr_reconstruct: gamut compression needs a 3-channel layout"); return kErrInvalid;
    }
    if (p->tmo < GMD_TMO_NONE || p->tmo > GMD_TMO_CUDA) { set_last_error("gmd_hdr_reconstruct: unknown tmo %d", p->tmo); return kErrInvalid; }
    int64_t total = p->layout == GMD_LAYOUT_PLANAR3 ? p->n_px * p->batch : p->n_px;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (p->minmax) {
        minmax_init_kernel<<<1, 1, 0, st>>>(p->minmax);
        count_launch(1);
    }
    if (total == 0) return check_launch("hdr_reconstruct(empty)");
    HdrConsts c;
    c.qmax = p->qmax; c.eps = p->eps; c.hi = p->qmax + 1.0f;
    c.inv_hi = (float)(1.0 / ((double)p->qmax + 1.0));
    c.flags = p->flags; c.tmo = p->tmo;
    double mu = p->tmo == GMD_TMO_CUDA ? 5000.0 : (double)p->mu;
    c.mu = (float)mu;
    c.inv_log1p_mu = (float)(1.0 / log1p(mu));
    c.log2_hi = (float)log2((double)p->qmax + 1.0);
    c.rgbe_div = p->rgbe_div != 0.0f ? p->rgbe_div : 1.0f;
    c.rgbe_min = (float)1e-32;
    if ((double)c.rgbe_min < 1e-32) c.rgbe_min = nextafterf(c.rgbe_min, INFINITY);
    if (p->in_dtype == GMD_F32) return dispatch<float>(p, c, st);
    if (p->in_dtype == GMD_BF16) return dispatch<__nv_bfloat16>(p, c, st);
    set_last_error("gmd_hdr_reconstruct: unknown in_dtype %d", p->in_dtype);
    return kErrInvalid;
}

extern "C" float gmd_decode_ordered(int32_t v) {
    int32_t i = v >= 0 ? v : v ^ 0x7fffffff;
    float f;
    memcpy(&f, &i, sizeof(f));
    return f;
}

extern "C" int gmd_hdr_reconstruct_bwd(const gmd_hdr_params* p, const float* grad_out, float* grad_sdr, float* grad_gm, int32_t wrt_tmo, void* stream) {
    using namespace gmd;
    if (!p || !p->sdr || !grad_out) { set_last_error("gmd_hdr_reconstruct_bwd: null input"); return kErrInvalid; }
    if (!grad_sdr && !grad_gm) { set_last_error("gmd_hdr_reconstruct_bwd: no gradient requested"); return kErrInvalid; }
    if ((p->flags & GMD_HDR_EQ1) && !p->gm) { set_last_error("gmd_hdr_reconstruct_bwd: Eq.(1) needs the gain map"); return kErrInvalid; }
    if (!(p->flags & GMD_HDR_EQ1) && grad_gm) { set_last_error("gmd_hdr_reconstruct_bwd: grad_gm without Eq.(1)"); return kErrInvalid; }
    if (p->in_dtype != GMD_F32 || (p->flags & (GMD_HDR_DENORM | GMD_HDR_EXP_GAIN))) {
        set_last_error("gmd_hdr_reconstruct_bwd: fp32 inputs in [0,1] with the linear gain only (the training path)"); return kErrUnsupported;
    }
    if ((p->flags & GMD_HDR_GAMUT) && p->layout == GMD_LAYOUT_FLAT) { set_last_error("gmd_hdr_reconstruct_bwd: gamut needs a 3-channel layout"); return kErrInvalid; }
    if (p->tmo < GMD_TMO_NONE || p->tmo > GMD_TMO_CUDA) { set_last_error("gmd_hdr_reconstruct_bwd: unknown tmo %d", p->tmo); return kErrInvalid; }
    BwdArgs a{};
    a.sdr = static_cast<const float*>(p->sdr); a.gm = static_cast<const float*>(p->gm); a.grad_out = grad_out; a.grad_sdr = grad_sdr; a.grad_gm = grad_gm;
    a.wrt_tmo = wrt_tmo;
    if (p->layout == GMD_LAYOUT_PLANAR3) { a.ch = 3; a.n_px = p->n_px; a.batch = p->batch; a.ch_stride = p->n_px; a.px_stride = 1; a.img_stride = 3 * p->n_px; }
    else if (p->layout == GMD_LAYOUT_INTERLEAVED3) { a.ch = 3; a.n_px = p->n_px; a.batch = 1; a.ch_stride = 1; a.px_stride = 3; a.img_stride = 0; }
    else { a.ch = 1; a.n_px = p->n_px; a.batch = 1; a.ch_stride = 0; a.px_stride = 1; a.img_stride = 0; }
    const int64_t total = a.batch * a.n_px;
    if (total == 0) return kOk;
    HdrConsts c;
    c.qmax = p->qmax; c.eps = p->eps; c.hi = p->qmax + 1.0f;
    c.inv_hi = (float)(1.0 / ((double)p->qmax + 1.0));
    c.flags = p->flags; c.tmo = p->tmo;
    double mu = p->tmo == GMD_TMO_CUDA ? 5000.0 : (double)p->mu;
    c.mu = (float)mu;
    c.inv_log1p_mu = (float)(1.0 / log1p(mu));
    c.log2_hi = 0.0f; c.rgbe_div = 1.0f; c.rgbe_min = 0.0f;
    int64_t want = (total + 255) / 256;
    int grid = (int)(want < 148 * 8 ? want : 148 * 8);
    hdr_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, c);
    count_launch(1);
    return check_launch("hdr_bwd_kernel");
}

extern "C" int gmd_exposure_adjust(const float* src, float* dst, int64_t n, int32_t stages, double n_curve, float sigma, float exposure, double gamma, void* stream) {
    using namespace gmd;
    if (n == 0) return kOk;
    if (!src || !dst || n < 0) { set_last_error("gmd_exposure_adjust: bad arguments"); return kErrInvalid; }
    if ((stages & 1) && !(n_curve > 0.0)) { set_last_error("gmd_exposure_adjust: camera-curve exponent n must be positive"); return kErrInvalid; }
    if ((stages & 4) && !(gamma > 0.0)) { set_last_error("gmd_exposure_adjust: gamma must be positive"); return kErrInvalid; }
    int64_t want = (n + 255) / 256;
    int grid = (int)(want < 148 * 8 ? want : 148 * 8);
    // 1/n and 1/gamma as Python computes them: a double division rounded to fp32 when torch.pow receives the scalar
    exposure_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, n, stages, (float)(1.0 / n_curve), sigma, exposure,
                                                                          (float)(1.0 / gamma));
    count_launch(1);
    return check_launch("exposure_kernel");
}
