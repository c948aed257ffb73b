// Kernel (c): classifier-free-guidance combine + x0 prediction + PLMS / DDIM update + SDR/GM latent concat,
// one launch per branch per step.  One thread owns one latent pixel (4 channels = one float4) so every
// access is a 16-byte vector; outputs for the next UNet calls are written directly in the channel-padded
// bf16 pixel-major layout the convolution kernels read.
// Follows stable_diffusion_dual_unet.py:1045-1048,1063-1080,1093 and diffusers PNDMScheduler.step_plms /
// DDIMScheduler.step.
#include "common.cuh"
#include "../../include/gmd_b200.h"

namespace gmd {
void count_launch(int n);
namespace {

__device__ __forceinline__ float4 ld4(const float* p, int64_t i) { return __ldg(reinterpret_cast<const float4*>(p) + i); }
__device__ __forceinline__ void st4(float* p, int64_t i, float4 v) { reinterpret_cast<float4*>(p)[i] = v; }

// write one pixel row of `ch` bf16 channels: ch0-3 = a, ch4-7 = b, rest zero
__device__ __forceinline__ void store_row(void* dst, int64_t px, int ch, float4 a, float4 b) {
    uint4* row = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(dst) + px * ch);
    row[0] = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
    for (int k = 1; k < ch / 8; ++k) row[k] = make_uint4(0, 0, 0, 0);
}

// per-sample sums for guidance rescale (stable_diffusion_dual_unet.py:71-94): [B][4] = {S(c), S(c^2), S(g), S(g^2)}
__global__ void __launch_bounds__(256) rescale_stats_kernel(const float* __restrict__ eu, const float* __restrict__ ec,
                                                           float* __restrict__ stats, int64_t px_per_sample, float g) {
    int64_t b = blockIdx.y;
    float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < px_per_sample; i += (int64_t)gridDim.x * blockDim.x) {
        float4 c = ld4(ec, b * px_per_sample + i), u = ld4(eu, b * px_per_sample + i);
        float cv[4] = {c.x, c.y, c.z, c.w}, uv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float e = uv[k] + g * (cv[k] - uv[k]);
            s0 += cv[k]; s1 += cv[k] * cv[k]; s2 += e; s3 += e * e;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(~0u, s0, o); s1 += __shfl_xor_sync(~0u, s1, o);
        s2 += __shfl_xor_sync(~0u, s2, o); s3 += __shfl_xor_sync(~0u, s3, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(stats + b * 4 + 0, s0); atomicAdd(stats + b * 4 + 1, s1);
        atomicAdd(stats + b * 4 + 2, s2); atomicAdd(stats + b * 4 + 3, s3);
    }
}

// All arithmetic below uses explicit round-to-nearest intrinsics in the SAME association order as the torch
// expressions it replaces, so that nvcc cannot contract mul+add into FMA: the fused step is then bit-identical
// to the unfused fp32 chain (except the guidance-rescale reduction, whose summation order differs).
#define MUL(a, b) __fmul_rn(a, b)
#define ADD(a, b) __fadd_rn(a, b)
#define SUB(a, b) __fsub_rn(a, b)
#define DIV(a, b) __fdiv_rn(a, b)

struct F4 { float v[4]; };
// (plain loads, not ld.global.nc: x, the history ring and the stash are rewritten in place by the same launch — each thread reads its
// element before it writes it, but the read-only path is formally undefined for memory the kernel writes)
__device__ __forceinline__ F4 ld(const float* p, int64_t i) { float4 t = *(reinterpret_cast<const float4*>(p) + i); return {{t.x, t.y, t.z, t.w}}; }
__device__ __forceinline__ float4 f4(const F4& a) { return make_float4(a.v[0], a.v[1], a.v[2], a.v[3]); }

__global__ void __launch_bounds__(256) sched_kernel(gmd_sched_params p) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= p.n_px) return;
    // --- CFG combine: u + g * (c - u)   (dual_unet.py:1063-1065) ---
    F4 e = ld(p.eps_cond, i);
    if (p.eps_uncond) {
        F4 u = ld(p.eps_uncond, i);
        const float g = p.guidance_scale;
        const bool rescale = p.guidance_rescale > 0.0f && p.rescale_stats;
        float ratio = 1.0f;
        if (rescale) {
            // rescale_noise_cfg (dual_unet.py:71-94): unbiased std over the sample
            const float* s = p.rescale_stats + (i / p.px_per_sample) * 4;
            float n = (float)(p.px_per_sample * 4);
            float var_c = (s[1] - s[0] * s[0] / n) / (n - 1.0f);
            float var_g = (s[3] - s[2] * s[2] / n) / (n - 1.0f);
            ratio = DIV(sqrtf(var_c), sqrtf(var_g));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float c = e.v[k];
            float cfg = ADD(u.v[k], MUL(g, SUB(c, u.v[k])));
            if (rescale) {
                // phi * (cfg * (std_text/std_cfg)) + (1 - phi) * cfg
                float phi = p.guidance_rescale;
                cfg = ADD(MUL(phi, MUL(cfg, ratio)), MUL(SUB(1.0f, phi), cfg));
            }
            e.v[k] = cfg;
        }
    }
    F4 x = ld(p.x, i);
    // --- x0 = (x - sqrt(1-a_t) * eps) / sqrt(a_t) from the PRE-step latents and the loop's t (dual_unet.py:1072-1075) ---
    F4 x0;
#pragma unroll
    for (int k = 0; k < 4; ++k) x0.v[k] = DIV(SUB(x.v[k], MUL(p.sqrt_1m_alpha_t, e.v[k])), p.sqrt_alpha_t);
    if (p.x0_out) st4(p.x0_out, i, f4(x0));
    if (p.stash_out) st4(p.stash_out, i, f4(x));
    if (p.eps_out && p.mode != GMD_SCHED_DPMPP) st4(p.eps_out, i, f4(e));
    // --- scheduler update ---
    F4 xn;
    if (p.mode == GMD_SCHED_DPMPP) {
        // diffusers DPMSolverMultistepScheduler (dpmsolver++, midpoint, order 2): the history holds x0 predictions.
        //   m0 = (x - sigma_s * eps) / alpha_s                                   (convert_model_output)
        //   1st order: x_t = (sigma_t/sigma_s) x - (alpha_t (exp(-h) - 1)) m0
        //   2nd order: ... - 0.5 (alpha_t (exp(-h) - 1)) * ((1/r0) (m0 - m1))
        F4 m0, h0;
        if (p.plms_kind == 1) h0 = ld(p.hist[0], i);
        const float c_half = MUL(0.5f, p.c_num);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            m0.v[k] = DIV(SUB(x.v[k], MUL(p.ddim_sqrt_1m_alpha_t, e.v[k])), p.ddim_sqrt_alpha_t);
            float v = SUB(MUL(p.c_sample, x.v[k]), MUL(p.c_num, m0.v[k]));
            if (p.plms_kind == 1) v = SUB(v, MUL(c_half, MUL(p.c_denom, SUB(m0.v[k], h0.v[k]))));
            xn.v[k] = v;
        }
        if (p.eps_out) st4(p.eps_out, i, f4(m0));
    } else if (p.mode == GMD_SCHED_LINEAR) {
        // diffusers PNDMScheduler.step_plms: the multistep combination, written exactly as the reference expressions
        F4 h0, h1, h2, ep;
        if (p.hist[0]) h0 = ld(p.hist[0], i);
        if (p.hist[1]) h1 = ld(p.hist[1], i);
        if (p.hist[2]) h2 = ld(p.hist[2], i);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float m = e.v[k];
            switch (p.plms_kind) {
                case 1: m = DIV(ADD(m, h0.v[k]), 2.0f); break;                                          // (eps + ets[-1]) / 2
                case 2: m = DIV(SUB(MUL(3.0f, m), h0.v[k]), 2.0f); break;                                // (3 e1 - e2) / 2
                case 3: m = DIV(ADD(SUB(MUL(23.0f, m), MUL(16.0f, h0.v[k])), MUL(5.0f, h1.v[k])), 12.0f); break;
                case 4: m = MUL(0.041666666666666664f,
                                SUB(ADD(SUB(MUL(55.0f, m), MUL(59.0f, h0.v[k])), MUL(37.0f, h1.v[k])), MUL(9.0f, h2.v[k])));
                        break;
                default: break;
            }
            ep.v[k] = m;
        }
        F4 xs = p.use_stash ? ld(p.x_stash, i) : x;
        // sample_coeff * sample - (a_prev - a_t) * model_output / denom
#pragma unroll
        for (int k = 0; k < 4; ++k) xn.v[k] = SUB(MUL(p.c_sample, xs.v[k]), DIV(MUL(p.c_num, ep.v[k]), p.c_denom));
    } else if (p.mode == GMD_SCHED_DDPM) {
        // diffusers DDPMScheduler.step (epsilon, fixed_small): pred_x0; prev = c_x0 * pred_x0 + c_xt * x (+ sqrt(var) * noise if t > 0)
        F4 z;
        const bool noisy = p.noise && p.ddim_sigma != 0.0f;
        if (noisy) z = ld(p.noise, i);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float p0 = DIV(SUB(x.v[k], MUL(p.ddim_sqrt_1m_alpha_t, e.v[k])), p.ddim_sqrt_alpha_t);
            float v = ADD(MUL(p.ddim_sqrt_alpha_prev, p0), MUL(p.ddim_dir_coeff, x.v[k]));
            if (noisy) v = ADD(v, MUL(p.ddim_sigma, z.v[k]));
            xn.v[k] = v;
        }
    } else {
        // diffusers DDIMScheduler.step: pred_x0, direction, prev (+ sigma * noise)
        F4 z;
        const bool noisy = p.noise && p.ddim_sigma != 0.0f;
        if (noisy) z = ld(p.noise, i);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float p0 = DIV(SUB(x.v[k], MUL(p.ddim_sqrt_1m_alpha_t, e.v[k])), p.ddim_sqrt_alpha_t);
            float dir = MUL(p.ddim_dir_coeff, e.v[k]);
            float v = ADD(MUL(p.ddim_sqrt_alpha_prev, p0), dir);
            if (noisy) v = ADD(v, MUL(p.ddim_sigma, z.v[k]));
            xn.v[k] = v;
        }
    }
    st4(p.x_next, i, f4(xn));
    // --- fused layout outputs (replace torch.cat at dual_unet.py:1045,1080 / gm.py:1045) ---
    float4 zero = make_float4(0, 0, 0, 0);
    if (p.unet_in_next) {
        store_row(p.unet_in_next, i, p.unet_in_ch, f4(xn), zero);
        if (p.unet_in_dup == 2) store_row(p.unet_in_next, i + p.n_px, p.unet_in_ch, f4(xn), zero);
    }
    if (p.concat_out) {
        float4 lead = p.concat_lead ? ld4(p.concat_lead, i) : f4(x0);
        float4 tail = p.concat_self ? f4(xn) : (p.concat_tail ? ld4(p.concat_tail, i) : zero);
        store_row(p.concat_out, i, p.unet_in_ch, lead, tail);
        if (p.concat_dup == 2) store_row(p.concat_out, i + p.n_px, p.unet_in_ch, lead, tail);
    }
}

__global__ void __launch_bounds__(256) nchw_to_px_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t batch, int64_t hw) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= batch * hw) return;
    int64_t b = i / hw, p = i - b * hw;
    const float* s = src + b * 4 * hw + p;
    st4(dst, i, make_float4(s[0], s[hw], s[2 * hw], s[3 * hw]));
}
__global__ void __launch_bounds__(256) px_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t batch, int64_t hw) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= batch * hw) return;
    int64_t b = i / hw, p = i - b * hw;
    float4 v = ld4(src, i);
    float* d = dst + b * 4 * hw + p;
    d[0] = v.x; d[hw] = v.y; d[2 * hw] = v.z; d[3 * hw] = v.w;
}
__global__ void __launch_bounds__(256) pack_unet_input_kernel(const float* __restrict__ lead, const float* __restrict__ tail, void* dst, int64_t n_px, int ch) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_px) return;
    float4 zero = make_float4(0, 0, 0, 0);
    store_row(dst, i, ch, ld4(lead, i), tail ? ld4(tail, i) : zero);
}

// image NCHW fp32 [B,C,H,W], C <= 8  ->  NHWC bf16 [B,H,W,8] (zero-padded channels): the VAE encoder's conv_in operand
__global__ void __launch_bounds__(256) pack_image_kernel(const float* __restrict__ src, void* dst, int64_t batch, int64_t hw, int c) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= batch * hw) return;
    int64_t b = i / hw, p = i - b * hw;
    const float* s = src + b * c * hw + p;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = k < c ? __ldg(s + k * hw) : 0.0f;
    store_row(dst, i, 8, make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]));
}

// diffusers DiagonalGaussianDistribution (what `vae.encode(x).latent_dist` is, generate_hdr.py:208): moments [n_px, 8] =
// (mean[4], logvar[4]); logvar = clamp(logvar, -30, 20); std = exp(0.5 * logvar); sample = mean + std * noise; mode = mean.
__global__ void __launch_bounds__(256) vae_sample_kernel(const float* __restrict__ moments, const float* __restrict__ noise,
                                                         float* __restrict__ out, int64_t n_px, float scale) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_px) return;
    F4 mean = ld(moments, 2 * i), lv = ld(moments, 2 * i + 1), z;
    if (noise) z = ld(noise, i);
    F4 o;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float v = mean.v[k];
        if (noise) {
            float l = fminf(fmaxf(lv.v[k], -30.0f), 20.0f);
            v = ADD(v, MUL(expf(MUL(0.5f, l)), z.v[k]));
        }
        o.v[k] = scale == 1.0f ? v : MUL(v, scale);
    }
    st4(out, i, f4(o));
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace gmd

extern "C" int gmd_cfg_sched_step(const gmd_sched_params* p, void* stream) {
    using namespace gmd;
    if (!p) { set_last_error("gmd_cfg_sched_step: null params"); return kErrInvalid; }
    if (p->n_px == 0) return kOk;
    if (!p->eps_cond || !p->x || !p->x_next) { set_last_error("gmd_cfg_sched_step: null eps/x"); return kErrInvalid; }
    if (p->n_px < 0) { set_last_error("gmd_cfg_sched_step: negative n_px"); return kErrInvalid; }
    if (p->use_stash && !p->x_stash) { set_last_error("gmd_cfg_sched_step: use_stash without x_stash"); return kErrInvalid; }
    if ((p->unet_in_next || p->concat_out) && (p->unet_in_ch < 8 || p->unet_in_ch % 8)) {
        set_last_error("gmd_cfg_sched_step: unet_in_ch must be a multiple of 8 (got %d)", p->unet_in_ch); return kErrInvalid;
    }
    if (p->mode < GMD_SCHED_LINEAR || p->mode > GMD_SCHED_DPMPP) { set_last_error("gmd_cfg_sched_step: bad mode %d", p->mode); return kErrInvalid; }
    const void* ptrs[] = {p->eps_uncond, p->eps_cond, p->x, p->x_stash, p->hist[0], p->hist[1], p->hist[2], p->noise, p->x_next,
                          p->stash_out, p->eps_out, p->unet_in_next, p->concat_out, p->concat_tail, p->concat_lead, p->x0_out};
    for (const void* q : ptrs)
        if (!al16(q)) { set_last_error("gmd_cfg_sched_step: pointers must be 16-byte aligned"); return kErrInvalid; }
    if (p->mode == GMD_SCHED_LINEAR) {
        if (p->plms_kind < 0 || p->plms_kind > 4) { set_last_error("gmd_cfg_sched_step: plms_kind %d out of range", p->plms_kind); return kErrInvalid; }
        const int need = p->plms_kind == 0 ? 0 : p->plms_kind <= 2 ? 1 : p->plms_kind - 1;
        for (int k = 0; k < need; ++k)
            if (!p->hist[k]) { set_last_error("gmd_cfg_sched_step: plms_kind %d needs %d history tensors", p->plms_kind, need); return kErrInvalid; }
    }
    if (p->mode == GMD_SCHED_DPMPP) {
        if (p->plms_kind < 0 || p->plms_kind > 1) { set_last_error("gmd_cfg_sched_step: DPM++ order index %d out of range", p->plms_kind); return kErrInvalid; }
        if (p->plms_kind == 1 && !p->hist[0]) { set_last_error("gmd_cfg_sched_step: DPM++ 2nd-order step needs the previous x0 prediction"); return kErrInvalid; }
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (p->eps_uncond && p->guidance_rescale > 0.0f) {
        if (!p->rescale_stats || p->px_per_sample <= 0 || p->n_px % p->px_per_sample) {
            set_last_error("gmd_cfg_sched_step: guidance_rescale needs rescale_stats and px_per_sample"); return kErrInvalid;
        }
        int64_t B = p->n_px / p->px_per_sample;
        cudaMemsetAsync(p->rescale_stats, 0, sizeof(float) * 4 * B, st);
        dim3 grid((unsigned)((p->px_per_sample + 255) / 256 < 64 ? (p->px_per_sample + 255) / 256 : 64), (unsigned)B);
        rescale_stats_kernel<<<grid, 256, 0, st>>>(p->eps_uncond, p->eps_cond, p->rescale_stats, p->px_per_sample, p->guidance_scale);
        count_launch(1);
    }
    unsigned grid = (unsigned)((p->n_px + 255) / 256);
    sched_kernel<<<grid, 256, 0, st>>>(*p);
    count_launch(1);
    return check_launch("sched_kernel");
}

extern "C" int gmd_latents_nchw_to_px(const float* src, float* dst, int64_t batch, int64_t hw, void* stream) {
    using namespace gmd;
    if (batch * hw == 0) return kOk;
    if (!src || !dst || !al16(dst)) { set_last_error("gmd_latents_nchw_to_px: bad pointers"); return kErrInvalid; }
    nchw_to_px_kernel<<<(unsigned)((batch * hw + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, batch, hw);
    count_launch(1);
    return check_launch("nchw_to_px");
}
extern "C" int gmd_latents_px_to_nchw(const float* src, float* dst, int64_t batch, int64_t hw, void* stream) {
    using namespace gmd;
    if (batch * hw == 0) return kOk;
    if (!src || !dst || !al16(src)) { set_last_error("gmd_latents_px_to_nchw: bad pointers"); return kErrInvalid; }
    px_to_nchw_kernel<<<(unsigned)((batch * hw + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, batch, hw);
    count_launch(1);
    return check_launch("px_to_nchw");
}
extern "C" int gmd_pack_unet_input(const float* lead, const float* tail, void* dst, int64_t n_px, int32_t ch, void* stream) {
    using namespace gmd;
    if (n_px == 0) return kOk;
    if (!lead || !dst || ch < 8 || ch % 8 || !al16(lead) || !al16(tail) || !al16(dst)) { set_last_error("gmd_pack_unet_input: bad arguments"); return kErrInvalid; }
    pack_unet_input_kernel<<<(unsigned)((n_px + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(lead, tail, dst, n_px, ch);
    count_launch(1);
    return check_launch("pack_unet_input");
}

extern "C" int gmd_pack_image_nchw(const float* src, void* dst, int64_t batch, int64_t hw, int32_t channels, void* stream) {
    using namespace gmd;
    if (batch * hw == 0) return kOk;
    if (!src || !dst || channels < 1 || channels > 8 || !al16(dst)) { set_last_error("gmd_pack_image_nchw: bad arguments"); return kErrInvalid; }
    pack_image_kernel<<<(unsigned)((batch * hw + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, batch, hw, channels);
    count_launch(1);
    return check_launch("pack_image_nchw");
}
extern "C" int gmd_vae_sample(const float* moments, const float* noise, float* out, int64_t n_px, float scale, void* stream) {
    using namespace gmd;
    if (n_px == 0) return kOk;
    if (!moments || !out || !al16(moments) || !al16(noise) || !al16(out)) { set_last_error("gmd_vae_sample: bad arguments"); return kErrInvalid; }
    vae_sample_kernel<<<(unsigned)((n_px + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(moments, noise, out, n_px, scale);
    count_launch(1);
    return check_launch("vae_sample");
}
