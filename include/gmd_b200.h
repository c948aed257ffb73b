/*
 * gmd_b200.h — C-ABI of the B200-native GM-Diffusion Stage-3 hot path.
 *
 * The reference (Guanys-dar/GM-Diffusion) is pure Python and has no FFI of its own; the seam this
 * library slots into is the set of tensor ops its pipelines and stage-1 functions execute
 * (SURVEY.md §8b).  Every entry point below names the reference lines it replaces.  A maintainer
 * binds these with ctypes from
 *   gm_diffusion/stage1/tone_mapping.py            (kernel d)
 *   gm_diffusion/pipelines/stable_diffusion_*.py   (kernels a, b, c inside the denoising loop)
 *
 * Conventions
 *   - all pointers are DEVICE pointers owned by the caller (torch); `stream` is a cudaStream_t
 *     passed as void*; functions never allocate device memory and never synchronise.
 *   - return 0 on success, <0 on error: -1 invalid argument (binding raises ValueError),
 *     -2 CUDA error, -3 unsupported configuration (binding raises RuntimeError /
 *     NotImplementedError).  gmd_last_error() returns a thread-local message.
 *   - activations are bf16 NHWC ("pixel-major": [N, H, W, C]) or token-major [M, C]; latent /
 *     scheduler state is fp32 [N, H, W, 4]; images for kernel (d) are fp32 (or bf16) planar
 *     [B, 3, H, W] / interleaved [.., 3] / flat.
 */
#ifndef GMD_B200_H_
#define GMD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GMD_VERSION 201

int gmd_version(void);
const char* gmd_last_error(void);
/* number of kernels launched by this library in this process since load / last reset */
int64_t gmd_launch_count(void);
void gmd_reset_launch_count(void);
/* adjust the counter: launches recorded while a CUDA graph was being captured did not run (n < 0); replays of that graph do (n > 0) */
void gmd_add_launch_count(int64_t n);

/* ------------------------------------------------------------------------------------------ */
/* (d) Eq.(1) HDR reconstruction + TMO + gamut + min/max                                       */
/*     replaces gm_diffusion/stage1/tone_mapping.py:14-90 (apply_gm_to_sdr :60-71,             */
/*     linear_scale_tmo :14-18, hard_clip_tmo :21-26, fix_mulog_tmo :29-36, tmo_cuda :39-47,   */
/*     random_tmo_cuda :50-57, gamut_compress :74-90) and the de-normalise + numpy Eq.(1)      */
/*     tail of scripts/inference/generate_hdr.py:225-233,256-268.                              */
/* ------------------------------------------------------------------------------------------ */
enum gmd_layout { GMD_LAYOUT_FLAT = 0, GMD_LAYOUT_PLANAR3 = 1 /* [B,3,H*W] */, GMD_LAYOUT_INTERLEAVED3 = 2 /* [n_px,3] */ };
enum gmd_dtype { GMD_F32 = 0, GMD_BF16 = 1 };
enum gmd_tmo {
    GMD_TMO_NONE = 0,
    GMD_TMO_LINEAR = 1,    /* x / (qmax+1) */
    GMD_TMO_HARD_CLIP = 2, /* clamp(x, 0, 1) */
    GMD_TMO_MULOG = 3,     /* clamp(log1p(mu * x/(qmax+1)) / log1p(mu), 0, 1); fix_mulog: mu = 500 */
    GMD_TMO_CUDA = 4       /* y = clamp(x/10,0,1); log1p(5000 y)/log1p(5000)  (tmo_cuda) */
};
enum gmd_stage_flags {
    GMD_HDR_EQ1 = 1,          /* run Eq.(1); otherwise `sdr` is taken as the HDR input of the TMO stage */
    GMD_HDR_DENORM = 2,       /* inputs are in [-1,1]: x <- clamp(x/2 + 0.5, 0, 1) first (generate_hdr.py:227,232) */
    GMD_HDR_CLAMP_OUT = 4,    /* clamp(hdr, 0, qmax+1) (tone_mapping.py:71; numpy twins do not clamp) */
    GMD_HDR_GAMUT = 8,        /* BT.2020 -> BT.709 3x3 + clamp(0,1) after the TMO (needs a 3-channel layout) */
    GMD_HDR_EXP_GAIN = 16     /* extra, never default: gain = 2^(gm * log2(qmax+1)) instead of 1 + gm*qmax */
};

typedef struct gmd_hdr_params {
    const void* sdr;   /* SDR image (or HDR input when GMD_HDR_EQ1 is not set) */
    const void* gm;    /* gain map; may be NULL when GMD_HDR_EQ1 is not set */
    float* hdr_out;    /* optional: Eq.(1) output */
    float* tmo_out;    /* optional: TMO (+gamut) output */
    int32_t* minmax;   /* optional: 2 ordered-int32 slots {min,max} of the hdr values; decode with gmd_decode_ordered.  A NaN anywhere makes the
                          maximum decode to NaN (code 0x7fc00000, above +inf), as torch.max would report it */
    int64_t n_px;      /* PLANAR3: H*W per image plane; INTERLEAVED3: pixels; FLAT: elements */
    int64_t batch;     /* PLANAR3: number of images (planes = 3*batch); otherwise 1 */
    int32_t layout;    /* enum gmd_layout */
    int32_t in_dtype;  /* enum gmd_dtype */
    int32_t flags;     /* enum gmd_stage_flags */
    int32_t tmo;       /* enum gmd_tmo */
    float qmax;
    float eps;
    float mu;          /* GMD_TMO_MULOG only */
    /* Host-tail byte outputs (all optional; scripts/inference/generate_hdr.py:27-30,243-244):                                 */
    float rgbe_div;      /* RGBE encodes hdr / rgbe_div — save_hdr_image's `apply_HDR / (qmax+1)`; 0 means 1                 */
    uint8_t* rgbe_out;   /* [pixels,4] Radiance R,G,B,E of the (pre-TMO) HDR value, quantised as cv2.imwrite("x.hdr") does    */
                         /* (Ward float2rgbe); needs PLANAR3 or INTERLEAVED3; pixel order = batch-major scan order            */
    uint8_t* sdr_u8_out; /* trunc(clamp01(sdr) * 255), same element order as `sdr` (after DENORM when that flag is set)       */
    uint8_t* gm_u8_out;  /* same for the gain map                                                                              */
} gmd_hdr_params;

int gmd_hdr_reconstruct(const gmd_hdr_params* p, void* stream);
/* Backward of the same chain for stage-1 training (scripts/stage1/train_vqgan_lora.py:1134-1141 back-propagates through
   apply_gm_to_sdr -> TMO -> gamut_compress, tone_mapping.py:14-90, with torch autograd).  `p` describes the forward call (inputs,
   layout, flags, tmo, qmax, eps, mu; fp32 inputs, no DENORM / EXP_GAIN); grad_out has the forward output's layout and is the
   gradient w.r.t. the TMO(+gamut) output (wrt_tmo = 1) or w.r.t. the Eq.(1) output (wrt_tmo = 0); grad_sdr / grad_gm are optional. */
/* RandomExposureAdjust (gm_diffusion/stage1/augmentations.py:24-73), fp32 elementwise: stages bit 0 = inverse camera curve
   ((sigma*y)/(1+sigma-y+1e-8))^(1/n), bit 1 = uint16 discretisation, bit 2 = clamp(x*exposure,0,1)^(1/gamma).
   n_curve and gamma are doubles: the exponents 1/n and 1/gamma are formed in double like the Python expressions and rounded once. */
int gmd_exposure_adjust(const float* src, float* dst, int64_t n, int32_t stages, double n_curve, float sigma, float exposure, double gamma, void* stream);
int gmd_hdr_reconstruct_bwd(const gmd_hdr_params* p, const float* grad_out, float* grad_sdr, float* grad_gm, int32_t wrt_tmo, void* stream);
float gmd_decode_ordered(int32_t v);

/* ------------------------------------------------------------------------------------------ */
/* (c) CFG combine + x0 prediction + PLMS/DDIM update + SDR/GM concat, one launch per branch   */
/*     replaces gm_diffusion/pipelines/stable_diffusion_dual_unet.py:1045-1048,1063-1080,1093  */
/*     and stable_diffusion_gm.py:1045-1048,1062-1071 plus diffusers PNDMScheduler.step_plms / */
/*     DDIMScheduler.step (called at :1077/:1093).                                             */
/* ------------------------------------------------------------------------------------------ */
enum gmd_sched_mode {
    GMD_SCHED_LINEAR = 0, /* PLMS: x' = c_sample*x_src - c_num*eps'/c_denom */
    GMD_SCHED_DDIM = 1,   /* x' = sqrt(a_prev)*x0 + dir_coeff*eps (+ sigma*noise) */
    GMD_SCHED_DDPM = 2,   /* ancestral: x' = c_x0*x0 + c_xt*x (+ sigma*noise); the scheduler every reference CLI passes (generate_hdr.py:162-176) */
    GMD_SCHED_DPMPP = 3   /* DPM-Solver++(2M, midpoint), swapped in by scripts/inference/experiments/formal_improved.py:195.
                             m0 = (x - ddim_sqrt_1m_alpha_t*eps)/ddim_sqrt_alpha_t (alpha_s, sigma_s of the current sigma) is written
                             to eps_out (the history holds x0 predictions); plms_kind 0: x' = c_sample*x - c_num*m0;
                             plms_kind 1: ... - 0.5*c_num*(c_denom*(m0 - hist[0])), c_denom = 1/r0 */
};
/* PLMS multistep combination eps' (diffusers PNDMScheduler.step_plms):
 *   0: eps   1: (eps + h0)/2   2: (3 eps - h0)/2   3: (23 eps - 16 h0 + 5 h1)/12   4: (55 eps - 59 h0 + 37 h1 - 9 h2)/24 */

typedef struct gmd_sched_params {
    /* model outputs, fp32 [n_px, 4] (pixel-major) */
    const float* eps_uncond; /* NULL when CFG is off: eps = eps_cond */
    const float* eps_cond;
    /* state, fp32 [n_px, 4] */
    const float* x;          /* current latents */
    const float* x_stash;    /* PLMS: sample stashed at counter 0; used instead of x when use_stash */
    const float* hist[3];    /* previous eps (most recent first); NULL where weight is 0 */
    const float* noise;      /* optional pre-drawn variance noise (DDIM eta>0 / DDPM), fp32 [n_px,4] */
    float* x_next;           /* updated latents (may alias x) */
    float* stash_out;        /* optional: copy of x (PLMS counter 0) */
    float* eps_out;          /* optional: post-CFG eps stored for the history ring */
    /* fused layout outputs, bf16 [n_px, unet_in_ch] pixel-major, channel-padded with zeros */
    void* unet_in_next;      /* optional: ch0-3 <- x_next (next SDR-UNet input; CFG dup is implicit) */
    void* concat_out;        /* optional: ch0-3 <- x0 (or `concat_lead` if given), ch4-7 <- concat_tail */
    const float* concat_tail;/* fp32 [n_px,4]: the other branch's current latents (gm_latents) */
    const float* concat_lead;/* optional fp32 [n_px,4]: fixed leading channels (single pipeline: sdr_latent) */
    float* x0_out;           /* optional fp32 [n_px,4]: x0 prediction */
    int64_t n_px;            /* B*h*w */
    int64_t px_per_sample;   /* h*w, for per-sample guidance rescale */
    int32_t unet_in_ch;      /* channel count of unet_in_next / concat_out rows (multiple of 8, >= 8) */
    int32_t unet_in_dup;     /* 1 or 2: write unet_in_next twice (rows i and i + n_px) = the CFG batch duplication of dual_unet.py:1045 */
    int32_t concat_dup;      /* 1 or 2: same duplication for concat_out (single pipeline under CFG, gm.py:1045-1047) */
    int32_t concat_self;     /* 1: concat_out ch4-7 <- x_next of THIS branch (single pipeline) instead of concat_tail */
    int32_t mode;            /* enum gmd_sched_mode */
    int32_t use_stash;
    float guidance_scale;
    float guidance_rescale;  /* phi; > 0 needs rescale_stats */
    float* rescale_stats;    /* workspace fp32 [B, 4]: sums for std(eps_cond), std(eps_cfg) */
    float sqrt_alpha_t;      /* x0 = (x - sqrt_1m_alpha_t * eps) / sqrt_alpha_t (dual_unet.py:1072-1075) */
    float sqrt_1m_alpha_t;
    int32_t plms_kind;       /* LINEAR mode: which multistep combination (see above) */
    float c_sample;          /* LINEAR mode: (a_prev/a_t)^0.5 */
    float c_num;             /*              a_prev - a_t */
    float c_denom;           /*              a_t*(1-a_prev)^0.5 + (a_t*(1-a_t)*a_prev)^0.5 */
    float ddim_sqrt_alpha_t, ddim_sqrt_1m_alpha_t, ddim_sqrt_alpha_prev, ddim_dir_coeff, ddim_sigma; /* DDIM mode; DDPM mode reuses
        them as sqrt(a_t), sqrt(1-a_t), c_x0, c_xt, sigma */
} gmd_sched_params;

int gmd_cfg_sched_step(const gmd_sched_params* p, void* stream);

/* latents fp32 NCHW [B,4,h,w] <-> pixel-major fp32 [B,h,w,4] (pipeline entry / exit) */
int gmd_latents_nchw_to_px(const float* src, float* dst, int64_t batch, int64_t hw, void* stream);
int gmd_latents_px_to_nchw(const float* src, float* dst, int64_t batch, int64_t hw, void* stream);
/* fp32 [n_px,4] -> bf16 [n_px, ch] zero-padded rows, optional second source in ch4-7 */
int gmd_pack_unet_input(const float* lead, const float* tail, void* dst, int64_t n_px, int32_t ch, void* stream);
/* SDR->HDR entry (`pipeline.vae.encode(sdr_image).latent_dist.sample()`, scripts/inference/generate_hdr.py:208):
   image NCHW fp32 [batch, channels<=8, hw] -> NHWC bf16 [batch, hw, 8] (zero-padded), the encoder conv_in operand */
int gmd_pack_image_nchw(const float* src, void* dst, int64_t batch, int64_t hw, int32_t channels, void* stream);
/* diffusers DiagonalGaussianDistribution on the encoder moments fp32 [n_px, 8] = (mean[4], logvar[4]):
   out[n_px,4] = scale * (mean + exp(0.5*clamp(logvar,-30,20)) * noise); noise NULL -> mode() (= mean) */
int gmd_vae_sample(const float* moments, const float* noise, float* out, int64_t n_px, float scale, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* (b) convolutions / linears as tcgen05 implicit GEMM; GroupNorm+SiLU; LayerNorm               */
/*     replaces the diffusers UNet2DConditionModel / AutoencoderKL ops reached from            */
/*     stable_diffusion_dual_unet.py:1052-1060,1083-1092 and generate_hdr.py:225-233.          */
/* ------------------------------------------------------------------------------------------ */
enum gmd_epilogue_flags {
    GMD_EPI_BIAS = 1,       /* + bias[n] (fp32) */
    GMD_EPI_ROW_BIAS = 2,   /* + row_bias[sample(m), n] (fp32; the time-embedding projection) */
    GMD_EPI_RESIDUAL = 4,   /* + residual[m, n] (bf16, ld = ldr) */
    GMD_EPI_GEGLU = 8,      /* out[m, j] = (acc[m, j] + b[j]) * gelu(acc[m, j + N/2] + b[j + N/2]); weight rows interleaved per tile */
    GMD_EPI_OUT_F32 = 16,   /* write fp32 instead of bf16 */
    GMD_EPI_SCALE = 32,     /* acc *= alpha before everything else */
    GMD_EPI_RESIDUAL_F32 = 64, /* the residual tensor is fp32 (the transformer token stream is kept in fp32) */
    GMD_CONV_PAD_END = 256     /* gmd_conv_fwd, stride 2 only: no leading pad, one zero row/column appended at the bottom/right —
                                  diffusers Downsample2D(padding=0) + F.pad(x,(0,1,0,1)) of the AutoencoderKL encoder
                                  (`pipeline.vae.encode`, scripts/inference/generate_hdr.py:208) */
};

typedef struct gmd_gemm_params {
    /* out[M, N] = A[M, K] * W[N, K]^T  (+ epilogue).  A, W bf16 K-contiguous. */
    const void* a;  int64_t lda;        /* row stride in elements */
    const void* w;  int64_t ldw;
    void* out;      int64_t ldo;
    const float* bias;
    const float* row_bias; int64_t ld_row_bias; int64_t rows_per_sample;
    const void* residual;  int64_t ldr;
    int64_t M, N, K;
    int64_t batch;                       /* batched: strides in elements between problems */
    int64_t stride_a, stride_w, stride_o;
    int32_t flags;
    float alpha;
    void* workspace;                     /* optional fp32 scratch for split-K (few output tiles, long K); NULL disables it */
    int64_t workspace_bytes;
    int32_t w_tiled;                     /* 0: w is [N, K] row-major.  T > 0: w is pre-tiled [ceil(N/T)][ceil(K/64)][T][64] (zero padded), T = the
                                          * N tile the kernel uses for this N (160 if N%160==0, else 128 if N%128==0 or N>128, else 64 / 32):
                                          * every operand tile is then one contiguous DRAM read instead of T strided 128-byte rows.
                                          * 1000 + T: the tiles are additionally stored as the SWIZZLE_128B shared-memory image (16-byte chunk c of
                                          * row r at chunk c ^ (r & 7)) and fetched with ONE 1-D bulk copy per tile instead of T tensor rows */
    void* gn_sums;                       /* optional: GroupNorm statistics of the OUTPUT accumulated by the epilogue (see below); NULL = none */
    int64_t gn_rows_per_sample;          /* with gn_sums: output rows of one sample (a multiple of 128) */
    /* LayerNorm folded into the GEMMs on either side of it (BasicTransformerBlock norm1 -> to_q/k/v, norm2 -> to_q).  The GEMM that
     * writes the fp32 token stream x (PRODUCER) also emits a bf16 copy of x and each row's sum / sum of squares; the projection behind
     * the LayerNorm (CONSUMER) multiplies the bf16 copy with W' = W * diag(gamma) and normalises in its epilogue:
     *   out[r, n] = rstd_r * (acc[r, n] - mean_r * c[n]) + bias[n],   c[n] = sum_k W'[n, k],   bias = b + W * beta
     * — the LayerNorm kernel (a read of the fp32 stream and a write of its bf16 image) disappears.  Both sides need full N tiles,
     * batch == 1, no GEGLU and no split-K (else GMD_ERR_UNSUPPORTED). */
    void* ln_out_sums;                   /* PRODUCER: int64 [M, 2] (sum, sum of squares of output row r in 2^-24 fixed point), ADDED to: zeroed by the caller.  NULL = none */
    void* ln_out_copy;                   /* PRODUCER: bf16 [M, N] copy of the output, row stride N */
    const void* ln_in_sums;              /* CONSUMER: the producer's statistics of the rows of A (K = the producer's N).  NULL = none */
    const float* ln_in_c;                /* CONSUMER: fp32 [N] */
    float ln_eps;                        /* CONSUMER */
} gmd_gemm_params;

int gmd_gemm_fwd(const gmd_gemm_params* p, void* stream);

typedef struct gmd_conv_params {
    /* 3x3 (pad 1, stride 1 or 2) or 1x1 convolution, NHWC bf16, weights [Cout, R*S*(C0+C1)] bf16
     * with k = (r*S + s)*(C0+C1) + c.  Two sources implement the up-block skip concat without a copy. */
    const void* x0; int32_t C0;
    const void* x1; int32_t C1;          /* optional second source (channels appended after x0's) */
    const void* w;
    void* out;                           /* [N, Ho, Wo, Cout] bf16 (or fp32 with GMD_EPI_OUT_F32) */
    const float* bias;
    const float* row_bias; int64_t ld_row_bias;   /* [N, Cout] time-embedding projection */
    const void* residual;                /* [N, Ho, Wo, Cout] bf16 */
    int32_t N, H, W;                     /* input spatial size */
    int32_t Cout;                        /* logical output channels written */
    int32_t Cout_pad;                    /* rows of w (multiple of the N tile; >= Cout) */
    int32_t ksize;                       /* 1 or 3 */
    int32_t stride;                      /* 1 or 2 */
    int32_t upsample;                    /* 1: nearest-2x upsample of the input folded into the gather */
    int32_t flags;
    void* workspace;                     /* optional fp32 scratch for split-K (8x8-resolution layers); NULL disables it */
    int64_t workspace_bytes;
    int32_t w_tiled;                     /* as gmd_gemm_params.w_tiled; k blocks ordered (tap, 64-channel chunk of x0 then x1), channels zero padded */
    void* gn_sums;                       /* optional: GroupNorm statistics of the OUTPUT accumulated by the epilogue (see below); NULL = none */
} gmd_conv_params;

int gmd_conv_fwd(const gmd_conv_params* p, void* stream);

/* ---- GroupNorm fused into the producing convolution / GEMM (north_star (b); replaces the statistics pass of gmd_groupnorm_silu) ----
 * The epilogue of gmd_conv_fwd / gmd_gemm_fwd, which holds every output value in registers anyway, also accumulates per (sample,
 * channel pair) the sum and the sum of squares of its output into `gn_sums` [N][C/2][2] int64: each warp transpose-reduces its 32
 * rows x 32 columns with shuffles and adds 32 totals as FIXED-POINT integers (units of 2^-24) with 64-bit atomics.  Integer addition
 * is associative, so the result is bit-reproducible and independent of tile order and batch size — no partial buffers, no fold pass.
 * The caller zeroes `gn_sums` beforehand.  Channel pairs serve any consumer grouping (the up-block GroupNorms normalise a concat of
 * two producers whose 32 groups straddle the sources).  gmd_groupnorm_apply is then ONE streaming pass over the activation:
 * scale / shift (+ SiLU) with statistics formed from the sums of its one or two sources.
 * The *_gn_sums_ok functions say whether a call can emit statistics (no: split-K convolutions, ragged tiles, samples smaller than a
 * 128-row tile, GEGLU / batched GEMMs); the caller then keeps gmd_groupnorm_silu for that tensor. */
int gmd_conv_gn_sums_ok(const gmd_conv_params* p);
int gmd_gemm_gn_sums_ok(const gmd_gemm_params* p, int64_t rows_per_sample);
int gmd_groupnorm_apply(const void* x0, int32_t C0, const void* sums0, const void* x1, int32_t C1, const void* sums1,
                        const float* gamma, const float* beta, void* out, int32_t N, int32_t HW, int32_t groups, float eps,
                        int32_t apply_silu, int32_t in_dtype, void* stream);

/* GroupNorm(+SiLU) over NHWC with an optional second (concatenated, bf16) source; x0 is bf16 or fp32 (in_dtype =
 * enum gmd_dtype), output bf16.  Deterministic fixed-order reductions. */
int gmd_groupnorm_silu(const void* x0, int32_t C0, const void* x1, int32_t C1, const float* gamma, const float* beta,
                       void* out, int32_t N, int32_t HW, int32_t groups, float eps, int32_t apply_silu, int32_t in_dtype,
                       float* stats_ws /* workspace of 1024 + N*34*groups*2 4-byte words (N <= 1024); the first 1024 (arrival counters) ZEROED before first use */, void* stream);
/* LayerNorm over the last dim of token-major [M, C] (bf16 or fp32 in, bf16 out). */
int gmd_layernorm(const void* x, const float* gamma, const float* beta, void* out, int64_t M, int32_t C, float eps,
                  int32_t in_dtype, void* stream);
/* sinusoidal timestep embedding (flip_sin_to_cos, shift 0) -> bf16 [B, dim] */
int gmd_timestep_embedding(float t, void* out, int32_t B, int32_t dim, void* stream);
/* y = silu(x) elementwise bf16 */
int gmd_silu(const void* x, void* out, int64_t n, void* stream);
/* row softmax over bf16 [M, N] with scale (VAE mid-block attention) */
int gmd_softmax_rows(const void* x, void* out, int64_t M, int64_t N, float scale, void* stream);
/* as gmd_softmax_rows with columns >= n_valid (key padding) and, for causal_period > 0, columns > (row % causal_period) masked to 0:
 * the causal text attention of the CLIP text encoder (one causal_period x causal_period score matrix per head) */
int gmd_softmax_rows_masked(const void* x, void* out, int64_t M, int64_t N, float scale, int32_t n_valid, int32_t causal_period, void* stream);

/* ------------------------------------------------------------------------------------------ */
/* (a) attention: softmax(Q K^T * scale) V, tcgen05 + TMEM + TMA, streaming softmax            */
/*     replaces F.scaled_dot_product_attention inside diffusers AttnProcessor2_0 (self- and    */
/*     text cross-attention of the UNets called at stable_diffusion_dual_unet.py:1052,1083).   */
/*     d in {40, 80, 160}; strides in elements, multiples of 8; pointers 16-byte aligned.      */
/*     Dispatch (no change of results, only of kernel): heads contiguous within a token row    */
/*     (stride_h == d) -> K / V tiles fetched through dense tensor maps; 64 < Nk <= 80 with    */
/*     d in {40, 80} -> the persistent text cross-attention kernel; Nk <= 128 otherwise -> the */
/*     short configuration; anything else -> the streaming kernel.                             */
/* ------------------------------------------------------------------------------------------ */
typedef struct gmd_attn_params {
    /* q: [B, Nq, H, d] view with element strides; k, v: [B, Nk, H, d]; o: [B, Nq, H*d] bf16 */
    const void* q; int64_t q_stride_b, q_stride_n, q_stride_h;
    const void* k; int64_t k_stride_b, k_stride_n, k_stride_h;
    const void* v; int64_t v_stride_b, v_stride_n, v_stride_h;
    void* o;       int64_t o_stride_b, o_stride_n, o_stride_h;
    int32_t B, H, Nq, Nk, d;
    float scale;
} gmd_attn_params;

int gmd_attn_fwd(const gmd_attn_params* p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GMD_B200_H_ */
