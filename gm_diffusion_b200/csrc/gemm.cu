// Kernel family (b): bf16 GEMM and implicit-GEMM convolution on the 5th-generation tensor cores.
//
//   out[M, N] = A[M, K] * W[N, K]^T (+ fused epilogue)
//
// One CTA computes a 128 x BN output tile.  Warp 0 is the TMA producer, warp 1 issues tcgen05.mma
// (one elected thread) into a TMEM accumulator, warps 2-5 are the epilogue (TMEM -> registers ->
// bias / time-embedding / residual / GEGLU -> global).  Operands are staged by TMA into a STAGES-deep
// shared-memory ring in the canonical K-major SWIZZLE_128B layout (64 bf16 = 128 B per row).
//
// Convolution (3x3 pad 1, stride 1/2, optional folded nearest-2x upsample; NHWC bf16) is the same
// main loop: the 128 rows of an M tile are a (bn x bh x bw) box of output pixels, and for every filter
// tap the A tile is ONE 4-D TMA box load at the shifted coordinates; TMA's out-of-bounds zero fill is
// the convolution padding.  A second source tensor supplies the appended channels of the up-block
// skip connection, so torch.cat([hidden, skip]) is never materialised.
#include "common.cuh"
#include "../../include/gmd_b200.h"

namespace gmd {
void count_launch(int n);
namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int A_TILE_BYTES = BM * BK * 2;

struct KernelArgs {
    int mode;       // 0 = GEMM (3-D maps: k, row, batch), 1 = conv (4-D A maps: c, w, h, n)
    int num_kb;     // number of 64-wide K blocks
    // GEMM
    int kb_src0;    // K blocks taken from A source 0 (the rest from source 1)
    // conv
    int ks, stride, upsample;
    int chunks0, chunks1;  // 64-channel blocks per tap from source 0 / 1
    int ctot;              // C0 + C1 (weight K pitch per tap)
    int bw, bh, bn;        // pixel box of one M tile (bw*bh*bn == 128)
    int tiles_w, tiles_h;  // tile grid over the (Wg x Hg) iteration grid
    int Wg, Hg, Ng;        // iteration grid (output grid; low-res grid when upsample)
    int Wo, Ho;            // output spatial dims
    // epilogue
    int64_t M;             // GEMM rows
    int N_out;             // valid output columns (after GEGLU halving)
    void* out; int64_t ldo;
    const float* bias;
    const float* row_bias; int64_t ld_row_bias; int64_t rows_per_sample;
    const __nv_bfloat16* residual; int64_t ldr;
    int64_t batch_stride_o, batch_stride_r;
    int flags;
    float alpha;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
            const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_a3,
            const __grid_constant__ CUtensorMap map_w, const KernelArgs args) {
    constexpr int B_TILE_BYTES = BN * BK * 2;
    constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
    constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : BN <= 256 ? 256 : 512;
    constexpr uint32_t IDESC = umma_idesc_bf16(BM, BN, false, false);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // ---- tile coordinates ----
    const int n0 = blockIdx.y * BN;  // first weight row of this tile
    int m0 = 0, tw0 = 0, th0 = 0, tn0 = 0;
    if (args.mode == 0) {
        m0 = blockIdx.x * BM;
    } else {
        int t = blockIdx.x;
        int tw = t % args.tiles_w; t /= args.tiles_w;
        int th = t % args.tiles_h; t /= args.tiles_h;
        tw0 = tw * args.bw; th0 = th * args.bh; tn0 = t * args.bn;
    }
    const int zb = blockIdx.z;  // GEMM: batch index; conv+upsample: output parity class

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_a0);
        tma_prefetch_desc(&map_w);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(acc_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (elect_one()) {
            const int chunks_per_tap = args.chunks0 + args.chunks1;
            for (int kb = 0; kb < args.num_kb; ++kb) {
                const int stage = kb % STAGES;
                const uint32_t phase = (kb / STAGES) & 1;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + stage * STAGE_BYTES;
                uint8_t* sb = sa + A_TILE_BYTES;
                mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
                if (args.mode == 0) {
                    if (kb < args.kb_src0) tma_load_3d(sa, &map_a0, &full_bar[stage], kb * BK, m0, zb);
                    else tma_load_3d(sa, &map_a1, &full_bar[stage], (kb - args.kb_src0) * BK, m0, zb);
                    tma_load_3d(sb, &map_w, &full_bar[stage], kb * BK, n0, zb);
                } else {
                    const int tap = kb / chunks_per_tap;
                    const int cc = kb - tap * chunks_per_tap;
                    const int r = tap / args.ks, s = tap - r * args.ks;
                    const int pad = args.ks >> 1;
                    int dh, dw;
                    const CUtensorMap* map;
                    int c;
                    if (args.stride == 2) {
                        // input row 2*oh - 1 + r: r=0 -> odd plane, oh-1; r=1 -> even plane, oh; r=2 -> odd plane, oh
                        const int ph = (r == 1) ? 0 : 1, pw = (s == 1) ? 0 : 1;
                        dh = (r == 0) ? -1 : 0; dw = (s == 0) ? -1 : 0;
                        const int sel = ph * 2 + pw;
                        map = sel == 0 ? &map_a0 : sel == 1 ? &map_a1 : sel == 2 ? &map_a2 : &map_a3;
                        c = cc * BK;
                    } else {
                        if (args.upsample) {
                            const int py = zb >> 1, px = zb & 1;
                            dh = py == 0 ? (r == 0 ? -1 : 0) : (r == 2 ? 1 : 0);
                            dw = px == 0 ? (s == 0 ? -1 : 0) : (s == 2 ? 1 : 0);
                        } else {
                            dh = r - pad; dw = s - pad;
                        }
                        if (cc < args.chunks0) { map = &map_a0; c = cc * BK; }
                        else { map = &map_a1; c = (cc - args.chunks0) * BK; }
                    }
                    tma_load_4d(sa, map, &full_bar[stage], c, tw0 + dw, th0 + dh, tn0);
                    tma_load_3d(sb, &map_w, &full_bar[stage], tap * args.ctot + cc * BK, n0, 0);
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        if (elect_one()) {
            for (int kb = 0; kb < args.num_kb; ++kb) {
                const int stage = kb % STAGES;
                const uint32_t phase = (kb / STAGES) & 1;
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
                const uint64_t da = umma_desc_k_sw128(sa);
                const uint64_t db = umma_desc_k_sw128(sa + A_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    // +32 bytes per K=16 step inside the 128-byte swizzle row (address field is >> 4)
                    umma_bf16_ss(tmem_acc, da + 2 * k, db + 2 * k, IDESC, (kb | k) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
            }
            umma_commit(acc_bar);  // accumulator complete
        }
    } else {
        // =============================== epilogue ===============================
        const int lg = warp & 3;             // TMEM lane group this warp may access
        const int row = lg * 32 + lane;      // row of the 128-row tile
        mbar_wait(acc_bar, 0);
        tc_fence_after();

        bool row_ok;
        int64_t out_row;       // row index into out / residual (pixel index for conv)
        int64_t sample;        // sample index for row_bias
        if (args.mode == 0) {
            int64_t m = (int64_t)m0 + row;
            row_ok = m < args.M;
            out_row = m;
            sample = args.rows_per_sample > 0 ? m / args.rows_per_sample : 0;
        } else {
            int w = row % args.bw; int t = row / args.bw;
            int h = t % args.bh; int n = t / args.bh;
            w += tw0; h += th0; n += tn0;
            row_ok = (w < args.Wg) && (h < args.Hg) && (n < args.Ng);
            int ow = w, oh = h;
            if (args.upsample) { oh = 2 * h + (zb >> 1); ow = 2 * w + (zb & 1); }
            out_row = ((int64_t)n * args.Ho + oh) * args.Wo + ow;
            sample = n;
        }
        const int64_t zoff_o = args.mode == 0 ? (int64_t)zb * args.batch_stride_o : 0;
        const int64_t zoff_r = args.mode == 0 ? (int64_t)zb * args.batch_stride_r : 0;
        const bool geglu = args.flags & GMD_EPI_GEGLU;
        const bool out_f32 = args.flags & GMD_EPI_OUT_F32;
        const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(lg * 32) << 16);

        const int out_cols_tile = geglu ? BN / 2 : BN;
        const int out_col0_tile = geglu ? blockIdx.y * (BN / 2) : n0;
        for (int ch = 0; ch < out_cols_tile / 16; ++ch) {
            uint32_t r0[16], r1[16];
            tmem_ld_32x16(taddr + ch * 16, r0);
            if (geglu) tmem_ld_32x16(taddr + BN / 2 + ch * 16, r1);
            tmem_wait_ld();
            const int col0 = out_col0_tile + ch * 16;  // output column of element 0
            if (!row_ok || col0 >= args.N_out) continue;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r0[j]);
            if (args.flags & GMD_EPI_SCALE) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] *= args.alpha;
            }
            const int ncol = args.N_out - col0 < 16 ? args.N_out - col0 : 16;
            if (geglu) {
                // value half uses bias[col], gate half bias[N_out + col]  (diffusers GEGLU: proj(x).chunk(2))
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float g = __uint_as_float(r1[j]);
                    if (args.bias && j < ncol) { v[j] += __ldg(args.bias + col0 + j); g += __ldg(args.bias + args.N_out + col0 + j); }
                    v[j] = v[j] * gelu_erf(g);
                }
            } else if (args.bias) {
#pragma unroll
                for (int j = 0; j < 16; ++j) if (j < ncol) v[j] += __ldg(args.bias + col0 + j);
            }
            if (args.row_bias) {
                const float* rb = args.row_bias + sample * args.ld_row_bias + col0;
#pragma unroll
                for (int j = 0; j < 16; ++j) if (j < ncol) v[j] += __ldg(rb + j);
            }
            if (args.residual && (args.flags & GMD_EPI_RESIDUAL_F32)) {
                const float* rp = reinterpret_cast<const float*>(args.residual) + zoff_r + out_row * args.ldr + col0;
                if (ncol == 16 && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 a = __ldg(reinterpret_cast<const float4*>(rp) + j);
                        v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w;
                    }
                } else {
                    for (int j = 0; j < ncol; ++j) v[j] += rp[j];
                }
            } else if (args.residual) {
                const __nv_bfloat16* rp = args.residual + zoff_r + out_row * args.ldr + col0;
                if (ncol == 16 && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
                    uint4 a = __ldg(reinterpret_cast<const uint4*>(rp));
                    uint4 b = __ldg(reinterpret_cast<const uint4*>(rp) + 1);
                    uint32_t rr[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) { v[2 * j] += bf16_lo(rr[j]); v[2 * j + 1] += bf16_hi(rr[j]); }
                } else {
                    for (int j = 0; j < ncol; ++j) v[j] += __bfloat162float(rp[j]);
                }
            }
            if (out_f32) {
                float* op = reinterpret_cast<float*>(args.out) + zoff_o + out_row * args.ldo + col0;
                if (ncol == 16 && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        reinterpret_cast<float4*>(op)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                } else {
                    for (int j = 0; j < ncol; ++j) op[j] = v[j];
                }
            } else {
                __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(args.out) + zoff_o + out_row * args.ldo + col0;
                if (ncol == 16 && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
                    uint4 a = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                    uint4 b = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
                    reinterpret_cast<uint4*>(op)[0] = a;
                    reinterpret_cast<uint4*>(op)[1] = b;
                } else {
                    for (int j = 0; j < ncol; ++j) op[j] = __float2bfloat16(v[j]);
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<TMEM_COLS>(tmem_acc);
    }
}

template <int BN, int STAGES>
constexpr size_t smem_bytes() { return (size_t)STAGES * (A_TILE_BYTES + BN * BK * 2) + (2 * STAGES + 1) * 8 + 16 + 1024; }

template <int BN, int STAGES>
int launch(const CUtensorMap* maps_a, const CUtensorMap& map_w, const KernelArgs& args, dim3 grid, cudaStream_t st) {
    static bool configured = false;
    constexpr size_t smem = smem_bytes<BN, STAGES>();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_last_error("gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return kErrCuda; }
        configured = true;
    }
    gemm_kernel<BN, STAGES><<<grid, 192, smem, st>>>(maps_a[0], maps_a[1], maps_a[2], maps_a[3], map_w, args);
    count_launch(1);
    return check_launch("gemm_kernel");
}

int launch_bn(int bn, const CUtensorMap* maps_a, const CUtensorMap& map_w, const KernelArgs& args, dim3 grid, cudaStream_t st) {
    switch (bn) {
        case 160: return launch<160, 5>(maps_a, map_w, args, grid, st);
        case 128: return launch<128, 6>(maps_a, map_w, args, grid, st);
        case 64: return launch<64, 8>(maps_a, map_w, args, grid, st);
        case 32: return launch<32, 8>(maps_a, map_w, args, grid, st);
        default: set_last_error("gemm: unsupported N tile %d", bn); return kErrUnsupported;
    }
}

// N tile: 160 divides every UNet width (320/640/1280 and their multiples); 128 for the VAE widths.
int pick_bn(int64_t n_rows, bool geglu) {
    if (n_rows % 160 == 0) return 160;
    if (n_rows % 128 == 0 || n_rows > 128) return 128;
    if (n_rows > 32 || geglu) return 64;
    return 32;
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace gmd

extern "C" int gmd_gemm_fwd(const gmd_gemm_params* p, void* stream) {
    using namespace gmd;
    if (!p || !p->a || !p->w || !p->out) { set_last_error("gmd_gemm_fwd: null pointer"); return kErrInvalid; }
    if (p->M <= 0 || p->N <= 0 || p->K <= 0) { set_last_error("gmd_gemm_fwd: empty problem M=%lld N=%lld K=%lld", (long long)p->M, (long long)p->N, (long long)p->K); return kErrInvalid; }
    if (p->K % 8 || p->lda % 8 || p->ldw % 8 || !al16(p->a) || !al16(p->w)) {
        set_last_error("gmd_gemm_fwd: K, lda, ldw must be multiples of 8 and A, W 16-byte aligned (TMA)"); return kErrInvalid;
    }
    const bool geglu = p->flags & GMD_EPI_GEGLU;
    if (geglu && (p->N % 2)) { set_last_error("gmd_gemm_fwd: GEGLU needs even N"); return kErrInvalid; }
    int64_t batch = p->batch > 0 ? p->batch : 1;
    int bn = pick_bn(p->N, geglu);
    if (geglu && (p->N % bn)) { set_last_error("gmd_gemm_fwd: GEGLU needs N %% tile == 0 (N=%lld tile=%d)", (long long)p->N, bn); return kErrInvalid; }

    CUtensorMap maps_a[4], map_w;
    {
        uint64_t dims[3] = {(uint64_t)p->K, (uint64_t)p->M, (uint64_t)batch};
        uint64_t strides[3] = {2, (uint64_t)p->lda * 2, (uint64_t)(batch > 1 ? p->stride_a : p->lda * p->M) * 2};
        uint32_t box[3] = {BK, BM, 1};
        int rc = encode_tensor_map_bf16(&maps_a[0], p->a, 3, dims, strides, box, true);
        if (rc) return rc;
        maps_a[1] = maps_a[2] = maps_a[3] = maps_a[0];
    }
    {
        uint64_t dims[3] = {(uint64_t)p->K, (uint64_t)p->N, (uint64_t)batch};
        uint64_t strides[3] = {2, (uint64_t)p->ldw * 2, (uint64_t)(batch > 1 ? p->stride_w : p->ldw * p->N) * 2};
        uint32_t box[3] = {BK, (uint32_t)bn, 1};
        int rc = encode_tensor_map_bf16(&map_w, p->w, 3, dims, strides, box, true);
        if (rc) return rc;
    }
    KernelArgs a{};
    a.mode = 0;
    a.num_kb = (int)((p->K + BK - 1) / BK);
    a.kb_src0 = a.num_kb;
    a.M = p->M;
    a.N_out = (int)(geglu ? p->N / 2 : p->N);
    a.out = p->out; a.ldo = p->ldo;
    a.bias = (p->flags & GMD_EPI_BIAS) ? p->bias : nullptr;
    a.row_bias = (p->flags & GMD_EPI_ROW_BIAS) ? p->row_bias : nullptr;
    a.ld_row_bias = p->ld_row_bias; a.rows_per_sample = p->rows_per_sample;
    a.residual = (p->flags & GMD_EPI_RESIDUAL) ? static_cast<const __nv_bfloat16*>(p->residual) : nullptr;
    a.ldr = p->ldr;
    a.batch_stride_o = p->stride_o; a.batch_stride_r = p->stride_o;
    a.flags = p->flags; a.alpha = p->alpha;
    if ((p->flags & GMD_EPI_BIAS) && !p->bias) { set_last_error("gmd_gemm_fwd: BIAS flag without bias"); return kErrInvalid; }
    if ((p->flags & GMD_EPI_RESIDUAL) && !p->residual) { set_last_error("gmd_gemm_fwd: RESIDUAL flag without residual"); return kErrInvalid; }
    if ((p->flags & GMD_EPI_ROW_BIAS) && (!p->row_bias || p->rows_per_sample <= 0)) { set_last_error("gmd_gemm_fwd: ROW_BIAS needs row_bias and rows_per_sample"); return kErrInvalid; }
    dim3 grid((unsigned)((p->M + BM - 1) / BM), (unsigned)((p->N + bn - 1) / bn), (unsigned)batch);
    return launch_bn(bn, maps_a, map_w, a, grid, static_cast<cudaStream_t>(stream));
}

extern "C" int gmd_conv_fwd(const gmd_conv_params* p, void* stream) {
    using namespace gmd;
    if (!p || !p->x0 || !p->w || !p->out) { set_last_error("gmd_conv_fwd: null pointer"); return kErrInvalid; }
    if (p->ksize != 3 && p->ksize != 1) { set_last_error("gmd_conv_fwd: ksize must be 1 or 3"); return kErrInvalid; }
    if (p->stride != 1 && p->stride != 2) { set_last_error("gmd_conv_fwd: stride must be 1 or 2"); return kErrInvalid; }
    if (p->stride == 2 && (p->upsample || p->x1 || (p->H % 2) || (p->W % 2) || p->ksize != 3)) {
        set_last_error("gmd_conv_fwd: stride 2 needs 3x3, even H/W, single source, no upsample"); return kErrInvalid;
    }
    if (p->C0 % 8 || p->C1 % 8 || (p->x1 && (p->C0 % BK))) { set_last_error("gmd_conv_fwd: channel counts must be multiples of 8 (first source multiple of 64 when concatenating)"); return kErrInvalid; }
    if (p->N <= 0 || p->H <= 0 || p->W <= 0 || p->Cout <= 0) { set_last_error("gmd_conv_fwd: empty problem"); return kErrInvalid; }
    const int C1 = p->x1 ? p->C1 : 0;
    const int ctot = p->C0 + C1;
    const int taps = p->ksize * p->ksize;
    // iteration grid
    int Wg = p->W, Hg = p->H, Wo = p->W, Ho = p->H;
    if (p->stride == 2) { Wg = Wo = p->W / 2; Hg = Ho = p->H / 2; }
    if (p->upsample) { Wo = 2 * p->W; Ho = 2 * p->H; }
    auto pick = [](int extent, int cap) {
        int b = 1;
        while (b * 2 <= cap && extent % (b * 2) == 0) b *= 2;     // largest power of two dividing extent
        if (b < 8 && b < cap) { b = 1; while (b < extent && b < cap) b *= 2; }  // ragged: round up, mask in the epilogue
        return b;
    };
    int bw = pick(Wg, BM);
    int bh = pick(Hg, BM / bw);
    int bn = BM / (bw * bh);
    KernelArgs a{};
    a.mode = 1;
    a.ks = p->ksize; a.stride = p->stride; a.upsample = p->upsample;
    a.chunks0 = (p->C0 + BK - 1) / BK; a.chunks1 = (C1 + BK - 1) / BK;
    a.ctot = ctot;
    a.num_kb = taps * (a.chunks0 + a.chunks1);
    a.bw = bw; a.bh = bh; a.bn = bn;
    a.tiles_w = (Wg + bw - 1) / bw; a.tiles_h = (Hg + bh - 1) / bh;
    a.Wg = Wg; a.Hg = Hg; a.Ng = p->N; a.Wo = Wo; a.Ho = Ho;
    a.N_out = p->Cout;
    a.out = p->out; a.ldo = p->Cout;
    a.bias = (p->flags & GMD_EPI_BIAS) ? p->bias : nullptr;
    a.row_bias = (p->flags & GMD_EPI_ROW_BIAS) ? p->row_bias : nullptr;
    a.ld_row_bias = p->ld_row_bias;
    a.residual = (p->flags & GMD_EPI_RESIDUAL) ? static_cast<const __nv_bfloat16*>(p->residual) : nullptr;
    a.ldr = p->Cout;
    a.flags = p->flags & ~GMD_EPI_GEGLU;
    a.alpha = 1.0f;
    if ((p->flags & GMD_EPI_BIAS) && !p->bias) { set_last_error("gmd_conv_fwd: BIAS flag without bias"); return kErrInvalid; }
    if ((p->flags & GMD_EPI_RESIDUAL) && !p->residual) { set_last_error("gmd_conv_fwd: RESIDUAL flag without residual"); return kErrInvalid; }
    if ((p->flags & GMD_EPI_ROW_BIAS) && !p->row_bias) { set_last_error("gmd_conv_fwd: ROW_BIAS flag without row_bias"); return kErrInvalid; }

    int rows_w = p->Cout_pad > 0 ? p->Cout_pad : p->Cout;
    int bnt = pick_bn(rows_w, false);
    CUtensorMap maps_a[4], map_w;
    const uint32_t box[4] = {BK, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
    if (p->stride == 2) {
        // four parity planes of the input: (c, w/2, h/2, n) with doubled strides
        for (int ph = 0; ph < 2; ++ph)
            for (int pw = 0; pw < 2; ++pw) {
                const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(p->x0) + ((int64_t)ph * p->W + pw) * p->C0;
                uint64_t dims[4] = {(uint64_t)p->C0, (uint64_t)p->W / 2, (uint64_t)p->H / 2, (uint64_t)p->N};
                uint64_t strides[4] = {2, (uint64_t)2 * p->C0 * 2, (uint64_t)2 * p->W * p->C0 * 2, (uint64_t)p->H * p->W * p->C0 * 2};
                int rc = encode_tensor_map_bf16(&maps_a[ph * 2 + pw], base, 4, dims, strides, box, true);
                if (rc) return rc;
            }
    } else {
        uint64_t dims[4] = {(uint64_t)p->C0, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->N};
        uint64_t strides[4] = {2, (uint64_t)p->C0 * 2, (uint64_t)p->W * p->C0 * 2, (uint64_t)p->H * p->W * p->C0 * 2};
        int rc = encode_tensor_map_bf16(&maps_a[0], p->x0, 4, dims, strides, box, true);
        if (rc) return rc;
        maps_a[1] = maps_a[0];
        if (p->x1) {
            uint64_t d1[4] = {(uint64_t)C1, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->N};
            uint64_t s1[4] = {2, (uint64_t)C1 * 2, (uint64_t)p->W * C1 * 2, (uint64_t)p->H * p->W * C1 * 2};
            rc = encode_tensor_map_bf16(&maps_a[1], p->x1, 4, d1, s1, box, true);
            if (rc) return rc;
        }
        maps_a[2] = maps_a[3] = maps_a[0];
    }
    {
        uint64_t dims[3] = {(uint64_t)taps * ctot, (uint64_t)rows_w, 1};
        uint64_t strides[3] = {2, (uint64_t)taps * ctot * 2, (uint64_t)taps * ctot * rows_w * 2};
        uint32_t boxw[3] = {BK, (uint32_t)bnt, 1};
        int rc = encode_tensor_map_bf16(&map_w, p->w, 3, dims, strides, boxw, true);
        if (rc) return rc;
    }
    int tiles_n = (p->N + bn - 1) / bn;
    dim3 grid((unsigned)(a.tiles_w * a.tiles_h * tiles_n), (unsigned)((p->Cout + bnt - 1) / bnt), p->upsample ? 4u : 1u);
    return launch_bn(bnt, maps_a, map_w, a, grid, static_cast<cudaStream_t>(stream));
}
