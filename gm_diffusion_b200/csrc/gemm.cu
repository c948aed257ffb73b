// Kernel family (b): bf16 GEMM and implicit-GEMM convolution on the 5th-generation tensor cores.
//
//   out[M, N] = A[M, K] * W[N, K]^T (+ fused epilogue)
//
// One CTA computes a 128 x BN output tile.  Warp 0 is the TMA producer, warp 1 issues tcgen05.mma
// (one elected thread) into a TMEM accumulator, warps 2-9 are the epilogue (TMEM -> registers ->
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

#ifndef GMD_GEMM_SVC_SLEEP
#define GMD_GEMM_SVC_SLEEP 0   // ns between barrier polls of the producer and MMA threads (0: plain try_wait loop); a polling thread takes issue slots from the epilogue warps of its sub-partition
#endif
#if GMD_GEMM_SVC_SLEEP > 0
#define SVC_WAIT(bar, parity) mbar_wait_sleep(bar, parity, GMD_GEMM_SVC_SLEEP)
#else
#define SVC_WAIT(bar, parity) mbar_wait(bar, parity)
#endif
namespace gmd {
void count_launch(int n);
namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row

struct KernelArgs {
    int mode;       // 0 = GEMM (3-D maps: k, row, batch), 1 = conv (4-D A maps: c, w, h, n)
    int num_kb;     // number of 64-wide K blocks
    // GEMM
    int kb_src0;    // K blocks taken from A source 0 (the rest from source 1)
    // conv
    int ks, stride, upsample;
    int bw_log2, bh_log2;     // the pixel box of a tile has power-of-two extents: row -> (w, h, n) by shifts
    int pad_end;              // stride 2: taps start at input row/col 2*o (zero row/col appended at the end) instead of 2*o - 1
    int chunks0, chunks1;  // 64-channel blocks per tap from source 0 / 1
    int ctot;              // C0 + C1 (weight K pitch per tap)
    int bw, bh, bn;        // pixel box of one 128-row tile (bw*bh*bn == 128)
    int tiles_w, tiles_h;  // tile grid over the (Wg x Hg) iteration grid
    int Wg, Hg, Ng;        // iteration grid (output grid; low-res grid when upsample)
    int Wo, Ho;            // output spatial dims
    // epilogue
    int64_t M;             // GEMM rows
    int N_out;             // valid output columns (after GEGLU halving)
    void* out; int64_t ldo;
    const float* bias;
    const float* row_bias; int64_t ld_row_bias; int64_t rows_per_sample;
    const void* residual; int64_t ldr;
    int64_t batch_stride_o, batch_stride_r;
    int flags;
    float alpha;
    // persistent tile loop: tile -> (super-tile along M, N tile, z)
    int tiles_mt, tiles_n, gz;
    // K = STAGES * 64 (the K = 320 token GEMMs): every k-block of a tile always lands in the same ring slot, so the weight half of the
    // slots is left in place while a CTA stays on one N tile — the CTA then owns a CONTIGUOUS range of tiles (M fastest) instead of a
    // strided one, and a tile costs 80 KB of L2 -> shared-memory traffic instead of 180 KB (these GEMMs ran at the L2's ~9 TB/s)
    int b_resident;
    // split-K: tile also carries a K slice; partial sums go to an fp32 workspace and are reduced by splitk_finalize_kernel
    int ksplit, kb_per_split;
    int64_t split_stride_o;   // elements between the partial-sum planes
    int w_tiled;              // weights pre-tiled as [N tile][k block][BN][64]: every B tile is one contiguous 128*BN-byte read
    const uint8_t* w_bulk;    // non-null: tiles are also PRE-SWIZZLED (smem image) -> one 1-D bulk copy per B tile instead of BN tensor rows
    // GroupNorm statistics of the output (north_star (b)): per (sample, channel pair) sum and sum of squares, accumulated by the
    // epilogue as 64-bit FIXED-POINT integers (2^-24 units): integer addition is associative, so the atomics leave the result
    // bit-reproducible and independent of tile order and batch size; no partial buffer, no fold pass
    long long* gn_sums;       // [sample][N_out / 2][2]; null = none.  Zeroed by the caller
    int gn_ncb;               // N_out / 32
    int64_t gn_rows;          // GEMM mode: output rows per sample
    // LayerNorm folding (gmd_b200.h): producer side — row statistics and a bf16 copy of the output; consumer side — normalise in the epilogue
    long long* ln_out_sums;   // [M][2] fixed point 2^-24, added to
    __nv_bfloat16* ln_out_copy;   // [M][N_out]
    const long long* ln_in_sums;  // [M][2]
    const float* ln_in_c;     // [N]
    float ln_eps, ln_inv_k;
};

// exact-erf GELU with erf from Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, two MUFU ops): erff() costs ~60
// instructions and made the GEGLU epilogue compute-bound (ncu: 112 M warp instructions for one 65536 x 2560 GEMM)
__device__ __forceinline__ float gelu_fast(float x) {
    // Phi(x) = 1 - w for x >= 0, w for x < 0, with w = 0.5 * (a1 t + ... + a5 t^5) * exp(-x^2 / 2), t = 1 / (1 + p |x| / sqrt(2)):
    // 14 instructions, two of them MUFU (the 0.5 is folded into the coefficients, exp(-x^2/2) = 2^(-0.7213475 x^2))
    const float t = rcp_ftz(fmaf(fabsf(x), 0.23164189f, 1.0f));
    const float e = ex2_ftz(x * x * -0.72134752f);
    float h = fmaf(t, 0.5307027145f, -0.7265760135f);
    h = fmaf(t, h, 0.7107068705f);
    h = fmaf(t, h, -0.142248368f);
    h = fmaf(t, h, 0.127414796f);
    const float w = h * t * e;
    return x * (x >= 0.0f ? 1.0f - w : w);
}

#ifndef GMD_GEMM_F32X2
#define GMD_GEMM_F32X2 1   // epilogue arithmetic on packed fp32 pairs (FFMA2 / FADD2 / FMUL2: half the issue slots; the K = 320 GEMMs are epilogue-bound); 0: scalar (A/B)
#endif
// (a + ba) * gelu(g + bg) for a register pair: the same A&S erf as gelu_fast with the polynomial, the products and the final blend on
// packed pairs, and x Phi(x) written without a select: 0.5 x + |x| (0.5 - w)  (19 instructions per pair instead of 2 x 17)
__device__ __forceinline__ void geglu_pair(float a0, float a1, float g0, float g1, float4 ba, float4 bg, int hi, float& o0, float& o1) {
    const uint64_t av = f32x2_add(f32x2_pack(a0, a1), hi ? f32x2_pack(ba.z, ba.w) : f32x2_pack(ba.x, ba.y));
    const uint64_t x = f32x2_add(f32x2_pack(g0, g1), hi ? f32x2_pack(bg.z, bg.w) : f32x2_pack(bg.x, bg.y));
    float x0, x1;
    f32x2_unpack(x, x0, x1);
    const float t0 = rcp_ftz(fmaf(fabsf(x0), 0.23164189f, 1.0f)), t1 = rcp_ftz(fmaf(fabsf(x1), 0.23164189f, 1.0f));
    float q0, q1;
    f32x2_unpack(f32x2_mul(f32x2_mul(x, x), f32x2_bcast(-0.72134752f)), q0, q1);
    const uint64_t e = f32x2_pack(ex2_ftz(q0), ex2_ftz(q1)), t = f32x2_pack(t0, t1);
    uint64_t h = f32x2_fma(t, f32x2_bcast(0.5307027145f), f32x2_bcast(-0.7265760135f));
    h = f32x2_fma(t, h, f32x2_bcast(0.7107068705f));
    h = f32x2_fma(t, h, f32x2_bcast(-0.142248368f));
    h = f32x2_fma(t, h, f32x2_bcast(0.127414796f));
    const uint64_t w = f32x2_mul(f32x2_mul(h, t), e);
    const uint64_t hm = f32x2_fma(w, f32x2_bcast(-1.0f), f32x2_bcast(0.5f));                 // 0.5 - w
    const uint64_t ge = f32x2_fma(f32x2_pack(fabsf(x0), fabsf(x1)), hm, f32x2_mul(x, f32x2_bcast(0.5f)));   // x Phi(x)
    f32x2_unpack(f32x2_mul(av, ge), o0, o1);
}

// v[0..3] += a (two FADD2 when the epilogue runs on packed pairs)
__device__ __forceinline__ void add4(float* v, float4 a) {
#if GMD_GEMM_F32X2
    f32x2_unpack(f32x2_add(f32x2_pack(v[0], v[1]), f32x2_pack(a.x, a.y)), v[0], v[1]);
    f32x2_unpack(f32x2_add(f32x2_pack(v[2], v[3]), f32x2_pack(a.z, a.w)), v[2], v[3]);
#else
    v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
#endif
}

// Shared-memory plan of one CTA.  RB = bytes per element of the residual prefetch buffer (4 holds fp32 or bf16 rows).
template <int MT, int BN, int STAGES, int RB, int HALO = 0, int PAIR = 0>
struct Plan {
    static constexpr int A_SUB = BM * BK * 2;               // one 128-row A sub-tile
    // HALO: one A slot holds the pixel rows of all MT sub-tiles PLUS one image row above and below (up to 64 pixels wide), so
    // the three vertical filter taps read the same copy at row offsets 0 / bw / 2*bw (see the HALO main loop)
    static constexpr int A_BYTES = HALO ? (MT * BM + 2 * 64) * BK * 2 : MT * A_SUB;
    // PAIR (cta_group::2): each CTA of the pair stages HALF of the weight tile; the MMA reads both halves
    static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;    // non-HALO ring slot
    static constexpr int A_SLOTS = 2;                         // HALO: A ring depth (B ring depth is STAGES)
    static constexpr int OFF_B = A_SLOTS * A_BYTES;           // HALO: start of the B ring
    static constexpr int RING_BYTES = HALO ? A_SLOTS * A_BYTES + STAGES * B_BYTES : STAGES * STAGE_BYTES;
    static constexpr int RES_ROW = BN * RB + 16;             // bytes of one private residual row (+16: rows 8 apart share banks, not all)
    static constexpr int OFF_RES = RING_BYTES;
    static constexpr int OFF_BIAS = OFF_RES + BM * RES_ROW;  // BN floats
    static constexpr int OFF_BARS = OFF_BIAS + BN * 4;
    static constexpr int ACC = 2 * MT * BN <= 512 ? 2 : 1;   // TMEM accumulator sets (2 = epilogue overlaps the next tile's main loop)
    static constexpr int ACC_COLS = MT * BN;
    // ROT: 256-row tiles whose two sets do not fit (2 x 2 x 160 > 512) rotate their two accumulators through THREE BN-column
    // slots: tile i uses slots (2i) % 3 and (2i+1) % 3, so the next tile's main loop starts as soon as the epilogue has drained
    // the FIRST sub-tile of this one (the other slot it needs was free all along) — half of the epilogue leaves the critical path
    static constexpr bool ROT = ACC == 1 && MT == 2 && 3 * BN <= 512;
    static constexpr int TOTAL = OFF_BARS + (2 * STAGES + 7 + 2 * A_SLOTS) * 8 + 16;   // the dynamic smem base is declared 1024-byte aligned
    static constexpr int TMEM_USED = ROT ? 3 * BN : ACC * ACC_COLS;
    static constexpr uint32_t TMEM_COLS = TMEM_USED <= 32 ? 32 : TMEM_USED <= 64 ? 64 : TMEM_USED <= 128 ? 128 : TMEM_USED <= 256 ? 256 : 512;
    static_assert(TMEM_USED <= 512, "accumulators exceed TMEM");
    static_assert(TOTAL <= 232448, "shared memory plan exceeds 227 KB");
};

// 1-D bulk copy global -> shared (TMA engine, UBLKCP), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// TMA loads multicast to every CTA of the cluster named in `mask` (same CTA-relative smem offset and mbarrier in each)
__device__ __forceinline__ void tma_load_3d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(mask)
        : "memory");
}
// tcgen05.commit arriving on the same mbarrier offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
// ---- CTA-pair (cta_group::2) forms; instruction strings as in CUTLASS cute/arch/{copy_sm100_tma,mma_sm100_umma,tmem_allocator_sm100}.hpp ----
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the peer bit of a shared-window address: the even (leader) CTA of the pair
// TMA load issued by EITHER CTA of the pair into its own shared memory, completing on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
// D[tmem of both CTAs] (+)= A[256 rows: 128 from each CTA's smem] * B[N rows: N/2 from each CTA's smem]; issued by the leader only
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
// arrive on the LEADER's copy of a barrier (from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_slot) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// 256-bit global store (STG.256, sm_100): one full 32-byte sector per thread per instruction
__device__ __forceinline__ void st_global_256(void* p, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5, uint32_t a6,
                                              uint32_t a7) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5), "r"(a6), "r"(a7)
                 : "memory");
}
// 16-byte load from shared memory by 32-bit shared address (a pointer derived from the dynamic smem base is generic to the compiler:
// it emitted LD.E, which goes through the global/generic path)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void st_f32x8(float* p, const float* v) {
    st_global_256(p, __float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]), __float_as_uint(v[4]),
                  __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7]));
}
__device__ __forceinline__ void st_bf16x16(__nv_bfloat16* p, const float* v) {
    st_global_256(p, pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]), pack_bf16x2(v[8], v[9]),
                  pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
}

// PERSISTENT kernel: one CTA per SM loops over output tiles (tile = blockIdx.x, += gridDim.x).  CTA tile = (MT*128) x BN.
//   * MT = 2 keeps TWO accumulators per tile and issues two MMAs per B tile: the B operand bytes fetched from L2 are
//     amortised over 256 rows (L2 -> SM bandwidth, ~40 B/clk/SM, is what bounds the main loop).
//   * when two accumulator sets fit in TMEM (ACC = 2) the epilogue of tile i runs while the main loop of tile i+1
//     already fills the other set; the smem ring and its mbarrier phases run continuously across tiles.
//   * CL = 2: a CLUSTER of two CTAs works on two neighbouring N tiles of the same M super-tile; each CTA fetches ONE of the two
//     128-row A sub-tiles and TMA-multicasts it into both CTAs' smem, so the A bytes crossing the L2 -> SM fabric are halved
//     (36 KB instead of 52 KB per k block at 256x160).  A smem slot is recycled only after BOTH CTAs' MMAs have read it
//     (tcgen05.commit multicast onto both empty barriers).
template <int MT, int BN, int STAGES, int RB, int CL, int HALO, int PAIR>
__global__ void __launch_bounds__(320, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
            const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_a3,
            const __grid_constant__ CUtensorMap map_w, const KernelArgs args) {
    using P = Plan<MT, BN, STAGES, RB, HALO, PAIR>;
    constexpr uint32_t IDESC = umma_idesc_bf16(PAIR ? 2 * BM : BM, BN, false, false);
    static_assert(!PAIR || CL == 2, "the CTA-pair MMA runs in 2-CTA clusters");
    constexpr int ACC = P::ACC;
    static_assert(!HALO || CL == 1 || PAIR, "the halo main loop is single-CTA unless it runs as a CTA pair");

    // SWIZZLE_128B tiles need a 1024-byte aligned base; declaring the alignment (instead of rounding the pointer up) gives the
    // 1 KB of slack back to the plan — it is what lets the fp32-residual instantiation hold a 4th operand stage
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) { printf("gemm_kernel: dynamic shared memory base is not 1024-byte aligned\n"); __trap(); }
    uint8_t* res_s = smem + P::OFF_RES;
    float* bias_s = reinterpret_cast<float*>(smem + P::OFF_BIAS);
    const uint32_t bias_sa = smem_u32(bias_s);
    // the c vector of a LayerNorm-folding consumer: such a GEMM has no residual (host check), so the private residual rows (>= 2 KB
    // even with RB = 0) are free to hold its BN floats
    static_assert(BM * P::RES_ROW >= BN * 4, "residual region too small for the LayerNorm c vector");
    float* lnc_s = reinterpret_cast<float*>(smem + P::OFF_RES);
    const uint32_t lnc_sa = smem_u32(lnc_s);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + P::OFF_BARS);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;   // [3] per accumulator set (or per rotating slot)
    uint64_t* acc_empty = acc_full + 3;        // [3]
    uint64_t* res_bar = acc_empty + 3;         // residual prefetch rounds
    uint64_t* a_full = res_bar + 1;            // HALO: A ring (P::A_SLOTS each)
    uint64_t* a_empty = a_full + P::A_SLOTS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + P::A_SLOTS);

    static_assert(CL == 1 || MT == 2 || PAIR, "the A multicast splits the two 128-row sub-tiles between the two CTAs");
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
    const int tn_c = PAIR ? args.tiles_n : args.tiles_n / CL;   // N tile groups (a multicast cluster covers CL neighbouring N tiles; a PAIR shares one)
    const int total_tiles = args.tiles_mt * tn_c * args.gz * args.ksplit;
    // b_resident: blocked assignment (CTA b owns tiles [b * per, (b + 1) * per)), else strided
    const int per_cta = (total_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int tile0 = args.b_resident ? (int)blockIdx.x * per_cta : (int)blockIdx.x / CL;
    const int tstride = args.b_resident ? 1 : (int)gridDim.x / CL;
    const int num_tiles = args.b_resident ? min(total_tiles, tile0 + per_cta) : total_tiles;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_a0);
        tma_prefetch_desc(&map_w);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], PAIR ? 1 : CL); }
        for (int s = 0; s < 3; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], PAIR ? 512 : 256); }   // PAIR: both CTAs' epilogues release the leader's
        mbar_init(res_bar, 256);
        for (int s = 0; s < P::A_SLOTS; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        fence_barrier_init();
    }
    if (PAIR) cluster_sync_all();     // both CTAs of the pair are running before the paired TMEM allocation
    if (warp == 1) { if (PAIR) tmem_alloc_2sm<P::TMEM_COLS>(tmem_slot); else tmem_alloc<P::TMEM_COLS>(tmem_slot); }
    tc_fence_before();
    if (CL > 1) cluster_sync_all();   // barrier inits visible cluster-wide before any multicast / remote arrive
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (HALO && elect_one()) {
            // HALO main loop (3x3, stride 1, tile = MT*bh consecutive image rows of one image, full width): for every
            // (horizontal tap sx, 64-channel chunk) ONE TMA box of MT*bh + 2 image rows is loaded and the three vertical taps
            // read it at row offsets 0, bw, 2*bw — the A bytes entering the SM drop from 3 x MT x 16 KB to (MT*128 + 2*bw) x 128 B
            // per three k blocks (-50 % at MT = 2, bw = 64); the weights keep their own ring, one tile per k block.
            const int chunks = args.chunks0 + args.chunks1;
            const uint32_t a_bytes = (uint32_t)(MT * BM + 2 * args.bw) * (BK * 2);
            uint32_t ag = 0, kbg = 0;
            for (int tile = tile0; tile < num_tiles; tile += tstride) {
                const int mt = tile % args.tiles_mt;
                const int n0 = ((tile / args.tiles_mt) % tn_c) * BN;
                int t = PAIR ? mt * (2 * MT) + crank * MT : mt * MT;   // PAIR: each CTA of the pair owns MT consecutive row blocks
                t /= args.tiles_w;                                  // tiles_w == 1
                const int th0 = (t % args.tiles_h) * args.bh, tn0 = t / args.tiles_h;
                for (int sx = 0; sx < 3; ++sx) {
                    for (int cc = 0; cc < chunks; ++cc, ++ag) {
                        const int slot = ag % P::A_SLOTS;
                        SVC_WAIT(&a_empty[slot], ((ag / P::A_SLOTS) & 1) ^ 1);
                        const CUtensorMap* am = cc < args.chunks0 ? &map_a2 : &map_a3;
                        const int ac = cc < args.chunks0 ? cc * BK : (cc - args.chunks0) * BK;
                        if (PAIR) {
                            // each CTA loads ITS tall box and ITS half of the weight tile; all bytes complete on the leader's barriers
                            if (crank == 0) mbar_expect_tx(&a_full[slot], 2 * a_bytes);
                            tma_load_4d_2sm(smem + slot * P::A_BYTES, am, &a_full[slot], ac, sx - 1, th0 - 1, tn0);
                        } else {
                            mbar_expect_tx(&a_full[slot], a_bytes);
                            tma_load_4d(smem + slot * P::A_BYTES, am, &a_full[slot], ac, sx - 1, th0 - 1, tn0);
                        }
                        for (int r = 0; r < 3; ++r, ++kbg) {
                            const int kb = (r * 3 + sx) * chunks + cc;
                            const int stage = kbg % STAGES;
                            SVC_WAIT(&empty_bar[stage], ((kbg / STAGES) & 1) ^ 1);
                            if (PAIR) {
                                if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * P::B_BYTES);
                                tma_load_3d_2sm(smem + P::OFF_B + stage * P::B_BYTES, &map_w, &full_bar[stage], 0,
                                                ((n0 / BN) * args.num_kb + kb) * BN + crank * (BN / 2), 0);
                            } else {
                                mbar_expect_tx(&full_bar[stage], P::B_BYTES);
                                bulk_copy_g2s(smem + P::OFF_B + stage * P::B_BYTES, args.w_bulk + (size_t)((n0 / BN) * args.num_kb + kb) * P::B_BYTES,
                                              P::B_BYTES, &full_bar[stage]);
                            }
                        }
                    }
                }
            }
        } else if (!HALO && elect_one()) {
            const int chunks_per_tap = args.chunks0 + args.chunks1;
            uint32_t kbg = 0;  // k-blocks issued so far (ring position carries across tiles)
            int resident_n0 = -1;   // b_resident: the N tile whose weights the ring slots hold
            for (int tile = tile0; tile < num_tiles; tile += tstride) {
                const int mt = tile % args.tiles_mt;
                const int rest = tile / args.tiles_mt;
                const int n0 = PAIR ? (rest % tn_c) * BN : ((rest % tn_c) * CL + crank) * BN;
                const int zk = rest / tn_c;
                const int zb = zk % args.gz, ksl = zk / args.gz;
                const int kb_begin = ksl * args.kb_per_split;
                const int kb_end = min(args.num_kb, kb_begin + args.kb_per_split);
                const bool tile_keep_b = !PAIR && CL == 1 && args.b_resident && n0 == resident_n0;
                resident_n0 = n0;
                int m0[MT], tw0[MT], th0[MT], tn0[MT];
#pragma unroll
                for (int s = 0; s < MT; ++s) {
                    int t = PAIR ? mt * (2 * MT) + crank * MT + s : mt * MT + s;   // PAIR: this CTA owns MT consecutive 128-row blocks of the pair tile
                    m0[s] = t * BM;
                    int tw = t % args.tiles_w; t /= args.tiles_w;
                    int th = t % args.tiles_h; t /= args.tiles_h;
                    tw0[s] = tw * args.bw; th0[s] = th * args.bh; tn0[s] = t * args.bn;
                }
                for (int kb = kb_begin; kb < kb_end; ++kb, ++kbg) {
                    const int stage = kbg % STAGES;
                    const uint32_t phase = (kbg / STAGES) & 1;
                    SVC_WAIT(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * P::STAGE_BYTES;
                    uint8_t* sb = sa + P::A_BYTES;
                    if (PAIR) {
                        // both CTAs load their own rows of A and their half of the weight tile; every byte completes on the leader's barrier
                        if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * P::STAGE_BYTES);
#pragma unroll
                        for (int s = 0; s < MT; ++s) tma_load_3d_2sm(sa + s * P::A_SUB, &map_a0, &full_bar[stage], kb * BK, m0[s], zb);
                        tma_load_3d_2sm(sb, &map_w, &full_bar[stage], 0, ((n0 / BN) * args.num_kb + kb) * BN + crank * (BN / 2), 0);
                        continue;
                    }
                    const bool keep_b = tile_keep_b;   // slot `stage` already holds W[n0 tile][kb]
                    mbar_expect_tx(&full_bar[stage], keep_b ? P::A_BYTES : P::STAGE_BYTES);
                    if (args.mode == 0) {
                        if (CL > 1) {
                            // this CTA's sub-tile, multicast into both CTAs of the pair
                            tma_load_3d_mc(sa + crank * P::A_SUB, &map_a0, &full_bar[stage], kb * BK, m0[crank], zb, (uint16_t)0x3);
                        } else {
#pragma unroll
                            for (int s = 0; s < MT; ++s) {
                                if (kb < args.kb_src0) tma_load_3d(sa + s * P::A_SUB, &map_a0, &full_bar[stage], kb * BK, m0[s], zb);
                                else tma_load_3d(sa + s * P::A_SUB, &map_a1, &full_bar[stage], (kb - args.kb_src0) * BK, m0[s], zb);
                            }
                        }
                        if (keep_b) {}
                        else if (args.w_bulk) bulk_copy_g2s(sb, args.w_bulk + (size_t)((n0 / BN) * args.num_kb + kb) * P::B_BYTES, P::B_BYTES, &full_bar[stage]);
                        else if (args.w_tiled) tma_load_3d(sb, &map_w, &full_bar[stage], 0, ((n0 / BN) * args.num_kb + kb) * BN, 0);
                        else tma_load_3d(sb, &map_w, &full_bar[stage], kb * BK, n0, zb);
                    } else {
                        const int tap = kb / chunks_per_tap;
                        const int cc = kb - tap * chunks_per_tap;
                        const int r = tap / args.ks, sx = tap - r * args.ks;
                        const int pad = args.ks >> 1;
                        int dh, dw, c;
                        const CUtensorMap* map;
                        if (args.stride == 2) {
                            // input row 2*oh - 1 + r: r=0 -> odd plane, oh-1; r=1 -> even plane, oh; r=2 -> odd plane, oh
                            int ph = (r == 1) ? 0 : 1, pw = (sx == 1) ? 0 : 1;
                            dh = (r == 0) ? -1 : 0; dw = (sx == 0) ? -1 : 0;
                            if (args.pad_end) {
                                // input row 2*oh + r: r=0 -> even plane, oh; r=1 -> odd plane, oh; r=2 -> even plane, oh+1 (TMA zero-fills past the end)
                                ph = (r == 1) ? 1 : 0; pw = (sx == 1) ? 1 : 0;
                                dh = (r == 2) ? 1 : 0; dw = (sx == 2) ? 1 : 0;
                            }
                            const int sel = ph * 2 + pw;
                            map = sel == 0 ? &map_a0 : sel == 1 ? &map_a1 : sel == 2 ? &map_a2 : &map_a3;
                            c = cc * BK;
                        } else {
                            if (args.upsample) {
                                const int py = zb >> 1, px = zb & 1;
                                dh = py == 0 ? (r == 0 ? -1 : 0) : (r == 2 ? 1 : 0);
                                dw = px == 0 ? (sx == 0 ? -1 : 0) : (sx == 2 ? 1 : 0);
                            } else {
                                dh = r - pad; dw = sx - pad;
                            }
                            if (cc < args.chunks0) { map = &map_a0; c = cc * BK; }
                            else { map = &map_a1; c = (cc - args.chunks0) * BK; }
                        }
                        if (CL > 1) {
                            tma_load_4d_mc(sa + crank * P::A_SUB, map, &full_bar[stage], c, tw0[crank] + dw, th0[crank] + dh, tn0[crank], (uint16_t)0x3);
                        } else {
#pragma unroll
                            for (int s = 0; s < MT; ++s)
                                tma_load_4d(sa + s * P::A_SUB, map, &full_bar[stage], c, tw0[s] + dw, th0[s] + dh, tn0[s]);
                        }
                        if (args.w_bulk) bulk_copy_g2s(sb, args.w_bulk + (size_t)((n0 / BN) * args.num_kb + kb) * P::B_BYTES, P::B_BYTES, &full_bar[stage]);
                        else if (args.w_tiled) tma_load_3d(sb, &map_w, &full_bar[stage], 0, ((n0 / BN) * args.num_kb + kb) * BN, 0);
                        else tma_load_3d(sb, &map_w, &full_bar[stage], tap * args.ctot + cc * BK, n0, 0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        if (HALO && (!PAIR || crank == 0) && elect_one()) {   // PAIR: only the leader issues MMAs
            const int groups = 3 * (args.chunks0 + args.chunks1);
            uint32_t ag = 0, kbg = 0, it = 0;
            for (int tile = tile0; tile < num_tiles; tile += tstride, ++it) {
                const int ab = it % ACC;
                uint32_t tsl[MT];                                     // TMEM column base of each sub-tile's accumulator
                if (P::ROT) {
#pragma unroll
                    for (int s = 0; s < MT; ++s) {
                        const uint32_t u = 2 * it + s;
                        SVC_WAIT(&acc_empty[u % 3], ((u / 3) & 1) ^ 1);
                        tsl[s] = tmem_acc + (u % 3) * BN;
                    }
                } else {
                    SVC_WAIT(&acc_empty[ab], ((it / ACC) & 1) ^ 1);
#pragma unroll
                    for (int s = 0; s < MT; ++s) tsl[s] = tmem_acc + ab * P::ACC_COLS + s * BN;
                }
                tc_fence_after();
                for (int g = 0; g < groups; ++g, ++ag) {
                    const int slot = ag % P::A_SLOTS;
                    SVC_WAIT(&a_full[slot], (ag / P::A_SLOTS) & 1);
                    const uint32_t sa = smem_u32(smem + slot * P::A_BYTES);
                    for (int r = 0; r < 3; ++r, ++kbg) {
                        const int stage = kbg % STAGES;
                        SVC_WAIT(&full_bar[stage], (kbg / STAGES) & 1);
                        tc_fence_after();
                        const uint64_t db = umma_desc_k_sw128(smem_u32(smem + P::OFF_B + stage * P::B_BYTES));
                        const uint32_t sar = sa + (uint32_t)(r * args.bw) * (BK * 2);   // vertical tap r: r image rows further down
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
#pragma unroll
                            for (int s = 0; s < MT; ++s) {
                                const uint64_t da = umma_desc_k_sw128(sar + s * P::A_SUB);
                                if (PAIR) umma_bf16_ss_2sm(tsl[s], da + 2 * k, db + 2 * k, IDESC, (g | r | k) != 0 ? 1u : 0u);
                                else umma_bf16_ss(tsl[s], da + 2 * k, db + 2 * k, IDESC, (g | r | k) != 0 ? 1u : 0u);
                            }
                        }
                        if (PAIR) umma_commit_2sm_mc(&empty_bar[stage], (uint16_t)0x3); else umma_commit(&empty_bar[stage]);
                    }
                    if (PAIR) umma_commit_2sm_mc(&a_empty[slot], (uint16_t)0x3); else umma_commit(&a_empty[slot]);
                }
                if (PAIR) {
                    if (P::ROT) { umma_commit_2sm_mc(&acc_full[(2 * it) % 3], (uint16_t)0x3); umma_commit_2sm_mc(&acc_full[(2 * it + 1) % 3], (uint16_t)0x3); }
                    else umma_commit_2sm_mc(&acc_full[ab], (uint16_t)0x3);
                } else if (P::ROT) { umma_commit(&acc_full[(2 * it) % 3]); umma_commit(&acc_full[(2 * it + 1) % 3]); }
                else umma_commit(&acc_full[ab]);
            }
        } else if (!HALO && (!PAIR || crank == 0) && elect_one()) {   // PAIR: only the leader CTA issues the MMAs
            uint32_t kbg = 0, it = 0;
            for (int tile = tile0; tile < num_tiles; tile += tstride, ++it) {
                const int ab = it % ACC;
                uint32_t tsl[MT];
                if (P::ROT) {
#pragma unroll
                    for (int s = 0; s < MT; ++s) {
                        const uint32_t u = 2 * it + s;
                        SVC_WAIT(&acc_empty[u % 3], ((u / 3) & 1) ^ 1);
                        tsl[s] = tmem_acc + (u % 3) * BN;
                    }
                } else {
                    SVC_WAIT(&acc_empty[ab], ((it / ACC) & 1) ^ 1);   // epilogue has drained this accumulator set
#pragma unroll
                    for (int s = 0; s < MT; ++s) tsl[s] = tmem_acc + ab * P::ACC_COLS + s * BN;
                }
                tc_fence_after();
                const int ksl = tile / (args.tiles_mt * tn_c * args.gz);
                const int kb_begin = ksl * args.kb_per_split;
                const int kb_end = min(args.num_kb, kb_begin + args.kb_per_split);
                for (int kb = kb_begin; kb < kb_end; ++kb, ++kbg) {
                    const int stage = kbg % STAGES;
                    const uint32_t phase = (kbg / STAGES) & 1;
                    SVC_WAIT(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * P::STAGE_BYTES);
                    const uint64_t db = umma_desc_k_sw128(sa + P::A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
#pragma unroll
                        for (int s = 0; s < MT; ++s) {
                            // +32 bytes per K=16 step inside the 128-byte swizzle row (address field is >> 4)
                            const uint64_t da = umma_desc_k_sw128(sa + s * P::A_SUB);
                            if (PAIR) umma_bf16_ss_2sm(tsl[s], da + 2 * k, db + 2 * k, IDESC, (kb > kb_begin || k != 0) ? 1u : 0u);
                            else umma_bf16_ss(tsl[s], da + 2 * k, db + 2 * k, IDESC, (kb > kb_begin || k != 0) ? 1u : 0u);
                        }
                    }
                    if (PAIR) umma_commit_2sm_mc(&empty_bar[stage], (uint16_t)0x3);
                    else if (CL > 1) umma_commit_mc(&empty_bar[stage], (uint16_t)0x3);  // both CTAs' producers write this slot
                    else umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
                }
                if (PAIR) {
                    if (P::ROT) { umma_commit_2sm_mc(&acc_full[(2 * it) % 3], (uint16_t)0x3); umma_commit_2sm_mc(&acc_full[(2 * it + 1) % 3], (uint16_t)0x3); }
                    else umma_commit_2sm_mc(&acc_full[ab], (uint16_t)0x3);
                } else if (P::ROT) { umma_commit(&acc_full[(2 * it) % 3]); umma_commit(&acc_full[(2 * it + 1) % 3]); }
                else umma_commit(&acc_full[ab]);  // accumulators of this tile complete
            }
        }
    } else {
        // =============================== epilogue (8 warps) ===============================
        // Two warps per TMEM lane group (a warp may only read lanes 32*(warp%4)..+31); the pair splits the tile's
        // column chunks, so every SM sub-partition has two epilogue warps to hide TMEM / smem / ALU latency.
        // One output row per thread.  Everything besides the accumulator is fetched ahead of time: the bias slice goes to
        // smem and each thread streams ITS OWN residual chunks into a private smem row with cp.async (fire-and-forget
        // 16-byte copies: full memory-level parallelism, no dependent global loads in the epilogue).
        const int lg = warp & 3;                  // TMEM lane group
        const int half = (warp - 2) >> 2;         // which of the two warps of this lane group
        const int row = lg * 32 + lane;           // row of the 128-row sub-tile
        const int et = threadIdx.x - 64;          // 0..255
        const bool geglu = args.flags & GMD_EPI_GEGLU;
        const bool out_f32 = args.flags & GMD_EPI_OUT_F32;
        const bool res_f32 = args.flags & GMD_EPI_RESIDUAL_F32;
        const int out_cols_tile = geglu ? BN / 2 : BN;
        const int res_esize = res_f32 ? 4 : 2;
        const int out_esize = out_f32 ? 4 : 2;
        const int cw = geglu ? 16 : 32;                              // output columns per chunk
        const int nchunks = (out_cols_tile + cw - 1) / cw;
        // the two warps of a lane group own contiguous halves of the tile's column chunks
        const int c_begin = half ? (nchunks + 1) / 2 : 0;
        const int c_end = half ? nchunks : (nchunks + 1) / 2;
        const uint32_t taddr_lane = tmem_acc + (static_cast<uint32_t>(lg * 32) << 16);
        uint8_t* my_res = res_s + row * P::RES_ROW;
        const bool ptrs_ok = ((reinterpret_cast<uintptr_t>(args.out) & 31) == 0) && ((args.ldo * out_esize) % 32 == 0) &&
                             (!args.row_bias || (((reinterpret_cast<uintptr_t>(args.row_bias) & 15) == 0) && (args.ld_row_bias % 4 == 0)));
        const bool res_ptr_ok = args.residual && (BN * res_esize + 16 <= P::RES_ROW) && ((reinterpret_cast<uintptr_t>(args.residual) & 15) == 0) &&
                                ((args.ldr * res_esize) % 16 == 0);

        struct TileInfo { int mt, n0, zb, out_col_tile; bool tile_full, res_async, fast; int64_t zoff_o, zoff_r; };
        auto tile_info = [&](int tile) {
            TileInfo ti;
            ti.mt = tile % args.tiles_mt;
            const int rest = tile / args.tiles_mt;
            const int nt = PAIR ? rest % tn_c : (rest % tn_c) * CL + crank;
            const int zk = rest / tn_c;
            ti.n0 = nt * BN; ti.zb = zk % args.gz;
            ti.out_col_tile = geglu ? nt * (BN / 2) : ti.n0;
            ti.tile_full = ti.out_col_tile + out_cols_tile <= args.N_out && (out_cols_tile % cw) == 0;
            ti.zoff_o = (args.mode == 0 ? (int64_t)ti.zb * args.batch_stride_o : 0) + (int64_t)(zk / args.gz) * args.split_stride_o;
            ti.zoff_r = args.mode == 0 ? (int64_t)ti.zb * args.batch_stride_r : 0;
            ti.res_async = res_ptr_ok && ti.tile_full && ((ti.zoff_r * res_esize) % 16 == 0);
            ti.fast = ti.tile_full && ptrs_ok && (!args.residual || ti.res_async) && ((ti.zoff_o * out_esize) % 32 == 0);
            return ti;
        };
        struct RowInfo { bool ok; int64_t out_row, sample; int blk128; };
        auto row_info = [&](const TileInfo& ti, int s) {
            RowInfo ri;
            int t = PAIR ? ti.mt * (2 * MT) + crank * MT + s : ti.mt * MT + s;
            ri.blk128 = t;
            if (args.mode == 0) {
                int64_t m = (int64_t)t * BM + row;
                ri.ok = m < args.M; ri.out_row = m;
                ri.sample = (args.row_bias && args.rows_per_sample > 0) ? (int64_t)((uint32_t)m / (uint32_t)args.rows_per_sample) : 0;
            } else {
                int tw = t % args.tiles_w; t /= args.tiles_w;
                int th = t % args.tiles_h; t /= args.tiles_h;
                int w = row & (args.bw - 1); int q = row >> args.bw_log2;
                int h = q & (args.bh - 1); int n = q >> args.bh_log2;
                w += tw * args.bw; h += th * args.bh; n += t * args.bn;
                ri.ok = (w < args.Wg) && (h < args.Hg) && (n < args.Ng);
                int ow = w, oh = h;
                if (args.upsample) { oh = 2 * h + (ti.zb >> 1); ow = 2 * w + (ti.zb & 1); }
                ri.out_row = ((int64_t)n * args.Ho + oh) * args.Wo + ow;
                ri.sample = n;
            }
            return ri;
        };
        // One prefetch ROUND: every epilogue thread arrives on res_bar (256 arrivals), threads with a valid row add the byte
        // count of ONE bulk copy (their contiguous column range of their own row).  Issued by a thread only after it has
        // finished reading its private row for the previous round, so rounds never overlap in a row.
        uint32_t res_round = 0;   // rounds issued by this thread (uniform across the epilogue threads)
        auto prefetch_residual = [&](const TileInfo& ti, const RowInfo& ri) {
            if (!ti.res_async) return;
            const uint32_t bytes = (uint32_t)((c_end - c_begin) * cw * res_esize);
            if (ri.ok && bytes > 0) {
                const uint8_t* src = reinterpret_cast<const uint8_t*>(args.residual) +
                                     (ti.zoff_r + ri.out_row * args.ldr + ti.out_col_tile + c_begin * cw) * res_esize;
                mbar_expect_tx(res_bar, bytes);
                bulk_copy_g2s(my_res + c_begin * cw * res_esize, src, bytes, res_bar);
            } else {
                mbar_arrive(res_bar);
            }
            ++res_round;
        };
        uint32_t res_waited = 0;

        uint32_t it = 0;
        int bias_n0 = -1;
        if (tile0 < num_tiles) {
            TileInfo t0 = tile_info(tile0);
            prefetch_residual(t0, row_info(t0, 0));
        }
        for (int tile = tile0; tile < num_tiles; tile += tstride, ++it) {
            const TileInfo ti = tile_info(tile);
            const int ab = it % ACC;
            // bias slice of this N tile -> smem, only when the N tile changes (a CTA's consecutive tiles mostly share it): the
            // reload is a dependent global load plus two 256-thread barriers on the epilogue's critical path, ~1 us per tile,
            // which is what bounded the short-K GEMMs (55 tiles per CTA).  The first barrier keeps slow warps of the previous
            // tile from losing their copy.
            if ((args.bias || args.ln_in_c) && ti.n0 != bias_n0) {
                bias_n0 = ti.n0;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                for (int i = et; i < BN; i += 256) {
                    int col = geglu ? (i < BN / 2 ? ti.out_col_tile + i : args.N_out + ti.out_col_tile + (i - BN / 2)) : ti.n0 + i;
                    int lim = geglu ? 2 * args.N_out : args.N_out;
                    if (args.bias) bias_s[i] = col < lim ? __ldg(args.bias + col) : 0.0f;
                    if (args.ln_in_c) lnc_s[i] = col < lim ? __ldg(args.ln_in_c + col) : 0.0f;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            if (!P::ROT) {
                mbar_wait(&acc_full[ab], (it / ACC) & 1);
                tc_fence_after();
            }
            const int out_col_tile = ti.out_col_tile;
            const int64_t zoff_o = ti.zoff_o, zoff_r = ti.zoff_r;
            const bool fast = ti.fast, res_async = ti.res_async;
            (void)res_async;
#pragma unroll 1
            for (int s = 0; s < MT; ++s) {
                const RowInfo ri = row_info(ti, s);
                if (ti.res_async) { mbar_wait(res_bar, res_waited & 1); ++res_waited; }
                const bool row_ok = ri.ok;
                const int64_t orow = ri.out_row;
                const float* rb_row = args.row_bias ? args.row_bias + ri.sample * args.ld_row_bias : nullptr;
                const uint32_t rot_u = 2 * it + s;
                if (P::ROT) {
                    mbar_wait(&acc_full[rot_u % 3], (rot_u / 3) & 1);
                    tc_fence_after();
                }
                const uint32_t tsub = P::ROT ? taddr_lane + (rot_u % 3) * BN : taddr_lane + ab * P::ACC_COLS + s * BN;
                // LayerNorm folding, consumer side: this row's mean / rstd from the producer's exact integer sums (double: E[x^2] - mean^2)
                float ln_a = 1.0f, ln_b = 0.0f;        // v = acc * ln_a + ln_b * c[n]
                if (args.ln_in_sums && row_ok) {
                    const longlong2 sm = *reinterpret_cast<const longlong2*>(args.ln_in_sums + 2 * orow);
                    const double mean = static_cast<double>(sm.x) * (1.0 / 16777216.0) * args.ln_inv_k;
                    const double var = static_cast<double>(sm.y) * (1.0 / 16777216.0) * args.ln_inv_k - mean * mean;
                    ln_a = rsqrtf(static_cast<float>(var) + args.ln_eps);
                    ln_b = -static_cast<float>(mean) * ln_a;
                }
                float ln_s1 = 0.0f, ln_s2 = 0.0f;      // producer side: this thread's part of the row's sum / sum of squares
                if (fast) {
                    if (geglu) {
#pragma unroll 1
                        for (int c = c_begin; c < c_end; ++c) {
                            uint32_t r0[16], r1[16];
                            tmem_ld_32x16(tsub + c * 16, r0);
                            tmem_ld_32x16(tsub + BN / 2 + c * 16, r1);
                            tmem_wait_ld();
                            if (!row_ok) continue;
                            float v[16];
                            const uint32_t bv = bias_sa + (c * 16) * 4, bg = bias_sa + (BN / 2 + c * 16) * 4;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float4 a = args.bias ? lds128(bv + 16 * j) : make_float4(0, 0, 0, 0), g = args.bias ? lds128(bg + 16 * j) : make_float4(0, 0, 0, 0);
#if GMD_GEMM_F32X2
                                geglu_pair(__uint_as_float(r0[4 * j + 0]), __uint_as_float(r0[4 * j + 1]), __uint_as_float(r1[4 * j + 0]), __uint_as_float(r1[4 * j + 1]), a, g, 0, v[4 * j + 0], v[4 * j + 1]);
                                geglu_pair(__uint_as_float(r0[4 * j + 2]), __uint_as_float(r0[4 * j + 3]), __uint_as_float(r1[4 * j + 2]), __uint_as_float(r1[4 * j + 3]), a, g, 1, v[4 * j + 2], v[4 * j + 3]);
                                continue;
#endif
                                v[4 * j + 0] = (__uint_as_float(r0[4 * j + 0]) + a.x) * gelu_fast(__uint_as_float(r1[4 * j + 0]) + g.x);
                                v[4 * j + 1] = (__uint_as_float(r0[4 * j + 1]) + a.y) * gelu_fast(__uint_as_float(r1[4 * j + 1]) + g.y);
                                v[4 * j + 2] = (__uint_as_float(r0[4 * j + 2]) + a.z) * gelu_fast(__uint_as_float(r1[4 * j + 2]) + g.z);
                                v[4 * j + 3] = (__uint_as_float(r0[4 * j + 3]) + a.w) * gelu_fast(__uint_as_float(r1[4 * j + 3]) + g.w);
                            }
                            const int col0 = out_col_tile + c * 16;
                            if (out_f32) {
                                float* op = reinterpret_cast<float*>(args.out) + zoff_o + orow * args.ldo + col0;
                                st_f32x8(op, v); st_f32x8(op + 8, v + 8);
                            } else {
                                st_bf16x16(reinterpret_cast<__nv_bfloat16*>(args.out) + zoff_o + orow * args.ldo + col0, v);
                            }
                        }
                    } else {
#pragma unroll 1
                        for (int c = c_begin; c < c_end; ++c) {
                            uint32_t r[32];
                            tmem_ld_32x32(tsub + c * 32, r);
                            tmem_wait_ld();
                            if (!row_ok) continue;
                            float v[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                            if (args.flags & GMD_EPI_SCALE) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] *= args.alpha;
                            }
                            const int col0 = out_col_tile + c * 32;
                            if (args.ln_in_sums) {
                                const uint32_t c4 = lnc_sa + (c * 32) * 4;
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const float4 cv = lds128(c4 + 16 * j);
                                    v[4 * j] = fmaf(v[4 * j], ln_a, ln_b * cv.x); v[4 * j + 1] = fmaf(v[4 * j + 1], ln_a, ln_b * cv.y);
                                    v[4 * j + 2] = fmaf(v[4 * j + 2], ln_a, ln_b * cv.z); v[4 * j + 3] = fmaf(v[4 * j + 3], ln_a, ln_b * cv.w);
                                }
                            }
                            if (args.bias) {
                                const uint32_t b4 = bias_sa + (c * 32) * 4;
#pragma unroll
                                for (int j = 0; j < 8; ++j) { float4 a = lds128(b4 + 16 * j); add4(v + 4 * j, a); }
                            }
                            if (rb_row) {
                                const float4* b4 = reinterpret_cast<const float4*>(rb_row + col0);
#pragma unroll
                                for (int j = 0; j < 8; ++j) { float4 a = __ldg(b4 + j); v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w; }
                            }
                            if (args.residual) {
                                if (res_f32) {
#pragma unroll
                                    for (int j = 0; j < 8; ++j) {
                                        float4 a = *reinterpret_cast<const float4*>(my_res + (c * 8 + j) * 16);
                                        add4(v + 4 * j, a);
                                    }
                                } else {
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        uint4 a = *reinterpret_cast<const uint4*>(my_res + (c * 4 + j) * 16);
                                        v[8 * j + 0] += bf16_lo(a.x); v[8 * j + 1] += bf16_hi(a.x); v[8 * j + 2] += bf16_lo(a.y); v[8 * j + 3] += bf16_hi(a.y);
                                        v[8 * j + 4] += bf16_lo(a.z); v[8 * j + 5] += bf16_hi(a.z); v[8 * j + 6] += bf16_lo(a.w); v[8 * j + 7] += bf16_hi(a.w);
                                    }
                                }
                            }
                            if (args.gn_sums) {
                                // GroupNorm statistics of this 32 x 32 block of the output (host: full tiles only, every row valid):
                                // per thread 16 channel-pair sums + 16 sums of squares, then a warp transpose-reduce (31 shuffles): lane
                                // L ends up with the block total of value L, added to the sample's accumulators as a fixed-point integer
                                float sv[32];
#pragma unroll
                                for (int j = 0; j < 16; ++j) { sv[j] = v[2 * j] + v[2 * j + 1]; sv[16 + j] = fmaf(v[2 * j], v[2 * j], v[2 * j + 1] * v[2 * j + 1]); }
#define GMD_GN_STEP(M_, CNT_)                                                                                   \
                                {                                                                               \
                                    const bool up = (lane & (M_)) != 0;                                         \
                                    _Pragma("unroll") for (int i = 0; i < (CNT_) / 2; ++i) {                    \
                                        const float send = up ? sv[i] : sv[i + (CNT_) / 2];                     \
                                        const float keep = up ? sv[i + (CNT_) / 2] : sv[i];                     \
                                        sv[i] = keep + __shfl_xor_sync(0xffffffffu, send, (M_));                \
                                    }                                                                           \
                                }
                                GMD_GN_STEP(16, 32) GMD_GN_STEP(8, 16) GMD_GN_STEP(4, 8) GMD_GN_STEP(2, 4) GMD_GN_STEP(1, 2)
#undef GMD_GN_STEP
                                // lane L < 16: sum of channel pair (col0/2 + L); L >= 16: sum of squares of pair (col0/2 + L - 16)
                                const int64_t smp = args.mode == 0 ? orow / args.gn_rows : ri.sample;       // (warp-uniform: 32-row blocks never straddle samples)
                                unsigned long long* dst = reinterpret_cast<unsigned long long*>(args.gn_sums) +
                                                          ((smp * args.gn_ncb + (col0 >> 5)) * 16 + (lane & 15)) * 2 + (lane >> 4);
                                atomicAdd(dst, static_cast<unsigned long long>(__float2ll_rn(sv[0] * 16777216.0f)));
                            }
                            if (args.ln_out_sums) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) { ln_s1 += v[j]; ln_s2 = fmaf(v[j], v[j], ln_s2); }
                                __nv_bfloat16* cp = args.ln_out_copy + orow * args.N_out + col0;
                                st_bf16x16(cp, v); st_bf16x16(cp + 16, v + 16);
                            }
                            if (out_f32) {
                                float* op = reinterpret_cast<float*>(args.out) + zoff_o + orow * args.ldo + col0;
#pragma unroll
                                for (int j = 0; j < 4; ++j) st_f32x8(op + 8 * j, v + 8 * j);
                            } else {
                                __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(args.out) + zoff_o + orow * args.ldo + col0;
                                st_bf16x16(op, v); st_bf16x16(op + 16, v + 16);
                            }
                        }
                    }
                } else {
                    // ---- generic path (edge tiles, unaligned or ragged outputs): 16 columns at a time, element-wise guards ----
#pragma unroll 1
                    for (int ch = half; ch < (out_cols_tile + 15) / 16; ch += 2) {
                        uint32_t r0[16], r1[16];
                        tmem_ld_32x16(tsub + ch * 16, r0);
                        if (geglu) tmem_ld_32x16(tsub + BN / 2 + ch * 16, r1);
                        tmem_wait_ld();
                        const int col0 = out_col_tile + ch * 16;  // output column of element 0
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
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                float g = __uint_as_float(r1[j]);
                                if (args.bias) { v[j] += bias_s[ch * 16 + j]; g += bias_s[BN / 2 + ch * 16 + j]; }
                                v[j] *= gelu_fast(g);
                            }
                        } else if (args.bias) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] += bias_s[ch * 16 + j];
                        }
                        if (rb_row) for (int j = 0; j < ncol; ++j) v[j] += __ldg(rb_row + col0 + j);
                        if (args.residual) {
                            if (res_f32) {
                                const float* rp = reinterpret_cast<const float*>(args.residual) + zoff_r + orow * args.ldr + col0;
                                for (int j = 0; j < ncol; ++j) v[j] += __ldg(rp + j);
                            } else {
                                const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(args.residual) + zoff_r + orow * args.ldr + col0;
                                for (int j = 0; j < ncol; ++j) v[j] += __bfloat162float(rp[j]);
                            }
                        }
                        if (out_f32) {
                            float* op = reinterpret_cast<float*>(args.out) + zoff_o + orow * args.ldo + col0;
                            for (int j = 0; j < ncol; ++j) op[j] = v[j];
                        } else {
                            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(args.out) + zoff_o + orow * args.ldo + col0;
                            for (int j = 0; j < ncol; ++j) op[j] = __float2bfloat16(v[j]);
                        }
                    }
                }
                if (args.ln_out_sums && fast && row_ok) {
                    // (the two warps of a lane group own different column chunks of the row; integer adds commute: order-independent)
                    unsigned long long* dst = reinterpret_cast<unsigned long long*>(args.ln_out_sums) + 2 * orow;
                    atomicAdd(dst, static_cast<unsigned long long>(__float2ll_rn(ln_s1 * 16777216.0f)));
                    atomicAdd(dst + 1, static_cast<unsigned long long>(__float2ll_rn(ln_s2 * 16777216.0f)));
                }
                // this thread is done with its private residual row: stream in the next sub-tile's / next tile's chunks
                if (s + 1 < MT) {
                    prefetch_residual(ti, row_info(ti, s + 1));
                } else if (tile + tstride < num_tiles) {
                    const TileInfo tn = tile_info(tile + tstride);
                    prefetch_residual(tn, row_info(tn, 0));
                }
                if (P::ROT) {   // this sub-tile's slot is drained: the next tile's main loop may already need it
                    tc_fence_before();
                    if (PAIR) mbar_arrive_leader(&acc_empty[rot_u % 3]); else mbar_arrive(&acc_empty[rot_u % 3]);
                }
            }
            // accumulator set drained -> the MMA warp may start the tile after next into it
            if (!P::ROT) {
                tc_fence_before();
                if (PAIR) mbar_arrive_leader(&acc_empty[ab]); else mbar_arrive(&acc_empty[ab]);
            }
        }
        tc_fence_before();
    }
    if (CL > 1) cluster_sync_all();   // no CTA may exit while its partner can still multicast into it / arrive on its barriers
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_2sm<P::TMEM_COLS>(tmem_acc); else tmem_dealloc<P::TMEM_COLS>(tmem_acc);
    }
}

// Deterministic split-K reduction + the fused epilogue (bias, time-embedding row bias, residual) in one elementwise pass.
__global__ void __launch_bounds__(256) splitk_finalize_kernel(const float* __restrict__ ws, int ksplit, int64_t split_stride, int64_t rows, int N,
                                                              const float* __restrict__ bias, const float* __restrict__ row_bias, int64_t ld_row_bias,
                                                              int64_t rows_per_sample, const void* __restrict__ residual, int res_f32, void* __restrict__ out,
                                                              int out_f32) {
    const int n4 = N / 4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < rows * n4; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / n4;
        const int c = (int)(i - r * n4) * 4;
        float4 acc = __ldg(reinterpret_cast<const float4*>(ws + r * N + c));
        for (int k = 1; k < ksplit; ++k) {
            float4 t = __ldg(reinterpret_cast<const float4*>(ws + k * split_stride + r * N + c));
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        if (bias) { float4 b = __ldg(reinterpret_cast<const float4*>(bias + c)); acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w; }
        if (row_bias) {
            float4 b = __ldg(reinterpret_cast<const float4*>(row_bias + (r / rows_per_sample) * ld_row_bias + c));
            acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
        }
        if (residual) {
            if (res_f32) {
                float4 b = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(residual) + r * N + c));
                acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
            } else {
                uint2 b = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(residual) + r * N + c));
                acc.x += bf16_lo(b.x); acc.y += bf16_hi(b.x); acc.z += bf16_lo(b.y); acc.w += bf16_hi(b.y);
            }
        }
        if (out_f32) *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + r * N + c) = acc;
        else *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + r * N + c) = make_uint2(pack_bf16x2(acc.x, acc.y), pack_bf16x2(acc.z, acc.w));
    }
}

int g_pair_min_kb = 8;            // GMD_PAIR_MIN_KB: shortest K loop (in 64-wide blocks) that takes the CTA-pair path
bool g_disable_pair = false;      // GMD_NO_PAIR=1: GEMMs stay on single-CTA MMAs (A/B measurements)
bool g_disable_halo = false;      // GMD_NO_HALO=1: 3x3 convolutions take the one-box-per-tap main loop (A/B measurements)
bool g_disable_bres = false;      // GMD_NO_BRES=1: no resident-weight ring for the K = 320 GEMMs (A/B measurements)
bool g_disable_cluster = false;   // GMD_NO_CLUSTER=1 in the environment falls back to single-CTA tiles (A/B measurements)

int sm_count() {
    static int sms_of[kMaxDevices] = {};
    static bool env_read = false;
    const int dev = device_ordinal();
    int& sms = sms_of[dev];
    if (!env_read) {
        env_read = true;
        const char* e = getenv("GMD_NO_CLUSTER");
        g_disable_cluster = e && e[0] == '1';
        e = getenv("GMD_NO_BRES");
        g_disable_bres = e && e[0] == '1';
        e = getenv("GMD_NO_HALO");
        g_disable_halo = e && e[0] == '1';
        e = getenv("GMD_NO_PAIR");
        g_disable_pair = e && e[0] == '1';
        e = getenv("GMD_PAIR_MIN_KB");
        if (e && atoi(e) > 0) g_pair_min_kb = atoi(e);
    }
    if (!sms) {
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

template <int MT, int BN, int STAGES, int RB, int CL = 1, int HALO = 0, int PAIR = 0>
int launch(const CUtensorMap* maps_a, const CUtensorMap& map_w, const KernelArgs& args, cudaStream_t st) {
    static bool configured[kMaxDevices] = {};
    const int dev = device_ordinal();
    constexpr size_t smem = Plan<MT, BN, STAGES, RB, HALO, PAIR>::TOTAL;
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(gemm_kernel<MT, BN, STAGES, RB, CL, HALO, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_last_error("gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return kErrCuda; }
        configured[dev] = true;
    }
    // persistent: one CTA per SM; with CL = 2 one cluster (two SMs) per pair of N tiles
    const int64_t items = (int64_t)args.tiles_mt * (PAIR ? args.tiles_n : args.tiles_n / CL) * args.gz * args.ksplit;
    const int max_items = sm_count() / CL;
    const int grid = (int)(items < max_items ? items : max_items) * CL;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = CL > 1 ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_kernel<MT, BN, STAGES, RB, CL, HALO, PAIR>, maps_a[0], maps_a[1], maps_a[2], maps_a[3], map_w, args);
    if (e != cudaSuccess) { set_last_error("gemm: launch: %s", cudaGetErrorString(e)); return kErrCuda; }
    count_launch(1);
    return check_launch("gemm_kernel");
}

// tiles_m = number of 128-row tiles; picks the CTA tile and launches
// 256-row CTA tiles (two accumulators per tile, B tile reused) for long K loops that still fill the 148 SMs
// (split-K launches count their K slices as tiles: the 8x8-resolution convolutions of a 16-sample batch are 64 output tiles x 4 slices —
// as 256-row tiles 128 items of 52 KB per k-block in one wave instead of 256 items of 36 KB in 1.73)
inline bool want_mt2(int bn, bool res_f32, int64_t tiles_m, int64_t tiles_n, unsigned gz, int num_kb, int ksplit = 1) {
    return (bn == 160 || bn == 128) && !res_f32 && tiles_m >= 2 && num_kb >= 16 && ((tiles_m + 1) / 2) * tiles_n * gz * (ksplit > 1 ? ksplit : 1) >= 120;
}

int launch_cfg(int bn, int64_t tiles_m, int64_t tiles_n, unsigned gz, const CUtensorMap* maps_a, const CUtensorMap& map_w,
               KernelArgs args, cudaStream_t st, bool halo_ok = false, bool pair_ok = false, int halo_mt = 0) {
    // 256-row CTA tiles (two accumulators per tile, B tile reused) for long K loops that still fill the 148 SMs; short K loops
    // keep 128-row tiles with two accumulator SETS so the epilogue overlaps the next tile's main loop.  The fp32 residual
    // stream needs the larger private-row buffer, which only fits next to the 128-row pipeline.
    const bool res_f32 = args.residual && (args.flags & GMD_EPI_RESIDUAL_F32);
    const bool no_res = !args.residual;
    const bool mt2 = want_mt2(bn, res_f32, tiles_m, tiles_n, gz, args.ksplit > 1 ? args.kb_per_split : args.num_kb, args.ksplit);
    args.tiles_mt = (int)(mt2 ? (tiles_m + 1) / 2 : tiles_m);
    args.tiles_n = (int)tiles_n;
    args.b_resident = 0;
    args.gz = (int)gz;
    if (args.ksplit < 1) { args.ksplit = 1; args.kb_per_split = args.num_kb; args.split_stride_o = 0; }
    // pair neighbouring N tiles into 2-CTA clusters that share (multicast) the A operand
    (void)sm_count();  // also reads GMD_NO_CLUSTER once
    const bool cl2 = mt2 && (tiles_n % 2 == 0) && args.ksplit == 1 && (args.mode == 1 || args.kb_src0 == args.num_kb) && !g_disable_cluster;
    if (pair_ok && halo_ok && args.ksplit == 1) {
        // halo convolution loop as CTA pairs: each CTA owns one 128-row block (its own tall box) and half of the weight tile, so the
        // weight ring is 10 KB per slot; maps_a[2..3] and map_w were built for this by the caller
        (void)halo_mt;   // pairs are used for the 128-row-tile cases only (see gmd_conv_fwd)
        args.tiles_mt = (int)((tiles_m + 1) / 2);
        return no_res ? launch<1, 160, 10, 0, 2, 1, 1>(maps_a, map_w, args, st) : launch<1, 160, 8, 2, 2, 1, 1>(maps_a, map_w, args, st);
    }
    if (pair_ok && args.ksplit == 1) {
        // CTA pairs (cta_group::2): map_w was built with the half-tile box by the caller.  512-row pair tiles (two 256-row MMAs)
        // when that still leaves ~one item per pair of SMs and K is long, else 256-row pair tiles with two accumulator sets.
        const bool pmt2 = args.num_kb >= 16 && ((tiles_m + 3) / 4) * tiles_n >= 60 && !res_f32;
        args.tiles_mt = (int)(pmt2 ? (tiles_m + 3) / 4 : (tiles_m + 1) / 2);
        if (no_res) return pmt2 ? launch<2, 160, 5, 0, 2, 0, 1>(maps_a, map_w, args, st) : launch<1, 160, 8, 0, 2, 0, 1>(maps_a, map_w, args, st);
        if (res_f32) return launch<1, 160, 5, 4, 2, 0, 1>(maps_a, map_w, args, st);        // fp32 residual rows (84 KB) + 5 stages of 26 KB
        return pmt2 ? launch<2, 160, 4, 2, 2, 0, 1>(maps_a, map_w, args, st) : launch<1, 160, 7, 2, 2, 0, 1>(maps_a, map_w, args, st);
    }
    if (halo_ok && args.ksplit == 1 && !g_disable_halo) {
        // vertical-tap halo reuse (maps_a[2], maps_a[3] hold the tall boxes, built for this MT by the caller): -31 % operand bytes
        // per k block at 256x160, -26 % at 128x160
        // (no residual: the private residual rows are not needed and their 43 KB go to deeper operand rings, RB = 0)
        if (bn == 160 && no_res) return mt2 ? launch<2, 160, 6, 0, 1, 1>(maps_a, map_w, args, st) : launch<1, 160, 7, 0, 1, 1>(maps_a, map_w, args, st);
        if (bn == 160) return mt2 ? launch<2, 160, 4, 2, 1, 1>(maps_a, map_w, args, st) : launch<1, 160, 4, 2, 1, 1>(maps_a, map_w, args, st);
        if (bn == 128) return mt2 ? launch<2, 128, 4, 2, 1, 1>(maps_a, map_w, args, st) : launch<1, 128, 4, 2, 1, 1>(maps_a, map_w, args, st);
    }
    switch (bn) {
        // 128-row tiles: the main loop is bound by the operand bytes in flight (ring capacity / TMA latency — the "40 B/clk/SM" of the
        // 3-stage ring is 108 KB per ~2700 cycles), so everything the residual row buffer does not need goes to more stages
        case 160:
            // K = 320 (five k-blocks = five ring slots): weights stay resident in the ring while a CTA stays on one N tile
            if (!mt2 && !res_f32 && args.mode == 0 && args.num_kb == 5 && args.ksplit == 1 && gz == 1 && !g_disable_bres) {
                args.b_resident = 1;
                return no_res ? launch<1, 160, 5, 0>(maps_a, map_w, args, st) : launch<1, 160, 5, 2>(maps_a, map_w, args, st);
            }
            if (no_res) return mt2 ? (cl2 ? launch<2, 160, 4, 0, 2>(maps_a, map_w, args, st) : launch<2, 160, 4, 0>(maps_a, map_w, args, st))
                                   : launch<1, 160, 6, 0>(maps_a, map_w, args, st);
            return mt2 ? (cl2 ? launch<2, 160, 3, 2, 2>(maps_a, map_w, args, st) : launch<2, 160, 3, 2>(maps_a, map_w, args, st))
                       : (res_f32 ? launch<1, 160, 4, 4>(maps_a, map_w, args, st) : launch<1, 160, 5, 2>(maps_a, map_w, args, st));
        case 128: return mt2 ? (cl2 ? launch<2, 128, 3, 2, 2>(maps_a, map_w, args, st) : launch<2, 128, 3, 2>(maps_a, map_w, args, st))
                             : (res_f32 ? launch<1, 128, 4, 4>(maps_a, map_w, args, st) : launch<1, 128, 5, 2>(maps_a, map_w, args, st));
        case 64: return launch<1, 64, 6, 4>(maps_a, map_w, args, st);
        case 32: return launch<1, 32, 6, 4>(maps_a, map_w, args, st);
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

// Split the K loop over several CTAs when the output has too few tiles to fill the machine (8x8-resolution layers: 32-64
// tiles, 90-360 k-blocks each).  Returns the number of splits (1 = none) for a problem whose epilogue the finalize kernel can do.
int plan_splitk(int64_t tiles, int num_kb, int64_t rows, int N, const void* workspace, int64_t workspace_bytes, bool plain_layout) {
    if (!workspace || !plain_layout || tiles >= 100 || num_kb < 32 || (N % 4)) return 1;
    int ks = (int)(sm_count() / tiles);
    if (ks > num_kb / 16) ks = num_kb / 16;
    if (ks > 8) ks = 8;
    while (ks > 1 && (int64_t)ks * rows * N * 4 > workspace_bytes) --ks;
    return ks < 2 ? 1 : ks;
}

// Split rule for the low-resolution convolutions (8x8 and below: 32-64 output tiles on 148 SMs, 180-360 k-blocks each, every CTA
// bound by its own operand ingest).  It depends only on the per-image geometry and K — never on the batch — so an image goes
// through the same summation order whatever batch it is part of, and image sharding stays bit-exact.
int conv_splitk_rule(int Ho, int Wo, int num_kb, int Cout, bool plain_layout) {
    if (!plain_layout || (Cout % 4) || Ho * Wo > 64 || num_kb < 64) return 1;
    return 4;
}

int launch_finalize(const KernelArgs& a, const float* ws, int ksplit, int64_t rows, int N, const float* bias, const float* row_bias,
                    int64_t ld_row_bias, int64_t rows_per_sample, const void* residual, bool res_f32, void* out, bool out_f32, cudaStream_t st) {
    int64_t items = rows * (N / 4);
    int grid = (int)((items + 255) / 256);
    if (grid > sm_count() * 8) grid = sm_count() * 8;
    splitk_finalize_kernel<<<grid, 256, 0, st>>>(ws, ksplit, rows * N, rows, N, bias, row_bias, ld_row_bias, rows_per_sample, residual, res_f32 ? 1 : 0,
                                                 out, out_f32 ? 1 : 0);
    count_launch(1);
    return check_launch("splitk_finalize");
}

}  // namespace
}  // namespace gmd

namespace gmd { namespace {
// the GEMM epilogue can emit GroupNorm statistics when every tile is full and takes the plain (non-GEGLU, unbatched) fast path
inline bool gemm_gn_ok(const gmd_gemm_params* p, int bn) {
    const int64_t batch = p->batch > 0 ? p->batch : 1;
    return batch == 1 && !(p->flags & (GMD_EPI_GEGLU | GMD_EPI_SCALE)) && (bn % 32) == 0 && (p->N % bn) == 0 && (p->M % BM) == 0 && p->ldo == p->N;
}
} }

extern "C" int gmd_gemm_fwd(const gmd_gemm_params* p, void* stream) {
    using namespace gmd;
    if (!p || !p->a || !p->w || !p->out) { set_last_error("gmd_gemm_fwd: null pointer"); return kErrInvalid; }
    if (p->M <= 0 || p->N <= 0 || p->K <= 0) { set_last_error("gmd_gemm_fwd: empty problem M=%lld N=%lld K=%lld", (long long)p->M, (long long)p->N, (long long)p->K); return kErrInvalid; }
    if (p->K % 8 || p->lda % 8 || (!p->w_tiled && (p->ldw % 8)) || !al16(p->a) || !al16(p->w)) {
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
    const int w_tile = p->w_tiled >= 1000 ? p->w_tiled - 1000 : p->w_tiled;
    (void)sm_count();   // reads the GMD_NO_* switches once
    const int64_t num_kb0 = (p->K + BK - 1) / BK, tm0 = (p->M + BM - 1) / BM, tn0 = (p->N + bn - 1) / bn;
    // CTA-pair main loop: tiled weights (one half-tile TMA box per CTA), no residual stream, enough rows for every pair of SMs,
    // K long enough for the weight bytes to matter, and not a candidate for the GEMM split-K
    const bool pair_ok = !g_disable_pair && bn == 160 && p->w_tiled && w_tile == bn && batch == 1 && num_kb0 >= g_pair_min_kb &&
                         ((tm0 + 1) / 2) * tn0 >= 60 && !p->workspace;
    if (pair_ok) {
        const uint64_t total_rows = (uint64_t)tn0 * num_kb0 * bn;
        uint64_t dims[3] = {BK, total_rows, 1};
        uint64_t strides[3] = {2, BK * 2, total_rows * BK * 2};
        uint32_t box[3] = {BK, (uint32_t)bn / 2, 1};
        // pre-swizzled tiles are already the shared-memory image: copy them verbatim; plain tiles get swizzled by the TMA
        int rc = encode_tensor_map_bf16(&map_w, p->w, 3, dims, strides, box, p->w_tiled < 1000);
        if (rc) return rc;
    } else if (p->w_tiled) {
        if (w_tile != bn || batch != 1) { set_last_error("gmd_gemm_fwd: tiled weights were packed for N tile %d, kernel picks %d (batch %lld)", p->w_tiled, bn, (long long)batch); return kErrInvalid; }
        const uint64_t total_rows = (uint64_t)((p->N + bn - 1) / bn) * ((p->K + BK - 1) / BK) * bn;
        uint64_t dims[3] = {BK, total_rows, 1};
        uint64_t strides[3] = {2, BK * 2, total_rows * BK * 2};
        uint32_t box[3] = {BK, (uint32_t)bn, 1};
        int rc = encode_tensor_map_bf16(&map_w, p->w, 3, dims, strides, box, true);
        if (rc) return rc;
    } else {
        uint64_t dims[3] = {(uint64_t)p->K, (uint64_t)p->N, (uint64_t)batch};
        uint64_t strides[3] = {2, (uint64_t)p->ldw * 2, (uint64_t)(batch > 1 ? p->stride_w : p->ldw * p->N) * 2};
        uint32_t box[3] = {BK, (uint32_t)bn, 1};
        int rc = encode_tensor_map_bf16(&map_w, p->w, 3, dims, strides, box, true);
        if (rc) return rc;
    }
    KernelArgs a{};
    a.mode = 0;
    a.w_tiled = p->w_tiled ? 1 : 0;
    a.w_bulk = p->w_tiled >= 1000 ? static_cast<const uint8_t*>(p->w) : nullptr;
    a.num_kb = (int)((p->K + BK - 1) / BK);
    a.kb_src0 = a.num_kb;
    a.M = p->M;
    a.N_out = (int)(geglu ? p->N / 2 : p->N);
    a.out = p->out; a.ldo = p->ldo;
    a.bias = (p->flags & GMD_EPI_BIAS) ? p->bias : nullptr;
    a.row_bias = (p->flags & GMD_EPI_ROW_BIAS) ? p->row_bias : nullptr;
    a.ld_row_bias = p->ld_row_bias; a.rows_per_sample = p->rows_per_sample;
    a.residual = (p->flags & GMD_EPI_RESIDUAL) ? p->residual : nullptr;
    a.ldr = p->ldr;
    a.batch_stride_o = p->stride_o; a.batch_stride_r = p->stride_o;
    a.tiles_w = a.tiles_h = 1; a.bw = BM; a.bh = a.bn = 1;  // (unused in GEMM mode; keep the index math well defined)
    a.flags = p->flags; a.alpha = p->alpha;
    if ((p->flags & GMD_EPI_BIAS) && !p->bias) { set_last_error("gmd_gemm_fwd: BIAS flag without bias"); return kErrInvalid; }
    if ((p->flags & GMD_EPI_RESIDUAL) && !p->residual) { set_last_error("gmd_gemm_fwd: RESIDUAL flag without residual"); return kErrInvalid; }
    if ((p->flags & GMD_EPI_ROW_BIAS) && (!p->row_bias || p->rows_per_sample <= 0)) { set_last_error("gmd_gemm_fwd: ROW_BIAS needs row_bias and rows_per_sample"); return kErrInvalid; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t tiles_m = (p->M + BM - 1) / BM, tiles_n = (p->N + bn - 1) / bn;
    const bool plain = batch == 1 && !geglu && !(p->flags & GMD_EPI_SCALE) && p->ldo == p->N && (!a.residual || p->ldr == p->N);
    const int ks = plan_splitk(tiles_m * tiles_n, a.num_kb, p->M, (int)p->N, p->workspace, p->workspace_bytes, plain);
    if (p->gn_sums) {
        if (!gemm_gn_ok(p, bn) || ks > 1 || p->gn_rows_per_sample <= 0 || (p->gn_rows_per_sample % BM) || (p->M % p->gn_rows_per_sample)) {
            set_last_error("gmd_gemm_fwd: gn_sums is not available for this call (see gmd_gemm_gn_sums_ok)"); return kErrUnsupported;
        }
        a.gn_sums = static_cast<long long*>(p->gn_sums); a.gn_ncb = (int)(p->N / 32); a.gn_rows = p->gn_rows_per_sample;
    }
    if (p->ln_out_sums || p->ln_in_sums) {
        const int oes = (p->flags & GMD_EPI_OUT_F32) ? 4 : 2;
        const bool ok = batch == 1 && !geglu && ks == 1 && !(p->ln_in_sums && a.residual) && (p->N % bn) == 0 && (reinterpret_cast<uintptr_t>(p->out) & 31) == 0 && ((p->ldo * oes) % 32) == 0 &&
                        (!a.residual || ((reinterpret_cast<uintptr_t>(p->residual) & 15) == 0 &&
                                         ((p->ldr * ((p->flags & GMD_EPI_RESIDUAL_F32) ? 4 : 2)) % 16) == 0)) &&
                        (!p->ln_out_sums || (p->ln_out_copy && (reinterpret_cast<uintptr_t>(p->ln_out_copy) & 31) == 0 && (reinterpret_cast<uintptr_t>(p->ln_out_sums) & 15) == 0)) &&
                        (!p->ln_in_sums || (p->ln_in_c && (reinterpret_cast<uintptr_t>(p->ln_in_c) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->ln_in_sums) & 15) == 0));
        if (!ok) { set_last_error("gmd_gemm_fwd: LayerNorm folding needs batch == 1, no GEGLU, no split-K, full N tiles and 32-byte aligned rows"); return kErrUnsupported; }
        a.ln_out_sums = static_cast<long long*>(p->ln_out_sums); a.ln_out_copy = static_cast<__nv_bfloat16*>(p->ln_out_copy);
        a.ln_in_sums = static_cast<const long long*>(p->ln_in_sums); a.ln_in_c = p->ln_in_c;
        a.ln_eps = p->ln_eps; a.ln_inv_k = 1.0f / static_cast<float>(p->K);
    }
    if (ks > 1) {
        KernelArgs b = a;
        b.ksplit = ks; b.kb_per_split = (a.num_kb + ks - 1) / ks; b.split_stride_o = p->M * p->N;
        b.out = p->workspace; b.ldo = p->N; b.bias = nullptr; b.row_bias = nullptr; b.residual = nullptr; b.flags = GMD_EPI_OUT_F32;
        int rc = launch_cfg(bn, tiles_m, tiles_n, 1u, maps_a, map_w, b, st);
        if (rc) return rc;
        return launch_finalize(a, static_cast<const float*>(p->workspace), ks, p->M, (int)p->N, a.bias, a.row_bias, a.ld_row_bias,
                               a.rows_per_sample > 0 ? a.rows_per_sample : p->M, a.residual, p->flags & GMD_EPI_RESIDUAL_F32, p->out,
                               p->flags & GMD_EPI_OUT_F32, st);
    }
    return launch_cfg(bn, tiles_m, tiles_n, (unsigned)batch, maps_a, map_w, a, st, false, pair_ok);
}

extern "C" int gmd_gemm_gn_sums_ok(const gmd_gemm_params* p, int64_t rows_per_sample) {
    using namespace gmd;
    // whole 128-row tiles per sample: whether statistics come from the epilogue must not depend on how many samples are in the batch
    if (!p || rows_per_sample <= 0 || (rows_per_sample % BM) || (p->M % rows_per_sample)) return 0;
    const int bn = pick_bn((int)p->N, (p->flags & GMD_EPI_GEGLU) != 0);
    return (gemm_gn_ok(p, bn) && !p->workspace) ? 1 : 0;
}

extern "C" int gmd_conv_gn_sums_ok(const gmd_conv_params* p) {
    using namespace gmd;
    if (!p || (p->ksize != 3 && p->ksize != 1) || p->N <= 0) return 0;
    int Wg = p->W, Hg = p->H, Wo = p->W, Ho = p->H;
    if (p->stride == 2) { Wg = Wo = p->W / 2; Hg = Ho = p->H / 2; }
    if (p->upsample) { Wo = 2 * p->W; Ho = 2 * p->H; }
    auto pick = [](int extent, int cap) {          // (as in gmd_conv_fwd)
        int b = 1;
        while (b * 2 <= cap && extent % (b * 2) == 0) b *= 2;
        if (b < 8 && b < cap) { b = 1; while (b < extent && b < cap) b *= 2; }
        return b;
    };
    const int bw = pick(Wg, BM), bh = pick(Hg, BM / bw), bn = BM / (bw * bh);
    const int C1 = p->x1 ? p->C1 : 0;
    const int num_kb = p->ksize * p->ksize * ((p->C0 + BK - 1) / BK + (C1 + BK - 1) / BK);
    const int rows_w = p->w_tiled ? p->Cout : (p->Cout_pad > 0 ? p->Cout_pad : p->Cout);
    const int bnt = pick_bn(rows_w, false);
    const int ks = p->workspace ? conv_splitk_rule(Ho, Wo, num_kb, p->Cout, !p->upsample && p->stride == 1) : 1;
    return (ks > 1 || bn != 1 || (Wg % bw) || (Hg % bh) || (p->Cout % bnt) || (bnt % 32)) ? 0 : 1;
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
    a.pad_end = (p->flags & GMD_CONV_PAD_END) ? 1 : 0;
    if (a.pad_end && p->stride != 2) { set_last_error("gmd_conv_fwd: GMD_CONV_PAD_END needs stride 2"); return kErrInvalid; }
    a.chunks0 = (p->C0 + BK - 1) / BK; a.chunks1 = (C1 + BK - 1) / BK;
    a.ctot = ctot;
    a.num_kb = taps * (a.chunks0 + a.chunks1);
    a.bw = bw; a.bh = bh; a.bn = bn;
    a.bw_log2 = 0; while ((1 << a.bw_log2) < bw) ++a.bw_log2;
    a.bh_log2 = 0; while ((1 << a.bh_log2) < bh) ++a.bh_log2;
    a.tiles_w = (Wg + bw - 1) / bw; a.tiles_h = (Hg + bh - 1) / bh;
    a.Wg = Wg; a.Hg = Hg; a.Ng = p->N; a.Wo = Wo; a.Ho = Ho;
    a.N_out = p->Cout;
    a.out = p->out; a.ldo = p->Cout;
    a.bias = (p->flags & GMD_EPI_BIAS) ? p->bias : nullptr;
    a.row_bias = (p->flags & GMD_EPI_ROW_BIAS) ? p->row_bias : nullptr;
    a.ld_row_bias = p->ld_row_bias;
    a.residual = (p->flags & GMD_EPI_RESIDUAL) ? p->residual : nullptr;
    a.ldr = p->Cout;
    a.flags = p->flags & ~(GMD_EPI_GEGLU | GMD_CONV_PAD_END);
    a.alpha = 1.0f;
    if ((p->flags & GMD_EPI_BIAS) && !p->bias) { set_last_error("gmd_conv_fwd: BIAS flag without bias"); return kErrInvalid; }
    if ((p->flags & GMD_EPI_RESIDUAL) && !p->residual) { set_last_error("gmd_conv_fwd: RESIDUAL flag without residual"); return kErrInvalid; }
    if ((p->flags & GMD_EPI_ROW_BIAS) && !p->row_bias) { set_last_error("gmd_conv_fwd: ROW_BIAS flag without row_bias"); return kErrInvalid; }

    int rows_w = p->w_tiled ? p->Cout : (p->Cout_pad > 0 ? p->Cout_pad : p->Cout);
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
    // HALO main loop eligibility: 3x3 stride 1 without upsample, a tile = whole image rows of one image (tiles_w == 1, bn == 1),
    // pairs of 128-row sub-tiles vertically adjacent inside an image, pre-swizzled weight tiles (bulk-copied)
    bool halo_ok = p->ksize == 3 && p->stride == 1 && !p->upsample && bw <= 64 && bw >= 8 && bw == p->W && bn == 1 && bw * bh == BM &&
                   p->w_tiled >= 1000 && !(p->flags & GMD_EPI_RESIDUAL_F32) && (bnt == 160 || bnt == 128) && a.num_kb >= 16;
    bool conv_pair = false;   // run the halo loop as CTA pairs (cta_group::2)
    int mtv = 1;              // 128-row blocks per CTA tile
    if (halo_ok && conv_splitk_rule(Ho, Wo, a.num_kb, p->Cout, true) > 1 && p->workspace) halo_ok = false;
    if (halo_ok) {
        (void)sm_count();
        const int64_t tm = (int64_t)a.tiles_w * a.tiles_h * ((p->N + bn - 1) / bn), tnn = (p->Cout + bnt - 1) / bnt;
        // CTA pairs only where the single-CTA choice would be 128-row tiles (batch-8 layers at 16x16: 128 CTAs on 148 SMs, +6 %);
        // the 256-row halo tiles are MMA-bound already and measured neutral as pairs (9.26 vs 9.22-9.28 ms of convolutions per step)
        const bool single_mt2 = want_mt2(bnt, false, tm, tnn, 1u, a.num_kb);
        conv_pair = !g_disable_pair && bnt == 160 && !single_mt2 && ((tm + 1) / 2) * tnn >= 60;
        if (conv_pair) {
            mtv = 1;
        } else {
            mtv = single_mt2 ? 2 : 1;
            if (mtv == 2 && (a.tiles_h % 2)) halo_ok = false;     // 256-row tiles must pair two row blocks of the SAME image
        }
    }
    if (halo_ok) {
        const uint32_t hbox[4] = {BK, (uint32_t)bw, (uint32_t)(mtv * bh + 2), 1};
        uint64_t dims[4] = {(uint64_t)p->C0, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->N};
        uint64_t strides[4] = {2, (uint64_t)p->C0 * 2, (uint64_t)p->W * p->C0 * 2, (uint64_t)p->H * p->W * p->C0 * 2};
        int rc = encode_tensor_map_bf16(&maps_a[2], p->x0, 4, dims, strides, hbox, true);
        if (rc) return rc;
        maps_a[3] = maps_a[2];
        if (p->x1) {
            uint64_t d1[4] = {(uint64_t)C1, (uint64_t)p->W, (uint64_t)p->H, (uint64_t)p->N};
            uint64_t s1[4] = {2, (uint64_t)C1 * 2, (uint64_t)p->W * C1 * 2, (uint64_t)p->H * p->W * C1 * 2};
            rc = encode_tensor_map_bf16(&maps_a[3], p->x1, 4, d1, s1, hbox, true);
            if (rc) return rc;
        }
    }
    if (halo_ok && conv_pair) {
        // half-tile boxes of the pre-swizzled weight image, copied verbatim (SWIZZLE_NONE): one per CTA of the pair
        if (p->w_tiled - 1000 != bnt) { set_last_error("gmd_conv_fwd: tiled weights were packed for N tile %d, kernel picks %d", p->w_tiled, bnt); return kErrInvalid; }
        const uint64_t total_rows = (uint64_t)((p->Cout + bnt - 1) / bnt) * a.num_kb * bnt;
        uint64_t dims[3] = {BK, total_rows, 1};
        uint64_t strides[3] = {2, BK * 2, total_rows * BK * 2};
        uint32_t boxw[3] = {BK, (uint32_t)bnt / 2, 1};
        int rc = encode_tensor_map_bf16(&map_w, p->w, 3, dims, strides, boxw, false);
        if (rc) return rc;
        a.w_tiled = 1;
        a.w_bulk = static_cast<const uint8_t*>(p->w);
    } else if (p->w_tiled) {
        // [N tile][k block = tap * chunks + chunk][bnt rows][64]: channels of every tap zero-padded to whole 64-blocks at pack time
        if ((p->w_tiled >= 1000 ? p->w_tiled - 1000 : p->w_tiled) != bnt) { set_last_error("gmd_conv_fwd: tiled weights were packed for N tile %d, kernel picks %d", p->w_tiled, bnt); return kErrInvalid; }
        const uint64_t total_rows = (uint64_t)((p->Cout + bnt - 1) / bnt) * a.num_kb * bnt;
        uint64_t dims[3] = {BK, total_rows, 1};
        uint64_t strides[3] = {2, BK * 2, total_rows * BK * 2};
        uint32_t boxw[3] = {BK, (uint32_t)bnt, 1};
        int rc = encode_tensor_map_bf16(&map_w, p->w, 3, dims, strides, boxw, true);
        if (rc) return rc;
        a.w_tiled = 1;
        a.w_bulk = p->w_tiled >= 1000 ? static_cast<const uint8_t*>(p->w) : nullptr;
    } else {
        uint64_t dims[3] = {(uint64_t)taps * ctot, (uint64_t)rows_w, 1};
        uint64_t strides[3] = {2, (uint64_t)taps * ctot * 2, (uint64_t)taps * ctot * rows_w * 2};
        uint32_t boxw[3] = {BK, (uint32_t)bnt, 1};
        int rc = encode_tensor_map_bf16(&map_w, p->w, 3, dims, strides, boxw, true);
        if (rc) return rc;
    }
    int tiles_img = (p->N + bn - 1) / bn;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t tiles_m = (int64_t)a.tiles_w * a.tiles_h * tiles_img, tiles_nn = (p->Cout + bnt - 1) / bnt;
    const int64_t rows = (int64_t)p->N * Ho * Wo;
    const int ks = p->workspace ? conv_splitk_rule(Ho, Wo, a.num_kb, p->Cout, !p->upsample && p->stride == 1) : 1;
    if (p->gn_sums) {
        // statistics come out of full 128-row tiles of ONE image each, written by the single-pass epilogue
        if (ks > 1 || bn != 1 || (Wg % bw) || (Hg % bh) || (p->Cout % bnt) || (bnt % 32)) {
            set_last_error("gmd_conv_fwd: gn_sums is not available for this call (see gmd_conv_gn_sums_ok)"); return kErrUnsupported;
        }
        a.gn_sums = static_cast<long long*>(p->gn_sums); a.gn_ncb = p->Cout / 32;
    }
    if (ks > 1) {
        // the fp32 partial planes of the whole batch must fit the workspace; otherwise run the batch in image chunks (whole tiles),
        // which leaves every image's arithmetic unchanged
        const int64_t per_img = (int64_t)ks * Ho * Wo * p->Cout * 4;
        int64_t fit = p->workspace_bytes / per_img;
        if (fit < p->N) {
            fit = fit / bn * bn;
            if (fit < 1) { set_last_error("gmd_conv_fwd: split-K workspace (%lld B) cannot hold one tile of %d images", (long long)p->workspace_bytes, bn); return kErrInvalid; }
            for (int64_t n0 = 0; n0 < p->N; n0 += fit) {
                gmd_conv_params q = *p;
                q.N = (int32_t)((p->N - n0) < fit ? (p->N - n0) : fit);
                q.x0 = static_cast<const __nv_bfloat16*>(p->x0) + n0 * p->H * p->W * p->C0;
                if (p->x1) q.x1 = static_cast<const __nv_bfloat16*>(p->x1) + n0 * p->H * p->W * C1;
                const int64_t o = n0 * Ho * Wo * p->Cout;
                q.out = (p->flags & GMD_EPI_OUT_F32) ? static_cast<void*>(static_cast<float*>(p->out) + o) : static_cast<void*>(static_cast<__nv_bfloat16*>(p->out) + o);
                if (p->residual)
                    q.residual = (p->flags & GMD_EPI_RESIDUAL_F32) ? static_cast<const void*>(static_cast<const float*>(p->residual) + o)
                                                                   : static_cast<const void*>(static_cast<const __nv_bfloat16*>(p->residual) + o);
                if (p->row_bias) q.row_bias = p->row_bias + n0 * p->ld_row_bias;
                int rc = gmd_conv_fwd(&q, stream);
                if (rc) return rc;
            }
            return kOk;
        }
        KernelArgs b = a;
        b.ksplit = ks; b.kb_per_split = (a.num_kb + ks - 1) / ks; b.split_stride_o = rows * p->Cout;
        b.out = p->workspace; b.ldo = p->Cout; b.bias = nullptr; b.row_bias = nullptr; b.residual = nullptr; b.flags = GMD_EPI_OUT_F32;
        int rc = launch_cfg(bnt, tiles_m, tiles_nn, 1u, maps_a, map_w, b, st);
        if (rc) return rc;
        return launch_finalize(a, static_cast<const float*>(p->workspace), ks, rows, p->Cout, a.bias, a.row_bias, a.ld_row_bias, (int64_t)Ho * Wo,
                               a.residual, p->flags & GMD_EPI_RESIDUAL_F32, p->out, p->flags & GMD_EPI_OUT_F32, st);
    }
    return launch_cfg(bnt, tiles_m, tiles_nn, p->upsample ? 4u : 1u, maps_a, map_w, a, st, halo_ok, halo_ok && conv_pair, mtv);
}
