// Kernel (a): flash attention on tcgen05 + TMEM, fed by TMA.   O = softmax(Q K^T * scale) V
//
// attn_kernel<D, SHORT>: one CTA owns 128 query rows of one (batch, head).  Warp 0 = TMA producer (lane 0: Q then the K ring,
// lane 1: the V ring), warp 1 = MMA issuer, warps 2.. = softmax (one thread per query row = one TMEM lane).
//   S_j = Q K_j^T : tcgen05.mma  A = Q smem (K-major), B = K smem (K-major)     -> TMEM S[j&1]   (double buffered)
//   P_j = 2^(..)  : softmax warps read S with tcgen05.ld, write bf16 P to smem (K-major, 128B swizzle)
//   O  += P_j V_j : tcgen05.mma  A = P smem (K-major), B = V smem (MN-major)    -> TMEM O
// xattn_kernel<D>: text cross-attention (77 keys) with K / V resident and a loop over query tiles (second half of this file).
// Design points
//   * the softmax denominator comes out of the SAME MMA as O: column D of the V tile is overwritten with ones, so
//     O[:, D] = sum_k P[:, k] in fp32, rescaled together with O.
//   * O is rescaled in TMEM lazily: only when a row maximum grew by more than 2^8 (the stale maximum keeps every exponent <= 8,
//     well inside bf16 / fp32 range), so after the first tiles the correction path is almost never taken.
//   * d = 40 runs two and d = 80 three softmax warp sets on alternating whole tiles (Cfg::ALT), exponentiating against the stale
//     maximum; the sets' partial results are merged once at the end.
//   * K / V tiles are fetched through a dense (channel, token, batch) map: no out-of-bounds fill on the 80-byte head rows (launch()).
//   * what bounds d = 40: the TMEM read of S (128x64 fp32 per tile at 64 B/clk/SM = 512 cycles) and its 8192 ex2 (16/clk/SM = 512
//     cycles) — two co-equal floors; the tensor pipe is 25 % busy.
// Head dims 40/80/160 (SD1.5: 8 heads at every level) are zero-padded to a multiple of 16 for free by TMA on the Q side: that
// tensor map's innermost dim is the true head dim, the box is 64 wide.
#include "common.cuh"
#include <type_traits>
#include "../../include/gmd_b200.h"

namespace gmd {
void count_launch(int n);
namespace {

// Timing-only knock-outs (wrong results; never set in the product build) — how DESIGN.md §3a located the bounds of these kernels:
// build a second library with -DGMD_ATTN_KO=<bits> and run profiles/bench_attn.py against it (GMD_AB_LIB).
//   attn_kernel d = 40:  1 no MUFU (exponentials replaced by their argument), 2 no ones column / V wait in the softmax warps,
//                        4 no P stores, 8 no row maxima, 16 no pv_done wait, 32 no s_full wait, 64 producer ignores the ring's
//                        empty barriers, 128 MMA warp ignores s_free, 256 no K / V loads
//   xattn_kernel:        1 no MUFU, 4 no P stores, 8 no row maxima
#ifndef GMD_ATTN_KO
#define GMD_ATTN_KO 0
#endif
#ifndef GMD_XATTN_KO
#define GMD_XATTN_KO 0
#endif
#ifndef GMD_ATTN_SKIPV
#define GMD_ATTN_SKIPV 1
#endif
#ifndef GMD_ATTN_NSET40
#define GMD_ATTN_NSET40 2   // softmax warp sets at d = 40.  2: two CTAs per SM, 721 us at B=16 N=4096.  3 / 4: one CTA per SM, 871 / 872 us —
                            // the same for both, i.e. paced by the ONE MMA-issuing thread (~935 cycles per tile: four barrier polls, seven
                            // MMAs, four commits); with two CTAs per SM there are two of them
#endif
#ifndef GMD_ATTN_NSET80
#define GMD_ATTN_NSET80 3   // softmax warp sets at d = 80 (1 = the single-set, two-CTAs-per-SM configuration)
#endif
constexpr int BQ = 128;   // query rows per CTA
constexpr int BKV = 64;   // keys per tile (one 128-byte swizzle row of P)

struct AttnArgs {
    __nv_bfloat16* o;
    int64_t o_stride_b, o_stride_n, o_stride_h;
    int Nq, Nk;
    float scale_log2;
    int kv_dense;   // K / V tensor maps are 3-D (channel, token, batch): see launch()
};

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 2^x on the FMA / ALU pipes (no MUFU): x = n + f with n = round(x), f in [-0.5, 0.5]; degree-3 polynomial for 2^f (relative error
// 1.1e-4 — the result is rounded to bf16, 3.9e-3); the exponent is added to the bit pattern.  x <= 8 (lazy window) on this path.
__device__ __forceinline__ float exp2_poly(float x) {
    x = fmaxf(x, -125.0f);
    const float xf = x + 12582912.0f;                 // 1.5 * 2^23: the low mantissa bits now hold round(x)
    const float f = x - (xf - 12582912.0f);
    float p = fmaf(f, 0.0555041086f, 0.2402265069f);
    p = fmaf(p, f, 0.6931471806f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(xf) << 23));
}
#ifndef GMD_ATTN2_F32X2
#define GMD_ATTN2_F32X2 1     // scale-and-shift and the row sums of attn2_kernel as packed fp32 pairs (FFMA2 / FADD2: half the issue slots; 0 = scalar, A/B)
#endif
#ifndef GMD_ATTN2_POLYPAIR
#define GMD_ATTN2_POLYPAIR 1  // 1: the polynomial exponentials of attn2_kernel come in adjacent pairs (elements 2*POLY-2, 2*POLY-1 of every 2*POLY) on FFMA2 / FADD2
#endif
// exp2_poly for a register pair: the same arithmetic, five packed instructions + clamp and exponent insertion per lane.
__device__ __forceinline__ void exp2_poly2(uint64_t x2, float& p0, float& p1, float& xmax) {
    float x0, x1;
    f32x2_unpack(x2, x0, x1);
    xmax = fmaxf(xmax, fmaxf(x0, x1));   // (exponents >= 128 wrap around in the bit arithmetic below: the caller must notice them)
    x2 = f32x2_pack(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));
    const uint64_t xf = f32x2_add(x2, f32x2_pack(12582912.0f, 12582912.0f));
    const uint64_t rn = f32x2_add(xf, f32x2_pack(-12582912.0f, -12582912.0f));
    const uint64_t f = f32x2_fma(rn, f32x2_pack(-1.0f, -1.0f), x2);
    uint64_t p = f32x2_fma(f, f32x2_pack(0.0555041086f, 0.0555041086f), f32x2_pack(0.2402265069f, 0.2402265069f));
    p = f32x2_fma(p, f, f32x2_pack(0.6931471806f, 0.6931471806f));
    p = f32x2_fma(p, f, f32x2_pack(1.0f, 1.0f));
    float pa, pb, xa, xb;
    f32x2_unpack(p, pa, pb);
    f32x2_unpack(xf, xa, xb);
    p0 = __int_as_float(__float_as_int(pa) + (__float_as_int(xa) << 23));
    p1 = __int_as_float(__float_as_int(pb) + (__float_as_int(xb) << 23));
}
#ifndef GMD_ATTN2_SUMGROW
#define GMD_ATTN2_SUMGROW 1   // attn2_kernel detects a row maximum leaving the lazy window from the tile's row SUM (> 2^16) instead of forming the maximum of
                              // every tile: no FMNMX on the fast path; the true maximum is read back from S only on the rare rescale path (0: maximum per tile, A/B)
#endif
#ifndef GMD_ATTN2_UNROLL
#define GMD_ATTN2_UNROLL 1    // unroll factor of the softmax warps' tile loop
#endif
#ifndef GMD_ATTN2_POLY
#define GMD_ATTN2_POLY 0      // every POLY-th exponential of attn2_kernel on the FMA pipe (0 = all on MUFU)
#endif
#ifdef GMD_ATTN2_TRACE
#ifndef GMD_ATTN2_TRACE_Z
#define GMD_ATTN2_TRACE_Z 0   // the traced CTA is (0, 0, z)
#endif
// debug build only (profiles/trace_attn.py): clock64 timestamps of CTA (0,0,0), 8 event slots per key tile
__device__ long long g_attn_trace[8 * 1024];
#define TRACE(slot, j) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == GMD_ATTN2_TRACE_Z && (j) < 1024) g_attn_trace[(j) * 8 + (slot)] = clock64(); } while (0)
// rows 1000 / 1001: (clock64, globaltimer ns) at the first and after the last tile -> the SM clock the kernel actually ran at
#define TRACE_CLK(i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == GMD_ATTN2_TRACE_Z) { unsigned long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_)); \
        g_attn_trace[(1000 + (i)) * 8] = clock64(); g_attn_trace[(1000 + (i)) * 8 + 1] = (long long)gt_; } } while (0)
#else
#define TRACE(slot, j) do { } while (0)
#define TRACE_CLK(i) do { } while (0)
#endif
#ifndef GMD_ATTN2_POLY40
#define GMD_ATTN2_POLY40 4     // the d = 40 default: a quarter of the exponentials on the FMA pipe
#endif
#ifndef GMD_ATTN2_POLY80
#define GMD_ATTN2_POLY80 4     // d = 80 likewise (B=16 N=1024: 67.6 -> 66.5 us, B=8: 41.0 -> 38.9 us)
#endif
#ifndef GMD_ATTN2_ALIAS80
#define GMD_ATTN2_ALIAS80 1    // d = 80 with P stored over S like d = 40 (B=16 N=1024: 67.6 -> 65.6 us; 0: separate single P buffer, A/B)
#endif
#ifndef GMD_ATTN2_KO
#define GMD_ATTN2_KO 0        // timing-only knock-outs of attn2_kernel (wrong results): 1 no MUFU, 2 no P stores, 4 no S loads, 8 no row sums / maxima, 16 no K / V traffic after the first ring fill
#endif

template <int D, bool SHORT = false>
struct Cfg {
    static constexpr int NDB = (D + 63) / 64;               // 64-wide d blocks
    static constexpr int DP = (D + 15) / 16 * 16;           // K extent of Q K^T
    static constexpr int DPV = (D + 1 + 15) / 16 * 16;      // N extent of P V: head dim + the ones column
    static_assert(DPV <= NDB * 64, "ones column must fall inside the loaded V blocks");
    // SHORT (key counts of at most two tiles that the text cross-attention kernel does not take): the whole CTA lives for ~2 tiles, so
    // what matters is how many CTAs are resident to overlap their prologues (TMEM allocation, Q / K / V round trip) — single S and P
    // buffers, one softmax set: 112 TMEM columns and ~64 KB of shared memory per CTA at d = 40.
    // NSET (d = 40: 2 sets, two CTAs per SM; d = 80: 3 sets, one CTA per SM): independent softmax warp sets of 4 warps, each with its
    // own running maximum, O accumulator and denominator (a flash-decoding style split over keys, merged once at the end): one
    // softmax warp per sub-partition per CTA was latency-bound (ncu: XU pipe 52 % busy, issue slots 39 %).
    static constexpr int NSET = SHORT ? 1 : (D == 40 ? GMD_ATTN_NSET40 : D == 80 ? GMD_ATTN_NSET80 : 1);
    // ALT: the sets own ALTERNATING WHOLE TILES (set s: tiles j = s mod NSET, S buffer s, P buffer s, O accumulator s) rather than
    // slices of every tile's keys: per synchronisation round a thread handles 64 keys and the sets of a CTA run out of phase.
    // A tile is exponentiated against the running (stale) maximum while its own maximum is formed on the side; growth beyond
    // the lazy window (rare after a set's first tile, which takes a maximum-only pass first) redoes the tile from TMEM.
    static constexpr bool ALT = NSET >= 2;
    static constexpr int SB = SHORT ? 1 : (NSET > 2 ? NSET : 2);   // S buffers in TMEM (ALT: one per set)
    // K / V ring depth: one slot per tile in flight (measured at d = 40: deeper rings change nothing — the kernel does not wait for K / V)
    static constexpr int KS = NSET > 2 ? NSET : 2, VS = KS;
    static constexpr int KW = BKV;                          // keys per softmax thread per tile
    static constexpr int NB = ALT ? NSET : 2;               // p_full / pv_done / s_free barriers (tile j uses slot j % NB)
    static constexpr int THREADS = 64 + 128 * NSET;
    static constexpr int PB = ALT ? NSET : ((D == 80 || SHORT) ? 1 : 2);   // P buffers in smem (ALT: one per set)
    static constexpr int Q_BYTES = NDB * BQ * 128;
    static constexpr int KV_BLOCK_BYTES = BKV * 128;        // one d block of a K or V tile
    static constexpr int K_BYTES = NDB * KV_BLOCK_BYTES;
    static constexpr int P_BYTES = BQ * 128;
    static constexpr int OFF_K = Q_BYTES;
    static constexpr int OFF_V = OFF_K + KS * K_BYTES;
    static constexpr int OFF_P = OFF_V + VS * K_BYTES;
    static constexpr int OFF_BAR = OFF_P + PB * P_BYTES;
    static constexpr int SMEM = OFF_BAR + 256 + 1024;
    static constexpr uint32_t TMEM_COLS = (SB * BKV + NSET * DPV) <= 128 ? 128 : (SB * BKV + NSET * DPV) <= 256 ? 256 : 512;
    static_assert(NSET == 1 || OFF_P >= (NSET - 1) * BQ * (DPV + 1) * 4, "merge scratch must fit in the (then idle) Q / K / V buffers");
    static constexpr int MIN_CTAS = SHORT ? (D == 40 ? 3 : D == 80 ? 2 : 1) : ((D <= 80 && TMEM_COLS <= 256) ? 2 : 1);
    static constexpr float LAZY_T = 8.0f;   // lazy-rescale window of the running maximum, in log2 units
    // Measured and not kept (numbers at B=16, N=4096, d=40 unless noted; DESIGN.md §3a has the full list): the next tile's TMEM read in
    // flight under the exponentials (serialises with the same sub-partition's MUFU stream: 993 us vs 785 without, both sets variants),
    // 8- / 16-column TMEM reads (756 / 747 vs 721 us), packed ex2.approx.bf16x2 (915 us: its narrow lazy window fires the redo
    // path), a degree-3 polynomial exp2 on the FMA pipe for 1/4, 1/3, 1/2 of the pairs (724 / 766 / 800 vs 731 us: issue-bound),
    // S_{j+3} queued ahead of P V_j, P chunks held across the pv_done wait, three CTAs per SM with single-buffered S.
};

template <int D, bool SHORT>
__global__ void __launch_bounds__(Cfg<D, SHORT>::THREADS, Cfg<D, SHORT>::MIN_CTAS)
attn_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
            const __grid_constant__ CUtensorMap map_v, const AttnArgs args) {
    using C = Cfg<D, SHORT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* q_smem = smem;
    uint8_t* k_smem = smem + C::OFF_K;
    uint8_t* v_smem = smem + C::OFF_V;
    uint8_t* p_smem = smem + C::OFF_P;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* q_full = bars;            // 1
    uint64_t* k_full = bars + 1;        // KS
    uint64_t* k_empty = k_full + C::KS; // KS
    uint64_t* v_full = k_empty + C::KS; // VS
    uint64_t* v_empty = v_full + C::VS; // VS
    constexpr int NB = C::NB;
    uint64_t* s_full = v_empty + C::VS; // NB (SB used)
    uint64_t* p_full = s_full + NB;     // NB
    uint64_t* pv_done = p_full + NB;    // NB
    uint64_t* s_free = pv_done + NB;    // NB (ALT only: S buffer drained by its softmax set)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + NB);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BQ, head = blockIdx.y, batch = blockIdx.z;
    const int T = (args.Nk + BKV - 1) / BKV;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
        mbar_init(q_full, 1);
        for (int s = 0; s < C::KS; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
        for (int s = 0; s < C::VS; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
        for (int s = 0; s < NB; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 128); mbar_init(&pv_done[s], 1); mbar_init(&s_free[s], 128); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_o = tmem_base + C::SB * BKV;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(q_full, C::Q_BYTES);
            for (int b = 0; b < C::NDB; ++b) tma_load_4d(q_smem + b * BQ * 128, &map_q, q_full, b * 64, head, q0, batch);
            for (int j = 0; j < T; ++j) {
                const int st = j % C::KS;
                if (!(GMD_ATTN_KO & 64)) mbar_wait(&k_empty[st], ((j / C::KS) & 1) ^ 1);
                if (GMD_ATTN_KO & 256) { mbar_arrive(&k_full[st]); continue; }
                mbar_expect_tx(&k_full[st], C::K_BYTES);
                for (int b = 0; b < C::NDB; ++b) {
                    if (args.kv_dense) tma_load_3d(k_smem + st * C::K_BYTES + b * C::KV_BLOCK_BYTES, &map_k, &k_full[st], head * D + b * 64, j * BKV, batch);
                    else tma_load_4d(k_smem + st * C::K_BYTES + b * C::KV_BLOCK_BYTES, &map_k, &k_full[st], b * 64, head, j * BKV, batch);
                }
            }
        } else if (lane == 1) {
            for (int j = 0; j < T; ++j) {
                const int st = j % C::VS;
                if (!(GMD_ATTN_KO & 64)) mbar_wait(&v_empty[st], ((j / C::VS) & 1) ^ 1);
                if (GMD_ATTN_KO & 256) { mbar_arrive(&v_full[st]); continue; }
                mbar_expect_tx(&v_full[st], C::K_BYTES);
                for (int b = 0; b < C::NDB; ++b) {
                    if (args.kv_dense) tma_load_3d(v_smem + st * C::K_BYTES + b * C::KV_BLOCK_BYTES, &map_v, &v_full[st], head * D + b * 64, j * BKV, batch);
                    else tma_load_4d(v_smem + st * C::K_BYTES + b * C::KV_BLOCK_BYTES, &map_v, &v_full[st], b * 64, head, j * BKV, batch);
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t IDESC_S = umma_idesc_bf16(BQ, BKV, false, false);
            constexpr uint32_t IDESC_O = umma_idesc_bf16(BQ, C::DPV, false, true);  // B = V is MN-major
            const uint32_t q_addr = smem_u32(q_smem);
            auto issue_s = [&](int j) {
                const int st = j % C::KS;
                mbar_wait(&k_full[st], (j / C::KS) & 1);
                tc_fence_after();
                const uint32_t k_addr = smem_u32(k_smem + st * C::K_BYTES);
#pragma unroll
                for (int ks = 0; ks < C::DP / 16; ++ks) {
                    const int blk = ks >> 2, within = ks & 3;
                    uint64_t da = umma_desc_k_sw128(q_addr + blk * (BQ * 128) + within * 32);
                    uint64_t db = umma_desc_k_sw128(k_addr + blk * C::KV_BLOCK_BYTES + within * 32);
                    umma_bf16_ss(tmem_base + (j % C::SB) * BKV, da, db, IDESC_S, ks != 0 ? 1u : 0u);
                }
                umma_commit(&k_empty[st]);
                umma_commit(&s_full[j % C::SB]);
            };
            mbar_wait(q_full, 0);
            for (int i = 0; i < C::SB && i < T; ++i) issue_s(i);
            for (int j = 0; j < T; ++j) {
                const int st = j % C::VS;
                const int sl = j % NB;                      // barrier slot (ALT: = owning set)
                const uint32_t sl_ph = (j / NB) & 1;
                if constexpr (C::ALT) {
                    // S_{j+NSET} reuses the set's buffer as soon as the set has taken the logits of tile j out of TMEM, ahead of P V_j
                    if (j + NB < T) {
                        if (!(GMD_ATTN_KO & 128)) mbar_wait(&s_free[sl], sl_ph);
                        tc_fence_after();
                        issue_s(j + NB);
                    }
                }
                mbar_wait(&p_full[sl], sl_ph);   // softmax_j: P_j in smem, ones column set (non-ALT: S[j&1] drained)
                // (ALT: the owning set polled v_full before it wrote the ones column and arrived on p_full — no second poll here; every
                // poll costs this thread ~90 cycles and it paces the kernel, see GMD_ATTN_NSET40)
                if (!(C::ALT && GMD_ATTN_SKIPV)) mbar_wait(&v_full[st], (j / C::VS) & 1);
                tc_fence_after();
                const uint32_t v_addr = smem_u32(v_smem + st * C::K_BYTES);
                const uint32_t p_addr = smem_u32(p_smem + (j % C::PB) * C::P_BYTES);
#pragma unroll
                for (int ks = 0; ks < BKV / 16; ++ks) {
                    // key half ks / (KW/16) accumulates into its own O (NSET = 2), else everything into one
                    constexpr int KSTEPS = C::KW / 16;
                    uint64_t da = umma_desc_k_sw128(p_addr + ks * 32);
                    uint64_t db = umma_desc_mn_sw128(v_addr + ks * 16 * 128, C::KV_BLOCK_BYTES);
                    if constexpr (C::ALT)   // whole tile into the accumulator of the set that owns it
                        umma_bf16_ss(tmem_o + sl * C::DPV, da, db, IDESC_O, (j >= NB || ks != 0) ? 1u : 0u);
                    else
                        umma_bf16_ss(tmem_o + (ks / KSTEPS) * C::DPV, da, db, IDESC_O, (j != 0 || (ks % KSTEPS) != 0) ? 1u : 0u);
                }
                umma_commit(&v_empty[st]);
                umma_commit(&pv_done[sl]);
                if (!C::ALT && j + C::SB < T) issue_s(j + C::SB);   // S[j % SB] was drained by softmax_j (p_full_j)
            }
        }
    } else {
        constexpr int KW = C::KW;
        const int lg = warp & 3;
        const int set = (warp - 2) >> 2;                      // which key half this warp set owns (NSET = 2), else 0
        const int row = lg * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(lg * 32) << 16;
        const uint32_t my_o = tmem_o + set * C::DPV + lane_off;
        const float c = args.scale_log2;
        float m = -INFINITY;   // running (possibly stale) row maximum of THIS key half in scaled log2 units
        // One key tile of the single-set configurations (d = 80 / 160 and the short text cross-attention)
        auto tile = [&](int j, uint32_t (&sr)[KW]) {
            mbar_wait(&s_full[j % C::SB], (j / C::SB) & 1);
            tc_fence_after();
            {
                const uint32_t a = tmem_base + (j % C::SB) * BKV + lane_off + set * KW;
                if constexpr (KW == 64) tmem_ld_32x64(a, sr); else tmem_ld_32x32(a, sr);
            }
            tmem_wait_ld();
            const int valid = args.Nk - j * BKV - set * KW;  // columns >= valid are padding keys (last tile only)
            if (valid < KW) {
#pragma unroll
                for (int k = 0; k < KW; ++k) if (k >= valid) sr[k] = 0xff800000u;  // -inf
            }
            // four independent running maxima: a short dependent chain
            float mx0 = __uint_as_float(sr[0]), mx1 = __uint_as_float(sr[1]), mx2 = __uint_as_float(sr[2]), mx3 = __uint_as_float(sr[3]);
#pragma unroll
            for (int k = 4; k < KW; k += 4) {
                mx0 = fmaxf(mx0, __uint_as_float(sr[k])); mx1 = fmaxf(mx1, __uint_as_float(sr[k + 1]));
                mx2 = fmaxf(mx2, __uint_as_float(sr[k + 2])); mx3 = fmaxf(mx3, __uint_as_float(sr[k + 3]));
            }
            const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            const float mt = mx * c;
            const bool grow = mt > m + C::LAZY_T;      // lazy: keep the stale maximum while exponents stay <= LAZY_T
            const float m_new = grow ? mt : m;
            const float alpha = grow ? ex2(m - m_new) : 1.0f;   // first tile: m = -inf -> 0
            const float m_sub = m_new == -INFINITY ? 0.0f : m_new;   // a key half that has seen no valid key yet: p = 2^-inf = 0
            // the P buffer we are about to overwrite was last read by PV_{j-PB} (long complete: two tiles ago)
            if (j >= C::PB) {
                const int jj = j - C::PB;
                mbar_wait(&pv_done[jj & 1], (jj >> 1) & 1);
            }
            // P row -> smem, K-major SWIZZLE_128B: 16-byte chunk c of row r lands at chunk (c ^ (r & 7)); every chunk is stored as
            // soon as its eight probabilities exist, so at most four packed registers are live next to the in-flight S_{j+1}
            uint8_t* prow = p_smem + (j % C::PB) * C::P_BYTES + row * 128;
            uint32_t pk[4];
#pragma unroll
            for (int k = 0; k < KW / 2; ++k) {
                const float x0 = fmaf(__uint_as_float(sr[2 * k]), c, -m_sub), x1 = fmaf(__uint_as_float(sr[2 * k + 1]), c, -m_sub);
                pk[k & 3] = pack_bf16x2(ex2(x0), ex2(x1));
                if ((k & 3) == 3) {
                    const int cc = set * (KW / 8) + (k >> 2);
                    *reinterpret_cast<uint4*>(prow + ((cc ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            }
            m = m_new;
            // ones column of V_j (column D of the zero padding) -> the P V MMA also accumulates the softmax denominator
            if (set == 0) {
                const int st = j % C::VS;
                mbar_wait(&v_full[st], (j / C::VS) & 1);
                if (row < BKV) {
                    constexpr int blk = D / 64, cc = (D % 64) / 8, within = (D % 8) * 2;
                    uint8_t* vrow = v_smem + st * C::K_BYTES + blk * C::KV_BLOCK_BYTES + row * 128;
                    *reinterpret_cast<uint16_t*>(vrow + ((cc ^ (row & 7)) << 4) + within) = 0x3F80;  // bf16 1.0
                }
            }
            if (j > 0 && __any_sync(0xffffffffu, grow)) {
                mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);   // O must be complete up to tile j-1
                tc_fence_after();
#pragma unroll
                for (int ch = 0; ch < C::DPV / 16; ++ch) {
                    uint32_t o[16];
                    tmem_ld_32x16(my_o + ch * 16, o);
                    tmem_wait_ld();
#pragma unroll
                    for (int k = 0; k < 16; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
                    tmem_st_32x16(my_o + ch * 16, o);
                }
                tmem_wait_st();
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(&p_full[j & 1]);
        };
        int n_own = 0;   // ALT: tiles this set has processed
        if constexpr (C::ALT) {
            const uint32_t s_addr = tmem_base + set * BKV + lane_off;
            uint8_t* prow = p_smem + set * C::P_BYTES + row * 128;
            // One pass over the tile in two 32-column TMEM reads: row maximum on the side and, with `do_exp`, the probabilities
            // 2^(s*c - m_use) as bf16 into the P row (K-major SWIZZLE_128B: 16-byte chunk cc of row r lands at chunk cc ^ (r & 7)),
            // every chunk stored as soon as its eight values exist.  `pv_parity` >= 0: the P V MMA that last read this P buffer is
            // awaited just before the first store.  (volatile ex2: keeps the exponentials in program order behind their TMEM read.)
            auto ex2v = [](float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; };
            auto tile_pass = [&](int valid, float m_use, int pv_parity, bool do_exp, float& mx_out) {
                constexpr int CH = 32, NCH = BKV / CH;
                uint32_t a[CH], b[CH];
                float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
                for (int cidx = 0; cidx < NCH; ++cidx) {
                    uint32_t (&cur)[CH] = (cidx & 1) ? b : a;
                    tmem_ld_32x32(s_addr + cidx * CH, cur);
                    tmem_wait_ld();
                    if (valid < (cidx + 1) * CH) {   // padding keys (last tile only)
#pragma unroll
                        for (int k = 0; k < CH; ++k) if (cidx * CH + k >= valid) cur[k] = 0xff800000u;  // -inf
                    }
#pragma unroll
                    for (int k = 0; k < CH; k += 2) {
                        mx0 = fmaxf(mx0, __uint_as_float(cur[k])); mx1 = fmaxf(mx1, __uint_as_float(cur[k + 1]));
                    }
                    if (do_exp) {
#pragma unroll
                        for (int q8 = 0; q8 < CH / 8; ++q8) {
                            uint32_t pk[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float x0 = fmaf(__uint_as_float(cur[q8 * 8 + 2 * k]), c, -m_use), x1 = fmaf(__uint_as_float(cur[q8 * 8 + 2 * k + 1]), c, -m_use);
                                pk[k] = (GMD_ATTN_KO & 1) ? pack_bf16x2(x0, x1) : pack_bf16x2(ex2v(x0), ex2v(x1));
                            }
                            if (cidx == 0 && q8 == 0 && pv_parity >= 0 && !(GMD_ATTN_KO & 16)) mbar_wait(&pv_done[set], pv_parity);
                            const int cc = cidx * (CH / 8) + q8;
                            if ((GMD_ATTN_KO & 4) && pk[0] != 0x12345678u) continue;
                            *reinterpret_cast<uint4*>(prow + ((cc ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        }
                    }
                }
                mx_out = fmaxf(mx0, mx1);
            };
            for (int j = set; j < T; j += C::NSET, ++n_own) {
                const int valid = args.Nk - j * BKV;
                if (!(GMD_ATTN_KO & 32)) mbar_wait(&s_full[set], n_own & 1);
                tc_fence_after();
                float mx;
                if (n_own == 0) {   // the set's first tile: maximum first
                    tile_pass(valid, 0.0f, -1, false, mx);
                    m = mx * c;
                    tile_pass(valid, m, -1, true, mx);
                } else {
                    tile_pass(valid, m, (n_own - 1) & 1, true, mx);
                    const float mt = mx * c;
                    const bool grow = mt > m + C::LAZY_T;
                    if (__any_sync(0xffffffffu, grow)) {   // rare: a row maximum left the lazy window -> new maximum, tile again, O rescaled
                        const float m_new = grow ? mt : m;
                        const float alpha = grow ? ex2(m - m_new) : 1.0f;
                        m = m_new;
                        tile_pass(valid, m, -1, true, mx);
                        tc_fence_after();   // (pv_done of the previous tile was awaited before the first P store)
#pragma unroll
                        for (int ch = 0; ch < C::DPV / 16; ++ch) {
                            uint32_t o[16];
                            tmem_ld_32x16(my_o + ch * 16, o);
                            tmem_wait_ld();
#pragma unroll
                            for (int k = 0; k < 16; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
                            tmem_st_32x16(my_o + ch * 16, o);
                        }
                        tmem_wait_st();
                    }
                }
                tc_fence_before();
                mbar_arrive(&s_free[set]);   // the tile is out of TMEM: S_{j+2} may overwrite the buffer
                // ones column of V_j (column D of its padding) -> the P V MMA also accumulates the softmax denominator
                if (!(GMD_ATTN_KO & 2)) {
                    const int st = j % C::VS;
                    mbar_wait(&v_full[st], (j / C::VS) & 1);
                    if (row < BKV) {
                        constexpr int blk = D / 64, cc = (D % 64) / 8, within = (D % 8) * 2;
                        uint8_t* vrow = v_smem + st * C::K_BYTES + blk * C::KV_BLOCK_BYTES + row * 128;
                        *reinterpret_cast<uint16_t*>(vrow + ((cc ^ (row & 7)) << 4) + within) = 0x3F80;  // bf16 1.0
                    }
                }
                fence_proxy_async_smem();
                tc_fence_before();
                mbar_arrive(&p_full[set]);
            }
        } else {
            uint32_t sa[KW], sb[KW];
            for (int j = 0; j < T; j += 2) {
                tile(j, sa);
                if (j + 1 < T) tile(j + 1, sb);
            }
        }
        if constexpr (C::ALT) {
            // every set waits for ITS last P V (the phase it has been tracking); after the named barrier both accumulators are final
            // and the P buffers (the merge scratch) are no longer read by the tensor core
            mbar_wait(&pv_done[set], (n_own - 1) & 1);
            asm volatile("bar.sync 2, %0;" ::"n"(128 * C::NSET) : "memory");
        } else {
            mbar_wait(&pv_done[(T - 1) & 1], ((T - 1) >> 1) & 1);
        }
        tc_fence_after();
        const int q = q0 + row;
        __nv_bfloat16* op = args.o + batch * args.o_stride_b + (int64_t)q * args.o_stride_n + head * args.o_stride_h;
        if constexpr (C::NSET == 1) {
            float inv_l;
            {
                // denominator accumulated by the MMA through the ones column (column D of O)
                uint32_t o[16];
                tmem_ld_32x16(my_o + (D / 16) * 16, o);
                tmem_wait_ld();
                inv_l = 1.0f / __uint_as_float(o[D % 16]);
            }
#pragma unroll
            for (int ch = 0; ch < (D + 15) / 16; ++ch) {
                uint32_t o[16];
                tmem_ld_32x16(my_o + ch * 16, o);
                tmem_wait_ld();
                if (q < args.Nq) {
#pragma unroll
                    for (int h8 = 0; h8 < 2; ++h8) {
                        const int d0 = ch * 16 + h8 * 8;
                        if (d0 < D) {  // D is a multiple of 8
                            uint4 v = make_uint4(pack_bf16x2(__uint_as_float(o[h8 * 8 + 0]) * inv_l, __uint_as_float(o[h8 * 8 + 1]) * inv_l),
                                                 pack_bf16x2(__uint_as_float(o[h8 * 8 + 2]) * inv_l, __uint_as_float(o[h8 * 8 + 3]) * inv_l),
                                                 pack_bf16x2(__uint_as_float(o[h8 * 8 + 4]) * inv_l, __uint_as_float(o[h8 * 8 + 5]) * inv_l),
                                                 pack_bf16x2(__uint_as_float(o[h8 * 8 + 6]) * inv_l, __uint_as_float(o[h8 * 8 + 7]) * inv_l));
                            *reinterpret_cast<uint4*>(op + d0) = v;
                        }
                    }
                }
            }
        } else {
            // merge the sets' partial results: O = sum_s O_s 2^(m_s - m) / sum_s l_s 2^(m_s - m), m = max_s m_s.
            // Sets 1.. park (m_s, O_s[0..D]) in the (now idle) Q / K / V buffers, set 0 combines and writes the output row.
            float ov[C::DPV];
#pragma unroll
            for (int ch = 0; ch < C::DPV / 16; ++ch) {
                uint32_t o[16];
                tmem_ld_32x16(my_o + ch * 16, o);
                tmem_wait_ld();
#pragma unroll
                for (int k = 0; k < 16; ++k) ov[ch * 16 + k] = __uint_as_float(o[k]);
            }
            constexpr int PITCH = C::DPV + 1;   // odd pitch: conflict-free rows
            float* scratch = reinterpret_cast<float*>(smem);
            if (set > 0) {
                float* mine = scratch + ((set - 1) * BQ + row) * PITCH;
                mine[C::DPV] = m;   // -inf for a set that owned no tile (Nk <= 64 * (NSET - 1)): its accumulator was never written
#pragma unroll
                for (int k = 0; k <= D; ++k) mine[k] = n_own > 0 ? ov[k] : 0.0f;
            }
            asm volatile("bar.sync 2, %0;" ::"n"(128 * C::NSET) : "memory");
            if (set == 0) {
                float mm = m;
#pragma unroll
                for (int t = 0; t < C::NSET - 1; ++t) mm = fmaxf(mm, scratch[(t * BQ + row) * PITCH + C::DPV]);
                float w[C::NSET];
                w[0] = ex2(m - mm);
                float l = ov[D] * w[0];
#pragma unroll
                for (int t = 0; t < C::NSET - 1; ++t) {
                    const float* other = scratch + (t * BQ + row) * PITCH;
                    const float mo = other[C::DPV];
                    w[t + 1] = mo == -INFINITY ? 0.0f : ex2(mo - mm);
                    l += other[D] * w[t + 1];
                }
                const float inv_l = 1.0f / l;
                if (q < args.Nq) {
#pragma unroll
                    for (int d0 = 0; d0 < D; d0 += 8) {
                        float r[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            float acc = ov[d0 + k] * w[0];
#pragma unroll
                            for (int t = 0; t < C::NSET - 1; ++t) acc += scratch[(t * BQ + row) * PITCH + d0 + k] * w[t + 1];
                            r[k] = acc * inv_l;
                        }
                        *reinterpret_cast<uint4*>(op + d0) = make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]), pack_bf16x2(r[4], r[5]), pack_bf16x2(r[6], r[7]));
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<C::TMEM_COLS>(tmem_base);
    }
}

template <int D, bool SHORT>
int launch(const gmd_attn_params* p, cudaStream_t st) {
    using C = Cfg<D, SHORT>;
    static bool configured[kMaxDevices] = {};
    const int dev = device_ordinal();
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(attn_kernel<D, SHORT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) { set_last_error("attn: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return kErrCuda; }
        configured[dev] = true;
    }
    CUtensorMap mq, mk, mv;
    auto enc = [&](CUtensorMap* m, const void* base, int64_t sb, int64_t sn, int64_t sh, int n, uint32_t rows) {
        uint64_t dims[4] = {(uint64_t)D, (uint64_t)p->H, (uint64_t)n, (uint64_t)p->B};
        uint64_t strides[4] = {2, (uint64_t)sh * 2, (uint64_t)sn * 2, (uint64_t)sb * 2};
        uint32_t box[4] = {64, 1, rows, 1};
        return encode_tensor_map_bf16(m, base, 4, dims, strides, box, true);
    };
    int rc;
    if ((rc = enc(&mq, p->q, p->q_stride_b, p->q_stride_n, p->q_stride_h, p->Nq, BQ))) return rc;
    // K / V: when the heads of a token are contiguous (stride_h == d: the fused QKV / KV GEMM outputs) the map is 3-D over the whole
    // channel row and the 64-wide box starts at channel head*d: the columns past the head's d are the NEXT head's values (finite,
    // multiplied by Q's zero padding in Q K^T and landing in never-read columns of O in P V) instead of TMA out-of-bounds fill.
    // Measured: the zero-filled 80-byte rows of the 4-D (d, head, token, batch) map paced the whole kernel (ncu / knock-out runs:
    // 808 us with the softmax removed vs 483 us without the K / V loads, B=16 N=4096 d=40).
    const bool kv_dense = p->k_stride_h == D && p->v_stride_h == D;
    auto enc_dense = [&](CUtensorMap* m, const void* base, int64_t sb, int64_t sn, int n) {
        uint64_t dims[3] = {(uint64_t)D * p->H, (uint64_t)n, (uint64_t)p->B};
        uint64_t strides[3] = {2, (uint64_t)sn * 2, (uint64_t)sb * 2};
        uint32_t box[3] = {64, (uint32_t)BKV, 1};
        return encode_tensor_map_bf16(m, base, 3, dims, strides, box, true);
    };
    if (kv_dense) {
        if ((rc = enc_dense(&mk, p->k, p->k_stride_b, p->k_stride_n, p->Nk))) return rc;
        if ((rc = enc_dense(&mv, p->v, p->v_stride_b, p->v_stride_n, p->Nk))) return rc;
    } else {
        if ((rc = enc(&mk, p->k, p->k_stride_b, p->k_stride_n, p->k_stride_h, p->Nk, BKV))) return rc;
        if ((rc = enc(&mv, p->v, p->v_stride_b, p->v_stride_n, p->v_stride_h, p->Nk, BKV))) return rc;
    }
    AttnArgs a;
    a.o = static_cast<__nv_bfloat16*>(p->o);
    a.o_stride_b = p->o_stride_b; a.o_stride_n = p->o_stride_n; a.o_stride_h = p->o_stride_h;
    a.Nq = p->Nq; a.Nk = p->Nk;
    a.scale_log2 = p->scale * 1.4426950408889634f;
    a.kv_dense = kv_dense ? 1 : 0;
    dim3 grid((p->Nq + BQ - 1) / BQ, p->H, p->B);
    attn_kernel<D, SHORT><<<grid, C::THREADS, C::SMEM, st>>>(mq, mk, mv, a);
    count_launch(1);
    return check_launch("attn_kernel");
}


// ---------------------------------------------------------------------------------------------------------------------------------
// attn2_kernel<D, BKV, SB, PB, KS>: self-attention with P KEPT IN TENSOR MEMORY (round 2).
//
// What the round-1 kernel was actually bound by (profiles/ubench_tmem_r02.txt, ubench_mma_issue_r02.txt, measured on B200):
//   * reading a 128x64 fp32 S tile out of TMEM costs ~35 cycles per SM, not 512 — tcgen05.ld is NOT a floor; the only hard floor of
//     the softmax is MUFU: 8192 ex2 per tile at 16 / clk / SM = 512 cycles, and a bare ld -> max -> fma -> ex2 -> pack loop reaches
//     548 cycles per tile with two warps per sub-partition;
//   * an SS-form tcgen05.mma with N <= 64 costs ~46 cycles whatever N is: it is bound by the 6 KB of A + B operand bytes it pulls
//     through the 128 B/clk shared-memory port.  Per 64-key tile the v1 kernel moved 42 KB of MMA operands + 16 KB of P stores +
//     16 KB of TMA writes = 74 KB = 578 cycles of shared-memory bandwidth — MORE than the MUFU floor;
//   * each poll of an mbarrier costs the MMA-issuing thread ~86 cycles, a tcgen05.commit ~5.
// Hence this kernel:
//   * P = 2^(S*c - m) goes registers -> tcgen05.st -> TMEM and is the A operand of the P V MMA in its TS form: no P stores, no
//     fence.proxy.async and no A-operand reads on the shared-memory port (26-36 KB per 64 keys instead of 74);
//   * the softmax denominator is accumulated by the softmax threads (one FADD per element; issue slots are not the bound) instead
//     of a ones column written into the V tile, so the softmax warps never touch shared memory or wait for V;
//   * ONE softmax warp set per CTA and two CTAs per SM: no end-of-kernel merge of partial results, prologue / epilogue of one CTA
//     overlap the main loop of the other, two MMA-issuing threads per SM;
//   * p_full(j) — "P_j is in TMEM" — also means "S_j has been read": one barrier, one poll per tile for the MMA thread
//     (k_full for the next S, p_full and v_full for P V: three polls, four commits per tile);
//   * the running maximum is checked per 32-column chunk BEFORE the chunk is exponentiated (growth beyond the lazy window 2^8 takes a
//     rare slow path that rescales O, the denominator and the tile's earlier P chunks in TMEM), so S can be released as soon as it
//     has been read and is never read twice.
//   * NQT = 2 (d = 40): one CTA owns TWO 128-row query tiles ("chains"), each with its own softmax warp set, S / P columns and O
//     accumulator, sharing every K / V tile — four softmax warps per sub-partition with two CTAs per SM.  A softmax warp spends
//     ~600 cycles per tile on the MUFU pipe and about as long NOT on it (barrier polls, TMEM round trips, branches: ncu source
//     view, profiles/ncu_attn2_r02_summary.txt), so two warps per sub-partition leave the pipe ~35 % idle.  TMEM then only has room
//     for ALIAS: P_j overwrites S_j in place (stored after the whole tile has been read and the growth check has passed) and
//     S_{j+1} follows P V_j in the tensor pipe's issue order.
//   * HS = 2 (round 2, second pass; profiles/ubench_mma_ts_r02.txt): the TS-form P V MMA runs at the speed of its math (26 cycles
//     at N = 48) and the whole tile's MMAs need ~260 tensor-pipe cycles, so what is left is keeping the MUFU pipe fed.  Every query
//     row is worked on by TWO threads, each owning one half of the tile's keys as an independent online-softmax chain (own running
//     maximum, denominator and O accumulator: P V of keys 0-31 accumulates into O_a, of keys 32-63 into O_b), merged once at the
//     end of the kernel.  The halves never talk to each other per tile, a thread holds 32 instead of 64 scores (~96 registers), and
//     two CTAs per SM put FOUR light softmax warps on every sub-partition instead of two heavy ones.
//   * WG = 1 (128-key tiles, two chains, one CTA per SM): the register file belongs to the sub-partitions (16 K registers each), so a
//     10-warp CTA is capped at 168 registers per thread like a 12-warp one.  The CTA therefore has three whole warpgroups — {TMA warp,
//     MMA warp, two idle warps} and one per chain — and re-balances with setmaxnreg: the first gives up all but 40 registers, the
//     softmax warpgroups grow to 232, enough for a 128-column row with the next 32-column tcgen05.ld in flight under the current
//     chunk's exponentials.  Twice the keys per barrier hand-off (~850 cycles per tile whatever its size).
template <int D, int BKV_, int NQT_, bool ALIAS_, int SB_, int PB_, int KS_, int HS_ = 1, int WG_ = 0, int POLY_ = GMD_ATTN2_POLY>
struct Cfg2 {
    static constexpr int BKV = BKV_;                        // keys per tile
    static constexpr int NQT = NQT_;                        // 128-row query tiles (chains) per CTA
    static constexpr bool ALIAS = ALIAS_;                   // P_j is stored over S_j (by the thread that read those columns of S_j)
    static constexpr int SB = SB_, PB = ALIAS ? SB_ : PB_;  // S / P buffers per chain in TMEM (ALIAS: the S buffers are the P buffers)
    static constexpr int KS = KS_;                          // K and V ring depth (separate rings, separate barriers)
    static constexpr int HS = HS_;                          // threads per query row (key halves of a tile as independent chains)
    static constexpr int WG = WG_;                          // 1: whole warpgroups + setmaxnreg (see above)
    static constexpr int SW0 = WG ? 4 : 2;                  // first softmax warp
    static constexpr int POLY = POLY_;                      // every POLY-th exponential as an FMA-pipe polynomial instead of a MUFU op (0: none)
    static constexpr int NDB = (D + 63) / 64;               // 64-wide d blocks
    static constexpr int DP = (D + 15) / 16 * 16;           // K extent of Q K^T
    static constexpr int DPV = (D + 15) / 16 * 16;          // N extent of P V
    static constexpr int PCOLS = BKV / 2;                   // TMEM columns of one bf16 P tile (two keys per 32-bit cell)
    static constexpr int CH_COLS = ALIAS ? SB * BKV : SB * BKV + PB * PCOLS;   // S (+ P) columns of one chain
    static constexpr int TM_O = NQT * CH_COLS;
    static constexpr int TMEM_USED = TM_O + NQT * HS * DPV;
    static constexpr uint32_t TMEM_COLS = TMEM_USED <= 128 ? 128 : TMEM_USED <= 256 ? 256 : 512;
    static constexpr int MIN_CTAS = TMEM_COLS <= 128 ? 3 : TMEM_COLS <= 256 ? 2 : 1;   // (3, not 4: 192 threads x 3 leaves 112 registers per thread)
    static constexpr int QT_BYTES = NDB * BQ * 128;         // one query tile
    static constexpr int Q_BYTES = NQT * QT_BYTES;
    static constexpr int KV_BLOCK_BYTES = BKV * 128;        // one d block of a K or V tile
    static constexpr int K_BYTES = NDB * KV_BLOCK_BYTES;
    static constexpr int OFF_K = Q_BYTES;
    static constexpr int OFF_V = OFF_K + KS * K_BYTES;
    static constexpr int OFF_BAR = OFF_V + KS * K_BYTES;
    static constexpr int OFF_MERGE = OFF_BAR + 256;         // (m, l) of every chain half for the final merge (HS > 1)
    static constexpr int SMEM = OFF_MERGE + (HS > 1 ? NQT * HS * BQ * 8 : 0) + 1024;
    static constexpr int THREADS = SW0 * 32 + 128 * NQT * HS;   // warp 0: TMA, warp 1: MMA, (WG: two idle warps,) then 4 * HS softmax warps per chain
    static constexpr int NBAR = 1 + 4 * KS + NQT * (SB + 2 * PB);
    static constexpr float LAZY_T = 8.0f;
    static_assert(BKV % (32 * HS) == 0 && BKV <= 256, "tile");
    static_assert(KS >= SB, "S_{j+SB} needs its K tile while V_j is still in use");
    static_assert(SMEM * MIN_CTAS <= 227 * 1024, "shared memory");
    static_assert(NBAR * 8 + 4 <= 256, "barrier block");
};

// Waits of attn2_kernel.  A waiting warp is not free: every failed poll is an issue slot and an MIO-queue instruction taken from the
// softmax warps of its sub-partition (a tight inline-asm poll loop cost 6 % at d = 40, the loop with a clock64 watchdog between polls
// was 4 % FASTER with try_wait than with test_wait), so each role polls as rarely as its latency budget allows:
//   TMA lanes (ring slots: tiles of slack)  GMD_ATTN2_TMA_SLEEP ns between polls
//   MMA thread (V / K operands, P hand-off) GMD_ATTN2_MMA_SLEEP ns
//   softmax warps (S hand-off)              GMD_ATTN2_SM_SLEEP ns            (0: try_wait loop without a sleep)
#ifndef GMD_ATTN2_TESTWAIT
#define GMD_ATTN2_TESTWAIT 0   // 1: hand-offs between the MMA thread and the softmax warps poll with mbarrier.test_wait (A/B)
#endif
#ifndef GMD_ATTN2_TMA_SLEEP
#define GMD_ATTN2_TMA_SLEEP 0
#endif
#ifndef GMD_ATTN2_MMA_SLEEP
#define GMD_ATTN2_MMA_SLEEP 0
#endif
#ifndef GMD_ATTN2_SM_SLEEP
#define GMD_ATTN2_SM_SLEEP 0
#endif
template <int NS>
__device__ __forceinline__ void attn2_wait(uint64_t* bar, uint32_t parity) {
    if (NS > 0) mbar_wait_sleep(bar, parity, NS);
    else if (GMD_ATTN2_TESTWAIT) mbar_wait_poll(bar, parity);
    else mbar_wait(bar, parity);
}
#define AWAIT(bar, parity) attn2_wait<GMD_ATTN2_SM_SLEEP>(bar, parity)
#define MWAIT(bar, parity) attn2_wait<GMD_ATTN2_MMA_SLEEP>(bar, parity)
#define TWAIT(bar, parity) attn2_wait<GMD_ATTN2_TMA_SLEEP>(bar, parity)
template <int D, int BKV, int NQT, bool ALIAS, int SB_, int PB_, int KS, int HS, int WG, int POLY_>
__global__ void __launch_bounds__(Cfg2<D, BKV, NQT, ALIAS, SB_, PB_, KS, HS, WG, POLY_>::THREADS, Cfg2<D, BKV, NQT, ALIAS, SB_, PB_, KS, HS, WG, POLY_>::MIN_CTAS)
attn2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
             const __grid_constant__ CUtensorMap map_v, const AttnArgs args) {
    using C = Cfg2<D, BKV, NQT, ALIAS, SB_, PB_, KS, HS, WG, POLY_>;
    constexpr int SB = C::SB, PB = C::PB, SW0 = C::SW0, POLY = C::POLY;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* q_smem = smem;
    uint8_t* k_smem = smem + C::OFF_K;
    uint8_t* v_smem = smem + C::OFF_V;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* q_full = bars;             // 1
    uint64_t* k_full = bars + 1;         // KS
    uint64_t* k_empty = k_full + KS;     // KS
    uint64_t* v_full = k_empty + KS;     // KS
    uint64_t* v_empty = v_full + KS;     // KS
    uint64_t* s_full = v_empty + KS;     // [NQT][SB]
    uint64_t* p_full = s_full + NQT * SB;    // [NQT][PB]  (128 arrivals: P_j in TMEM, S_j read)
    uint64_t* pv_done = p_full + NQT * PB;   // [NQT][PB]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + NQT * PB);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * (BQ * NQT), head = blockIdx.y, batch = blockIdx.z;
    const int T = (args.Nk + BKV - 1) / BKV;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
        mbar_init(q_full, 1);
        for (int s = 0; s < KS; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
        for (int s = 0; s < NQT * SB; ++s) mbar_init(&s_full[s], 1);
        for (int s = 0; s < NQT * PB; ++s) { mbar_init(&p_full[s], 4 * HS); mbar_init(&pv_done[s], 1); }   // (one arrival per softmax warp)
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // TMEM columns of chain t: S buffers, then (unless aliased) P buffers; O accumulators of all chains behind
    auto tm_s = [&](int t, int sb) { return tmem_base + t * C::CH_COLS + sb * BKV; };
    // P of key half h: packed behind the S buffers, or (ALIAS) over the first half of the S columns that half's threads own
    auto tm_p = [&](int t, int pb, int h) {
        return ALIAS ? tmem_base + t * C::CH_COLS + pb * BKV + h * (BKV / HS) : tmem_base + t * C::CH_COLS + SB * BKV + pb * C::PCOLS + h * (C::PCOLS / HS);
    };
    auto tm_o = [&](int t, int h) { return tmem_base + C::TM_O + (t * HS + h) * C::DPV; };

    if (warp < SW0) {
    // (WG: the whole first warpgroup — TMA warp, MMA warp, two idle warps — hands its registers to the softmax warpgroups)
    if (WG) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(q_full, C::Q_BYTES);
            for (int t = 0; t < NQT; ++t)
                for (int b = 0; b < C::NDB; ++b) tma_load_4d(q_smem + t * C::QT_BYTES + b * BQ * 128, &map_q, q_full, b * 64, head, q0 + t * BQ, batch);
            // (the K and the V stream have their own lane: a K slot is free again as soon as S_j has read it, long before the V slot of
            // the same tile — one in-order stream would hold K_{j+KS} back behind V_{j+KS-1}, i.e. behind P V_{j-1})
            for (int j = 0; j < T; ++j) {
                const int st = j % KS;
                TWAIT(&k_empty[st], ((j / KS) & 1) ^ 1);
                if ((GMD_ATTN2_KO & 16) && j >= KS) { mbar_arrive(&k_full[st]); continue; }   // timing only: no K / V traffic after the first ring fill
                mbar_expect_tx(&k_full[st], C::K_BYTES);
                for (int b = 0; b < C::NDB; ++b) tma_load_3d(k_smem + st * C::K_BYTES + b * C::KV_BLOCK_BYTES, &map_k, &k_full[st], head * D + b * 64, j * BKV, batch);
                TRACE(0, j);
            }
        } else if (lane == 1) {
            for (int j = 0; j < T; ++j) {
                const int st = j % KS;
                TWAIT(&v_empty[st], ((j / KS) & 1) ^ 1);
                if ((GMD_ATTN2_KO & 16) && j >= KS) { mbar_arrive(&v_full[st]); continue; }
                mbar_expect_tx(&v_full[st], C::K_BYTES);
                for (int b = 0; b < C::NDB; ++b) tma_load_3d(v_smem + st * C::K_BYTES + b * C::KV_BLOCK_BYTES, &map_v, &v_full[st], head * D + b * 64, j * BKV, batch);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t IDESC_S = umma_idesc_bf16(BQ, BKV, false, false);
            constexpr uint32_t IDESC_O = umma_idesc_bf16(BQ, C::DPV, false, true);   // A = P from TMEM (K-major), B = V is MN-major
            const uint32_t q_addr = smem_u32(q_smem);
            // S_t(j) = Q_t K_j^T.  The K tile is polled once per j (chain 0) and released after the last chain's MMAs.
            auto issue_s = [&](int t, int j, bool polled) {
                const int st = j % KS;
                if (t == 0 && !polled) { MWAIT(&k_full[st], (j / KS) & 1); tc_fence_after(); TRACE(1, j); }
                const uint32_t k_addr = smem_u32(k_smem + st * C::K_BYTES);
#pragma unroll
                for (int ks = 0; ks < C::DP / 16; ++ks) {
                    const int blk = ks >> 2, within = ks & 3;
                    const uint64_t da = umma_desc_k_sw128(q_addr + t * C::QT_BYTES + blk * (BQ * 128) + within * 32);
                    const uint64_t db = umma_desc_k_sw128(k_addr + blk * C::KV_BLOCK_BYTES + within * 32);
                    umma_bf16_ss(tm_s(t, j % SB), da, db, IDESC_S, ks != 0 ? 1u : 0u);
                }
                umma_commit(&s_full[t * SB + j % SB]);
                if (t == NQT - 1) umma_commit(&k_empty[st]);
                if (t == 0) TRACE(2, j);
            };
            MWAIT(q_full, 0);
            for (int i = 0; i < SB && i < T; ++i)
                for (int t = 0; t < NQT; ++t) issue_s(t, i, false);
            for (int j = 0; j < T; ++j) {
                const int st = j % KS, pb = j % PB;
                const uint32_t v_addr = smem_u32(v_smem + st * C::K_BYTES);
                // The operands of this iteration first — V_j and K_{j+SB} have been in flight for whole tiles, and every poll costs this
                // thread 100-230 cycles while the softmax warps keep the sub-partition's MIO queue busy (clock64 trace) — so that nothing
                // but the MMA issue itself stands between "P_j is in TMEM" and S_{j+SB}, the tile the softmax warps will wait for.
                const bool next_s = j + SB < T;
                MWAIT(&v_full[st], (j / KS) & 1);
                TRACE(4, j);
                // (K_{j+SB} only where it has a ring slot of its own: with KS == SB its load starts when S_j completes and may still be in flight)
                constexpr bool EARLY_K = KS > SB;
                if (EARLY_K && next_s) { MWAIT(&k_full[(j + SB) % KS], ((j + SB) / KS) & 1); TRACE(1, j + SB); }
#pragma unroll
                for (int t = 0; t < NQT; ++t) {
                    MWAIT(&p_full[t * PB + pb], (j / PB) & 1);       // P_t(j) in TMEM; S_t(j) read (its buffer may be overwritten)
                    tc_fence_after();
                    if (t == 0) TRACE(3, j);
#pragma unroll
                    for (int ks = 0; ks < BKV / 16; ++ks) {
                        constexpr int KPH = BKV / 16 / HS;   // K steps per key half: half h accumulates into its own O
                        const uint64_t db = umma_desc_mn_sw128(v_addr + ks * 16 * 128, C::KV_BLOCK_BYTES);
                        umma_bf16_ts(tm_o(t, ks / KPH), tm_p(t, pb, ks / KPH) + (ks % KPH) * 8, db, IDESC_O, (j != 0 || (ks % KPH) != 0) ? 1u : 0u);
                    }
                    umma_commit(&pv_done[t * PB + pb]);
                    if (t == NQT - 1) umma_commit(&v_empty[st]);
                    if (t == 0) TRACE(5, j);
                    // the chain's next S right behind its P V (ALIAS: it overwrites P_t(j), after P V_t(j) in the pipe's issue order)
                    if (next_s) issue_s(t, j + SB, EARLY_K);
                }
            }
        }
    }
    } else {
        if (WG) asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        const int lg = warp & 3;                 // TMEM lane quadrant of this warp (fixed by warp % 4)
        const int hw = ((warp - SW0) >> 2) % HS;   // key half of the tile this thread owns
        const int t = (warp - SW0) / (4 * HS);     // chain (query tile) of this softmax warp
        const int row = lg * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(lg * 32) << 16;
        const uint32_t my_o = tm_o(t, hw) + lane_off;
        const float c = args.scale_log2;
        float m = -INFINITY;     // running (possibly stale) row maximum, scaled log2 units
        float l = 0.0f;          // running denominator, relative to m
        constexpr int NCH = BKV / 32 / HS;
        // One pass over the S tile in 32-column TMEM reads.  EXP: probabilities 2^(s*c - m_use) as packed bf16 into pk[] (stored to
        // TMEM by the caller once the growth check has passed), their fp32 sum into `lsum`; the row maximum of the raw logits is
        // formed on the side (never on the exponentials' dependency chain).  MASK: columns >= valid are padding keys (last tile of
        // a ragged key count only).  The exponentials are plain (non-volatile) asm behind tmem_wait_ld_regs, so ptxas interleaves the
        // 32 independent fma -> ex2 -> add / pack chains of a chunk freely (one MUFU every ~4 issue slots in the SASS).
        auto tile_pass = [&](auto exp_tag, auto mask_tag, uint32_t s_addr, int valid, float m_use, uint32_t* pk, float& mx_out, float& lsum, float& px_out) {
            constexpr bool EXP = decltype(exp_tag)::value, MASK = decltype(mask_tag)::value;
            float mx0 = -INFINITY, mx1 = -INFINITY, l0 = 0.0f, l1 = 0.0f, l2 = 0.0f, l3 = 0.0f;
            float px = -INFINITY;   // largest exponent that went through the polynomial (which wraps around beyond 2^127 instead of returning +inf)
#if GMD_ATTN2_F32X2
            uint64_t l01 = f32x2_pack(0.0f, 0.0f), l23 = l01;
            const uint64_t c2 = f32x2_pack(c, c), nm2 = f32x2_pack(-m_use, -m_use);
#endif
            // PIPE (one CTA per SM: registers to spare): the next chunk's tcgen05.ld is in flight while this chunk is exponentiated
            constexpr bool PIPE = WG && NCH > 1 && !(GMD_ATTN2_KO & 4);
            uint32_t rr[PIPE ? 2 : 1][32];
            if (PIPE) tmem_ld_32x32(s_addr, rr[0]);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                uint32_t (&r)[32] = rr[PIPE ? (ch & 1) : 0];
                if (GMD_ATTN2_KO & 4) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(m_use + k);
                } else if (PIPE) {
                    tmem_wait_ld_regs(r);
                    if (ch + 1 < NCH) tmem_ld_32x32(s_addr + (ch + 1) * 32, rr[(ch + 1) & 1]);
                } else {
                    tmem_ld_32x32(s_addr + ch * 32, r);
                    tmem_wait_ld_regs(r);
                }
                if (MASK) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) if (ch * 32 + k >= valid) r[k] = 0xff800000u;   // -inf
                }
                if (!EXP || !GMD_ATTN2_SUMGROW) {
#pragma unroll
                    for (int k = 0; k < ((GMD_ATTN2_KO & 8) ? 4 : 32); k += 4) {
                        mx0 = fmaxf(mx0, fmaxf(__uint_as_float(r[k]), __uint_as_float(r[k + 1])));
                        mx1 = fmaxf(mx1, fmaxf(__uint_as_float(r[k + 2]), __uint_as_float(r[k + 3])));
                    }
                }
                if (EXP) {
                    float pr[32];
                    constexpr bool PAIR = GMD_ATTN2_F32X2 && GMD_ATTN2_POLYPAIR && POLY > 0;
                    auto expo = [&](int k, float x) {
                        if (GMD_ATTN2_KO & 1) return x;
                        if (!PAIR && POLY > 0 && (k % (POLY > 0 ? POLY : 1)) == POLY - 1) { px = fmaxf(px, x); return exp2_poly(x); }
                        return ex2(x);
                    };
#if GMD_ATTN2_F32X2
#pragma unroll
                    for (int k = 0; k < 32; k += 2) {
                        const uint64_t x2 = f32x2_fma(f32x2_pack(__uint_as_float(r[k]), __uint_as_float(r[k + 1])), c2, nm2);
                        if (PAIR && !(GMD_ATTN2_KO & 1) && (k % (2 * (POLY > 0 ? POLY : 1))) == 2 * POLY - 2) {
                            exp2_poly2(x2, pr[k], pr[k + 1], px);
                        } else {
                            float x0, x1;
                            f32x2_unpack(x2, x0, x1);
                            pr[k] = expo(k, x0); pr[k + 1] = expo(k + 1, x1);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < ((GMD_ATTN2_KO & 8) ? 4 : 32); k += 4) { l01 = f32x2_add(l01, f32x2_pack(pr[k], pr[k + 1])); l23 = f32x2_add(l23, f32x2_pack(pr[k + 2], pr[k + 3])); }
#else
#pragma unroll
                    for (int k = 0; k < 32; ++k) pr[k] = expo(k, fmaf(__uint_as_float(r[k]), c, -m_use));
#pragma unroll
                    for (int k = 0; k < ((GMD_ATTN2_KO & 8) ? 4 : 32); k += 4) { l0 += pr[k]; l1 += pr[k + 1]; l2 += pr[k + 2]; l3 += pr[k + 3]; }
#endif
#pragma unroll
                    for (int k = 0; k < 16; ++k) pk[ch * 16 + k] = pack_bf16x2(pr[2 * k], pr[2 * k + 1]);
                }
            }
            mx_out = fmaxf(mx0, mx1);
#if GMD_ATTN2_F32X2
            f32x2_unpack(f32x2_add(l01, l23), l0, l1);
            l2 = l3 = 0.0f;
#endif
            lsum = (l0 + l1) + (l2 + l3);
            px_out = px;
        };
        auto do_tile = [&](auto mask_tag, auto first_tag, int j) {
            constexpr bool FIRST = decltype(first_tag)::value;   // j == 0 (its own copy of the tile body: no per-tile branch in the steady state)
            const int sb = j % SB, pb = j % PB;
            const uint32_t s_addr = tm_s(t, sb) + lane_off + hw * (BKV / HS), p_addr = tm_p(t, pb, hw) + lane_off;
            const int valid = args.Nk - j * BKV - hw * (BKV / HS);
            AWAIT(&s_full[t * SB + sb], (j / SB) & 1);
            tc_fence_after();
            if (warp == SW0 && lane == 0) { TRACE(6, j); if (j == 0) TRACE_CLK(0); }
            float mx, ls, px;
            uint32_t pk[NCH * 16];
            if (FIRST) {   // first tile: maximum first
                tile_pass(std::false_type{}, mask_tag, s_addr, valid, 0.0f, pk, mx, ls, px);
                m = mx * c;
            }
            const float m_use = m == -INFINITY ? 0.0f : m;   // no valid key yet: 2^(-inf) = 0
            tile_pass(std::true_type{}, mask_tag, s_addr, valid, m_use, pk, mx, ls, px);
#if GMD_ATTN2_SUMGROW
            // every probability is positive, so a row sum <= 2^16 proves that no exponent of this thread's keys exceeded 16 (the window is
            // wide because P is bf16 with fp32's exponent range and O, l are fp32: nothing overflows below 2^100); +inf (an exponent > 128) trips it too
            bool grow = ls > 65536.0f || px > 16.0f;
#else
            float mt = mx * c;
            const bool grow = mt > m + C::LAZY_T;
#endif
            if (__builtin_expect(__any_sync(0xffffffffu, grow), 0)) {
                // rare: a row maximum left the lazy window.  S_j is still intact in TMEM (P_j has not been stored yet and the buffer is
                // only released by the p_full arrival below): take the new maximum, bring O and l to the new scale and exponentiate
                // the tile again.
#if GMD_ATTN2_SUMGROW
                tile_pass(std::false_type{}, mask_tag, s_addr, valid, 0.0f, pk, mx, ls, px);   // the tile's true maximum, from S
                const float mt = mx * c;
                grow = grow && mt > m;
#endif
                const float m_new = grow ? mt : m;
                const float alpha = grow ? ex2(m - m_new) : 1.0f;    // (m is finite here: j > 0 or the maximum-first pass ran)
                m = m_new;
                l *= alpha;
                if (j > 0) {
                    AWAIT(&pv_done[t * PB + (j - 1) % PB], ((j - 1) / PB) & 1);   // O complete up to tile j-1 (P V_j waits for our p_full arrival)
                    tc_fence_after();
#pragma unroll
                    for (int oc = 0; oc < C::DPV / 16; ++oc) {
                        uint32_t o[16];
                        tmem_ld_32x16(my_o + oc * 16, o);
                        tmem_wait_ld();
#pragma unroll
                        for (int k = 0; k < 16; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
                        tmem_st_32x16(my_o + oc * 16, o);
                    }
                    tmem_wait_st();
                }
                tile_pass(std::true_type{}, mask_tag, s_addr, valid, m, pk, mx, ls, px);
            }
            l += ls;
            // The P buffer was last read by P V_{j-PB}.  ALIAS: P V_{j-1} was issued before S_j by the same thread, so s_full(j) implies
            // it is complete; PB == SB likewise (P V_{j-PB} precedes S_j in the issue order).  Otherwise wait for it.
            if (!ALIAS && PB != SB && j >= PB) { AWAIT(&pv_done[t * PB + pb], ((j - PB) / PB) & 1); tc_fence_after(); }
            if (!(GMD_ATTN2_KO & 2)) {
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) tmem_st_32x16p(p_addr + ch * 16, pk + ch * 16);
                tmem_wait_st();
            } else if (pk[0] == 0x12345678u && pk[17] == 0x9abcdef0u) {
                tmem_st_32x16p(p_addr, pk);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[t * PB + pb]);
            if (warp == SW0 && lane == 0) { TRACE(7, j); if (j == T - 1) TRACE_CLK(1); }
        };
        const bool ragged = (args.Nk % BKV) != 0;
        if (T == 1) {
            if (ragged) do_tile(std::true_type{}, std::true_type{}, 0); else do_tile(std::false_type{}, std::true_type{}, 0);
        } else {
            do_tile(std::false_type{}, std::true_type{}, 0);
            constexpr int UNROLL_J = GMD_ATTN2_UNROLL;
#pragma unroll UNROLL_J
            for (int j = 1; j < T - 1; ++j) do_tile(std::false_type{}, std::false_type{}, j);
            if (ragged) do_tile(std::true_type{}, std::false_type{}, T - 1); else do_tile(std::false_type{}, std::false_type{}, T - 1);
        }
        AWAIT(&pv_done[t * PB + (T - 1) % PB], ((T - 1) / PB) & 1);
        tc_fence_after();
        const int q = q0 + t * BQ + row;
        __nv_bfloat16* op = args.o + batch * args.o_stride_b + (int64_t)q * args.o_stride_n + head * args.o_stride_h;
        // merge of the row's HS chains: common maximum, weights 2^(m_h - m), one denominator
        float wgt[HS];
        float inv_l;
        if (HS == 1) {
            wgt[0] = 1.0f;
            inv_l = 1.0f / l;
        } else {
            float2* mg = reinterpret_cast<float2*>(smem + C::OFF_MERGE) + t * HS * BQ;
            mg[hw * BQ + row] = make_float2(m, l);
            asm volatile("bar.sync %0, %1;" ::"r"(1 + t), "r"(128 * HS) : "memory");
            float mf = -INFINITY;
#pragma unroll
            for (int h = 0; h < HS; ++h) mf = fmaxf(mf, mg[h * BQ + row].x);
            float lf = 0.0f;
#pragma unroll
            for (int h = 0; h < HS; ++h) { const float2 ml = mg[h * BQ + row]; wgt[h] = ex2(ml.x - mf); lf = fmaf(ml.y, wgt[h], lf); }
            inv_l = 1.0f / lf;
        }
#pragma unroll
        for (int oc = 0; oc < C::DPV / 16; ++oc) {
            if (oc % HS != hw) continue;     // the row's threads share the output columns
            float acc[16];
#pragma unroll
            for (int h = 0; h < HS; ++h) {
                uint32_t o[16];
                tmem_ld_32x16(tm_o(t, h) + lane_off + oc * 16, o);
                tmem_wait_ld();
#pragma unroll
                for (int k = 0; k < 16; ++k) acc[k] = h == 0 ? __uint_as_float(o[k]) * wgt[0] : fmaf(__uint_as_float(o[k]), wgt[h], acc[k]);
            }
            if (q < args.Nq) {
#pragma unroll
                for (int h8 = 0; h8 < 2; ++h8) {
                    const int d0 = oc * 16 + h8 * 8;
                    if (d0 < D) {  // D is a multiple of 8
                        uint4 v = make_uint4(pack_bf16x2(acc[h8 * 8 + 0] * inv_l, acc[h8 * 8 + 1] * inv_l), pack_bf16x2(acc[h8 * 8 + 2] * inv_l, acc[h8 * 8 + 3] * inv_l),
                                             pack_bf16x2(acc[h8 * 8 + 4] * inv_l, acc[h8 * 8 + 5] * inv_l), pack_bf16x2(acc[h8 * 8 + 6] * inv_l, acc[h8 * 8 + 7] * inv_l));
                        *reinterpret_cast<uint4*>(op + d0) = v;
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<C::TMEM_COLS>(tmem_base);
    }
}

int g_attn_v1 = -1;     // GMD_ATTN_V1=1 in the environment: keep the round-1 kernel for d = 40 / 80 (A/B measurements)
int g_attn2_cfg40 = 0;  // GMD_ATTN2_CFG40: alternative d = 40 configurations for A/B measurements, see gmd_attn_fwd

template <int D, int BKV, int NQT, bool ALIAS, int SB, int PB, int KS, int HS = 1, int WG = 0, int POLY = GMD_ATTN2_POLY>
int launch2(const gmd_attn_params* p, cudaStream_t st) {
    using C = Cfg2<D, BKV, NQT, ALIAS, SB, PB, KS, HS, WG, POLY>;
    static bool configured[kMaxDevices] = {};
    const int dev = device_ordinal();
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(attn2_kernel<D, BKV, NQT, ALIAS, SB, PB, KS, HS, WG, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) { set_last_error("attn2: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return kErrCuda; }
        configured[dev] = true;
    }
    CUtensorMap mq, mk, mv;
    int rc;
    {
        uint64_t dims[4] = {(uint64_t)D, (uint64_t)p->H, (uint64_t)p->Nq, (uint64_t)p->B};
        uint64_t strides[4] = {2, (uint64_t)p->q_stride_h * 2, (uint64_t)p->q_stride_n * 2, (uint64_t)p->q_stride_b * 2};
        uint32_t box[4] = {64, 1, (uint32_t)BQ, 1};
        if ((rc = encode_tensor_map_bf16(&mq, p->q, 4, dims, strides, box, true))) return rc;
    }
    auto enc_dense = [&](CUtensorMap* m, const void* base, int64_t sb, int64_t sn, int n) {
        uint64_t dims[3] = {(uint64_t)D * p->H, (uint64_t)n, (uint64_t)p->B};
        uint64_t strides[3] = {2, (uint64_t)sn * 2, (uint64_t)sb * 2};
        uint32_t box[3] = {64, (uint32_t)BKV, 1};
        return encode_tensor_map_bf16(m, base, 3, dims, strides, box, true);
    };
    if ((rc = enc_dense(&mk, p->k, p->k_stride_b, p->k_stride_n, p->Nk))) return rc;
    if ((rc = enc_dense(&mv, p->v, p->v_stride_b, p->v_stride_n, p->Nk))) return rc;
    AttnArgs a;
    a.o = static_cast<__nv_bfloat16*>(p->o);
    a.o_stride_b = p->o_stride_b; a.o_stride_n = p->o_stride_n; a.o_stride_h = p->o_stride_h;
    a.Nq = p->Nq; a.Nk = p->Nk;
    a.scale_log2 = p->scale * 1.4426950408889634f;
    a.kv_dense = 1;
    dim3 grid((p->Nq + BQ * NQT - 1) / (BQ * NQT), p->H, p->B);
    attn2_kernel<D, BKV, NQT, ALIAS, SB, PB, KS, HS, WG, POLY><<<grid, C::THREADS, C::SMEM, st>>>(mq, mk, mv, a);
    count_launch(1);
    return check_launch("attn2_kernel");
}


// ---------------------------------------------------------------------------------------------------------------------------------
// Text cross-attention (64 < Nk <= 80: SD1.5's 77 CLIP tokens), d = 40 / 80.  K and V of one (batch, head) stay resident in shared
// memory and the CTA walks over QT consecutive 128-row query tiles; all keys are visible at once, so the softmax is exact in one pass
// (no running maximum, no rescale).  Two softmax warp sets own alternating query tiles together with their Q slot, S / O
// accumulators and P buffer, so the S MMA, the softmax, the P V MMA and the output write of neighbouring tiles overlap.  The
// per-tile CTA of the general kernel lived for two key tiles and was bound by its prologue (TMEM allocation, Q/K/V round trip).
template <int D>
struct XCfg {
    static constexpr int NDB = (D + 63) / 64;
    static constexpr int DP = (D + 15) / 16 * 16;
    static constexpr int DPV = (D + 1 + 15) / 16 * 16;
    static constexpr int NKP = 80;                        // keys padded to a multiple of 16 (N of Q K^T, K of P V)
    static constexpr int Q_BYTES = NDB * BQ * 128;
    static constexpr int KV_BLOCK = NKP * 128;            // one 64-wide d block of K or V
    static constexpr int KV_BYTES = NDB * KV_BLOCK;
    static constexpr int P_BLOCK = BQ * 128;              // P block 0: keys 0-63, block 1: keys 64-79 (first 32 bytes of each row)
    static constexpr int P_BYTES = 2 * P_BLOCK;
    // softmax warp sets = query tiles in flight.  The per-tile chain (S ready -> 80-column TMEM read -> exponentials -> P -> P V ->
    // O read -> global store) is ~5000 cycles of latency, so d = 40 runs four sets (all 512 TMEM columns: 4 x (80 + 48), 212 KB of
    // shared memory); d = 80 has room for two.
    static constexpr int NSET = D == 40 ? 4 : 2;
    static constexpr int SSTRIDE = NSET == 4 ? NKP : 128;  // TMEM columns between the S accumulators of consecutive sets
    static constexpr int OFF_K = NSET * Q_BYTES;
    static constexpr int OFF_V = OFF_K + KV_BYTES;
    static constexpr int OFF_P = OFF_V + KV_BYTES;
    static constexpr int OFF_BAR = OFF_P + NSET * P_BYTES;
    static constexpr int SMEM = OFF_BAR + 512 + 1024;
    static constexpr int THREADS = 64 + 128 * NSET;
    static_assert(KV_BLOCK % 1024 == 0, "swizzle atoms");
    static_assert(NSET * (SSTRIDE + DPV) <= 512, "TMEM");
    static_assert(SMEM <= 227 * 1024, "shared memory");
};

template <int D>
__global__ void __launch_bounds__(XCfg<D>::THREADS, 1)
xattn_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
             const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_o, const AttnArgs args, const int QT) {
    using C = XCfg<D>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* q_smem = smem;
    uint8_t* k_smem = smem + C::OFF_K;
    uint8_t* v_smem = smem + C::OFF_V;
    uint8_t* p_smem = smem + C::OFF_P;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    constexpr int NS = C::NSET;
    uint64_t* kv_full = bars;           // 1
    uint64_t* q_full = bars + 1;        // NS each from here on
    uint64_t* q_empty = q_full + NS;
    uint64_t* s_full = q_empty + NS;
    uint64_t* s_free = s_full + NS;
    uint64_t* p_full = s_free + NS;
    uint64_t* o_full = p_full + NS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + NS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int head = blockIdx.y, batch = blockIdx.z;
    const int nq_tiles = (args.Nq + BQ - 1) / BQ;
    const int i0 = blockIdx.x * QT;
    const int n = min(QT, nq_tiles - i0);   // query tiles of this CTA

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v); tma_prefetch_desc(&map_o);
        mbar_init(kv_full, 1);
        for (int s = 0; s < NS; ++s) {
            mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); mbar_init(&s_full[s], 1);
            mbar_init(&s_free[s], 128); mbar_init(&p_full[s], 128); mbar_init(&o_full[s], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_o = tmem_base + NS * C::SSTRIDE;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(kv_full, 2 * C::KV_BYTES);
            for (int b = 0; b < C::NDB; ++b) {
                tma_load_4d(k_smem + b * C::KV_BLOCK, &map_k, kv_full, b * 64, head, 0, batch);   // zero-filled past d (once per CTA)
                tma_load_3d(v_smem + b * C::KV_BLOCK, &map_v, kv_full, head * D + b * 64, 0, batch);
            }
            for (int i = 0; i < n; ++i) {
                const int slot = i % NS;
                mbar_wait(&q_empty[slot], ((i / NS) & 1) ^ 1);
                mbar_expect_tx(&q_full[slot], C::Q_BYTES);
                for (int b = 0; b < C::NDB; ++b)
                    tma_load_3d(q_smem + slot * C::Q_BYTES + b * BQ * 128, &map_q, &q_full[slot], head * D + b * 64, (i0 + i) * BQ, batch);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t IDESC_S = umma_idesc_bf16(BQ, C::NKP, false, false);
            constexpr uint32_t IDESC_O = umma_idesc_bf16(BQ, C::DPV, false, true);  // B = V is MN-major
            const uint32_t q_addr = smem_u32(q_smem), k_addr = smem_u32(k_smem), v_addr = smem_u32(v_smem), p_addr = smem_u32(p_smem);
            auto issue_s = [&](int i) {
                const int slot = i % NS;
                mbar_wait(&q_full[slot], (i / NS) & 1);
                if (i >= NS) mbar_wait(&s_free[slot], ((i / NS) - 1) & 1);   // the set has its previous S in registers
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < C::DP / 16; ++ks) {
                    const int blk = ks >> 2, within = ks & 3;
                    uint64_t da = umma_desc_k_sw128(q_addr + slot * C::Q_BYTES + blk * (BQ * 128) + within * 32);
                    uint64_t db = umma_desc_k_sw128(k_addr + blk * C::KV_BLOCK + within * 32);
                    umma_bf16_ss(tmem_base + slot * C::SSTRIDE, da, db, IDESC_S, ks != 0 ? 1u : 0u);
                }
                umma_commit(&q_empty[slot]);
                umma_commit(&s_full[slot]);
            };
            mbar_wait(kv_full, 0);
            for (int i = 0; i < NS && i < n; ++i) issue_s(i);
            for (int i = 0; i < n; ++i) {
                const int slot = i % NS;
                mbar_wait(&p_full[slot], (i / NS) & 1);   // P_i in smem, ones column set, O[slot] drained by the previous tile's epilogue
                tc_fence_after();
#pragma unroll
                for (int ks = 0; ks < C::NKP / 16; ++ks) {
                    uint64_t da = umma_desc_k_sw128(p_addr + slot * C::P_BYTES + (ks >> 2) * C::P_BLOCK + (ks & 3) * 32);
                    uint64_t db = umma_desc_mn_sw128(v_addr + ks * 16 * 128, C::KV_BLOCK);
                    umma_bf16_ss(tmem_o + slot * C::DPV, da, db, IDESC_O, ks != 0 ? 1u : 0u);
                }
                umma_commit(&o_full[slot]);
                if (i + NS < n) issue_s(i + NS);
            }
        }
    } else {
        const int lg = warp & 3;
        const int set = (warp - 2) >> 2;
        const int row = lg * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(lg * 32) << 16;
        const uint32_t my_s = tmem_base + set * C::SSTRIDE + lane_off;
        const uint32_t my_o = tmem_o + set * C::DPV + lane_off;
        const float c = args.scale_log2;
        uint8_t* prow = p_smem + set * C::P_BYTES + row * 128;
        const bool issuer = ((warp - 2) & 3) == 0 && lane == 0;   // one thread per set owns its TMA-store bulk groups
        // ones column of V (column D of its padding) -> the P V MMA also accumulates the softmax denominator.  Both sets write it
        // (idempotent) so that each set's first p_full arrival orders it before that set's first P V.
        mbar_wait(kv_full, 0);
        if (row < C::NKP) {
            constexpr int blk = D / 64, cc = (D % 64) / 8, within = (D % 8) * 2;
            uint8_t* vrow = v_smem + blk * C::KV_BLOCK + row * 128;
            *reinterpret_cast<uint16_t*>(vrow + ((cc ^ (row & 7)) << 4) + within) = 0x3F80;  // bf16 1.0
        }
        for (int i = set; i < n; i += NS) {
            const uint32_t ph = (i / NS) & 1;
            mbar_wait(&s_full[set], ph);
            tc_fence_after();
            uint32_t sr[C::NKP];
            {
                uint32_t t0[32], t1[32], t2[16];
                tmem_ld_32x32(my_s, t0);
                tmem_ld_32x32(my_s + 32, t1);
                tmem_ld_32x16(my_s + 64, t2);
                tmem_wait_ld();
#pragma unroll
                for (int k = 0; k < 32; ++k) { sr[k] = t0[k]; sr[32 + k] = t1[k]; }
#pragma unroll
                for (int k = 0; k < 16; ++k) sr[64 + k] = t2[k];
            }
            tc_fence_before();
            mbar_arrive(&s_free[set]);   // S[set] may be overwritten by S_{i+2}
#pragma unroll
            for (int k = 64; k < C::NKP; ++k) if (k >= args.Nk) sr[k] = 0xff800000u;  // padding keys -> -inf
            float mx0 = __uint_as_float(sr[0]), mx1 = __uint_as_float(sr[1]), mx2 = __uint_as_float(sr[2]), mx3 = __uint_as_float(sr[3]);
#pragma unroll
            for (int k = 4; k < ((GMD_XATTN_KO & 8) ? 4 : C::NKP); k += 4) {
                mx0 = fmaxf(mx0, __uint_as_float(sr[k])); mx1 = fmaxf(mx1, __uint_as_float(sr[k + 1]));
                mx2 = fmaxf(mx2, __uint_as_float(sr[k + 2])); mx3 = fmaxf(mx3, __uint_as_float(sr[k + 3]));
            }
            const float m = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * c;
            // P[set] was last read by P V of this set's previous tile (awaited before its epilogue) and then by the TMA store of that
            // tile's output, which used it as staging: the issuing thread waits for the store to have read it
            if (i >= NS) {
                if (issuer) tma_store_wait_read();
                asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");
            }
#pragma unroll
            for (int ch = 0; ch < C::NKP / 8; ++ch) {
                uint32_t pk[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float x0 = fmaf(__uint_as_float(sr[ch * 8 + 2 * k]), c, -m), x1 = fmaf(__uint_as_float(sr[ch * 8 + 2 * k + 1]), c, -m);
                    pk[k] = (GMD_XATTN_KO & 1) ? pack_bf16x2(x0, x1) : pack_bf16x2(ex2(x0), ex2(x1));
                }
                if ((GMD_XATTN_KO & 4) && pk[0] != 0x12345678u) continue;
                // K-major SWIZZLE_128B: 16-byte chunk cc of row r lands at chunk (cc ^ (r & 7)) of its 128-byte row
                uint8_t* dst = prow + (ch >> 3) * C::P_BLOCK + (((ch & 7) ^ (row & 7)) << 4);
                *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(&p_full[set]);
            // epilogue of this tile: O / l -> global
            mbar_wait(&o_full[set], ph);
            tc_fence_after();
            // O / l as bf16 into this set's P buffer (its P V is complete) in the SWIZZLE_128B layout, then ONE asynchronous TMA store
            // per 64-wide d block: rows past Nq and columns past d are clipped by the tensor map.  (Per-thread 16-byte global stores
            // of 80-byte rows cost 15 of 42 us at B=16, Nq=4096.)
            float inv_l;
            {
                uint32_t o[16];
                tmem_ld_32x16(my_o + (D / 16) * 16, o);
                tmem_wait_ld();
                inv_l = 1.0f / __uint_as_float(o[D % 16]);
            }
            uint8_t* orow = p_smem + set * C::P_BYTES + row * 128;
#pragma unroll
            for (int ch = 0; ch < (D + 15) / 16; ++ch) {
                uint32_t o[16];
                tmem_ld_32x16(my_o + ch * 16, o);
                tmem_wait_ld();
#pragma unroll
                for (int h8 = 0; h8 < 2; ++h8) {
                    const int d0 = ch * 16 + h8 * 8;
                    if (d0 < D) {  // D is a multiple of 8
                        uint4 v = make_uint4(pack_bf16x2(__uint_as_float(o[h8 * 8 + 0]) * inv_l, __uint_as_float(o[h8 * 8 + 1]) * inv_l),
                                             pack_bf16x2(__uint_as_float(o[h8 * 8 + 2]) * inv_l, __uint_as_float(o[h8 * 8 + 3]) * inv_l),
                                             pack_bf16x2(__uint_as_float(o[h8 * 8 + 4]) * inv_l, __uint_as_float(o[h8 * 8 + 5]) * inv_l),
                                             pack_bf16x2(__uint_as_float(o[h8 * 8 + 6]) * inv_l, __uint_as_float(o[h8 * 8 + 7]) * inv_l));
                        *reinterpret_cast<uint4*>(orow + (d0 / 64) * C::P_BLOCK + ((((d0 % 64) / 8) ^ (row & 7)) << 4)) = v;
                    }
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");
            if (issuer) {
                for (int b = 0; b < C::NDB; ++b)
                    tma_store_4d(&map_o, p_smem + set * C::P_BYTES + b * C::P_BLOCK, b * 64, head, (i0 + i) * BQ, batch);
                tma_store_commit();
            }
        }
        if (issuer) tma_store_wait_all();
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

template <int D>
int launch_x(const gmd_attn_params* p, cudaStream_t st) {
    using C = XCfg<D>;
    static bool configured[kMaxDevices] = {};
    static int sms_of[kMaxDevices] = {};
    const int dev = device_ordinal();
    if (!configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(xattn_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) { set_last_error("xattn: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return kErrCuda; }
        cudaDeviceGetAttribute(&sms_of[dev], cudaDevAttrMultiProcessorCount, dev);
        if (sms_of[dev] <= 0) sms_of[dev] = 148;
        configured[dev] = true;
    }
    const int sms = sms_of[dev];
    CUtensorMap mq, mk, mv;
    int rc;
    // Q (one tile per step of the loop) and V go through dense 3-D maps over the whole channel row: the columns past the head's d
    // hold the next head's values.  K — loaded once per CTA — keeps the 4-D (d, head, token, batch) map whose out-of-bounds fill
    // zeroes its columns d..63, so the padded part of the Q K^T contraction (d = 40: columns 40-47) contributes exactly 0.
    auto enc_dense = [&](CUtensorMap* m, const void* base, int64_t sb, int64_t sn, int n, int rows) {
        uint64_t dims[3] = {(uint64_t)D * p->H, (uint64_t)n, (uint64_t)p->B};
        uint64_t strides[3] = {2, (uint64_t)sn * 2, (uint64_t)sb * 2};
        uint32_t box[3] = {64, (uint32_t)rows, 1};
        return encode_tensor_map_bf16(m, base, 3, dims, strides, box, true);
    };
    if ((rc = enc_dense(&mq, p->q, p->q_stride_b, p->q_stride_n, p->Nq, BQ))) return rc;
    {
        uint64_t dims[4] = {(uint64_t)D, (uint64_t)p->H, (uint64_t)p->Nk, (uint64_t)p->B};
        uint64_t strides[4] = {2, (uint64_t)p->k_stride_h * 2, (uint64_t)p->k_stride_n * 2, (uint64_t)p->k_stride_b * 2};
        uint32_t box[4] = {64, 1, (uint32_t)C::NKP, 1};
        if ((rc = encode_tensor_map_bf16(&mk, p->k, 4, dims, strides, box, true))) return rc;
    }
    if ((rc = enc_dense(&mv, p->v, p->v_stride_b, p->v_stride_n, p->Nk, C::NKP))) return rc;
    CUtensorMap mo;
    {
        uint64_t dims[4] = {(uint64_t)D, (uint64_t)p->H, (uint64_t)p->Nq, (uint64_t)p->B};
        uint64_t strides[4] = {2, (uint64_t)p->o_stride_h * 2, (uint64_t)p->o_stride_n * 2, (uint64_t)p->o_stride_b * 2};
        uint32_t box[4] = {64, 1, (uint32_t)BQ, 1};
        if ((rc = encode_tensor_map_bf16(&mo, p->o, 4, dims, strides, box, true))) return rc;
    }
    AttnArgs a;
    a.o = static_cast<__nv_bfloat16*>(p->o);
    a.o_stride_b = p->o_stride_b; a.o_stride_n = p->o_stride_n; a.o_stride_h = p->o_stride_h;
    a.Nq = p->Nq; a.Nk = p->Nk;
    a.scale_log2 = p->scale * 1.4426950408889634f;
    a.kv_dense = 1;
    // query tiles per CTA: fewest waves x (prologue + QT tiles), in units of ~950 cycles per tile and ~2500 per prologue
    const int nq_tiles = (p->Nq + BQ - 1) / BQ;
    int best_qt = nq_tiles;
    double best = 1e30;
    for (int parts = 1; parts <= nq_tiles; parts *= 2) {
        const int qt = (nq_tiles + parts - 1) / parts;
        const int64_t ctas = (int64_t)p->B * p->H * ((nq_tiles + qt - 1) / qt);
        const double cost = (double)((ctas + sms - 1) / sms) * (2500.0 + 950.0 * qt);
        if (cost < best) { best = cost; best_qt = qt; }
    }
    dim3 grid((nq_tiles + best_qt - 1) / best_qt, p->H, p->B);
    xattn_kernel<D><<<grid, C::THREADS, C::SMEM, st>>>(mq, mk, mv, mo, a, best_qt);
    count_launch(1);
    return check_launch("xattn_kernel");
}

}  // namespace
}  // namespace gmd

#ifdef GMD_ATTN2_TRACE
extern "C" int gmd_attn_trace_dump(long long* host, int n) {
    return (int)cudaMemcpyFromSymbol(host, gmd::g_attn_trace, sizeof(long long) * n);
}
#endif

extern "C" int gmd_attn_fwd(const gmd_attn_params* p, void* stream) {
    using namespace gmd;
    if (!p || !p->q || !p->k || !p->v || !p->o) { set_last_error("gmd_attn_fwd: null pointer"); return kErrInvalid; }
    if (p->B <= 0 || p->H <= 0 || p->Nq <= 0 || p->Nk <= 0) { set_last_error("gmd_attn_fwd: empty problem"); return kErrInvalid; }
    if (!(p->scale > 0.0f)) { set_last_error("gmd_attn_fwd: scale must be positive"); return kErrInvalid; }
    const int64_t strides[] = {p->q_stride_b, p->q_stride_n, p->q_stride_h, p->k_stride_b, p->k_stride_n, p->k_stride_h,
                               p->v_stride_b, p->v_stride_n, p->v_stride_h, p->o_stride_b, p->o_stride_n, p->o_stride_h};
    for (int64_t s : strides)
        if (s % 8) { set_last_error("gmd_attn_fwd: strides must be multiples of 8 elements (16 bytes)"); return kErrInvalid; }
    const void* ptrs[] = {p->q, p->k, p->v, p->o};
    for (const void* q : ptrs)
        if (reinterpret_cast<uintptr_t>(q) & 15) { set_last_error("gmd_attn_fwd: pointers must be 16-byte aligned"); return kErrInvalid; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // text cross-attention (SD1.5: 77 keys): K / V resident, persistent over query tiles
    const bool xattn = p->Nk > 64 && p->Nk <= 80 && p->q_stride_h == p->d && p->v_stride_h == p->d;
    if (g_attn_v1 < 0) {
        const char* e = getenv("GMD_ATTN_V1");
        g_attn_v1 = (e && e[0] == '1') ? 1 : 0;
        e = getenv("GMD_ATTN2_CFG40");
        if (e) g_attn2_cfg40 = atoi(e);
    }
    // self-attention (key tiles beyond the short configurations, K / V rows dense over the heads): P-in-TMEM kernel
    const bool v2 = !g_attn_v1 && !xattn && p->Nk > 2 * BKV && p->k_stride_h == p->d && p->v_stride_h == p->d;
    switch (p->d) {
        case 40:
            if (xattn) return launch_x<40>(p, st);
            // default (same-box A/B, B=16 N=4096: 688 vs 704 us): two threads per query row on independent key halves (P over S, two S
            // buffers, four light softmax warps per sub-partition) with every 8th exponential as an FMA-pipe polynomial.  Kept for A/B
            // runs: 1 = one thread per row, S and P double-buffered (the round's first default), 2 = three aliased S
            // buffers, 3 = 128-key tiles, two chains per CTA, one CTA per SM with setmaxnreg-rebalanced warpgroups (917 us)
            if (v2) return g_attn2_cfg40 == 1 ? launch2<40, 64, 1, false, 2, 2, 3>(p, st) : g_attn2_cfg40 == 2 ? launch2<40, 64, 1, true, 3, 3, 4, 1>(p, st)
                         : g_attn2_cfg40 == 3 ? launch2<40, 128, 2, true, 1, 1, 3, 1, 1>(p, st) : launch2<40, 64, 1, true, 2, 2, 3, 2, 0, GMD_ATTN2_POLY40>(p, st);
            return p->Nk <= 2 * BKV ? launch<40, true>(p, st) : launch<40, false>(p, st);
        case 80:
            if (xattn) return launch_x<80>(p, st);
            if (v2) return GMD_ATTN2_ALIAS80 ? launch2<80, 64, 1, true, 2, 2, 2, 1, 0, GMD_ATTN2_POLY80>(p, st) : launch2<80, 64, 1, false, 2, 1, 2, 1, 0, GMD_ATTN2_POLY80>(p, st);
            return launch<80, false>(p, st);   // (the short configuration does not add a third resident CTA at d = 80: measured slightly slower)
        case 160: return launch<160, false>(p, st);
        default: set_last_error("gmd_attn_fwd: head dim %d not instantiated (40, 80, 160)", p->d); return kErrUnsupported;
    }
}
