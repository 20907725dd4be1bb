// window_tc.cu -- windowed attention of the SAM encoder (14 x 14 windows, head dim 80) on the tcgen05 tensor cores,
// one kernel, no pre-passes.
//
//   out = softmax(scale * q k^T + rel_h[q, kh(k)] + rel_w[q, kw(k)]) v     (image_encoder.py:235-260, 354-392;
//   zero-pad tokens of window_partition :263-288 are ordinary keys / queries, as in the reference)
//
// Work item = one (window, head): 196 queries as two groups of 128 TMEM lanes (rows past 196 belong to the next
// window: computed, never stored), 196 keys padded to 208. Persistent CTAs (one per SM) loop over the items.
//
// Everything a score needs is produced by the tensor core:
//   G   = Q Rcat^T            (128 x 64 x 80)   Rcat = [rel_h table rows | rel_w table rows]; row j <-> offset q - k + 13
//   S   = [Q | Gsel] [K | Sel]^T  (128 x 208 x 112)   Gsel[q] = [G[q][qh+13-kh], kh = 0..13 | G[q][32+qw+13-kw]] / scale (fp16,
//         written by the softmax threads into the second A box), Sel = constant 0/1 selection matrix built in shared memory
//   O   = P V                 (128 x 80 x 208)  P from tensor memory (TS-mode UMMA), V read in place as an MN-MAJOR B operand:
//         the TMA box [keys][64 head-dim columns] of the qkv rows is exactly the canonical 128B-swizzled MN-major atom
//         (8 keys x 64 columns), so no V^T copy exists anywhere
//   l   = P 1                 (128 x 16 x 208)  row sums from the same fp16-rounded weights (constant ones operand)
// The softmax is column-split: TWO threads per score row (keys 0..95 / 96..207), i.e. 16 softmax warps per CTA, which
// halves every serial per-row chain (the one-thread-per-row designs were latency-bound at ~2 warps per scheduler).
// Each query group has its own MMA-issuer warp, so the two groups are independent pipelines that share the tensor pipe.
//
// Tensor-memory columns of a group (base 256 * group):  G 0..63 (dead once Gsel is written)  ->  S 0..207  ->
//   P (packed fp16 pairs, in place) 0..47 (keys 0..95) and 96..151 (keys 96..207)  ->  O 160..239, l 240..255.
#include <cudaTypedefs.h>

#include "ops.h"

namespace {

constexpr int W_S = 196, W_G = 14, W_HD = 80, W_KP = 208, W_BQ = 128;
constexpr float W_L2E = 1.4426950408889634f;
constexpr uint32_t W_QB = W_BQ * 128;      // one 128-row x 64-column box
constexpr uint32_t W_KB = W_KP * 128;      // one 208-row x 64-column box
constexpr uint32_t W_RB = 64 * 128;        // rel-pos table box (64 rows)
constexpr uint32_t W_ONES = 4 * 16 * 128;  // ones operand: 4 key blocks x 16 rows x 128 B
constexpr int W_THREADS = 640;             // warps: 0 TMA, 1 issuer group 0, 2 TMEM alloc + issuer group 1, 3 idle, 4..19 softmax
constexpr uint32_t W_SMEM = 1024 + 4 * W_QB + 5 * W_KB + 2 * W_RB + W_ONES + 2 * 2 * 128 * 4 + 384;
constexpr uint32_t W_COL_PB = 96, W_COL_O = 160, W_COL_L = 240;
constexpr int W_KA = 96;                   // keys of the first half (6 k-steps); the second half has 112 (7 k-steps, 100 valid)

__device__ __forceinline__ uint32_t sel_b32(bool c, uint32_t a, uint32_t b) { return c ? a : b; }

// in[i] (i = 0..27, fp16 pairs in p[0..13], p[14] = 0) -> out[k] = in[s + k], k = 0..13 (7 packed registers), 0 <= s <= 13
__device__ __forceinline__ void barrel14(const uint32_t (&p)[15], int s, uint32_t (&out)[7]) {
    uint32_t t1[11], t2[9], t3[8];
    const bool b8 = (s & 8) != 0, b4 = (s & 4) != 0, b2 = (s & 2) != 0;
#pragma unroll
    for (int i = 0; i < 11; ++i) t1[i] = sel_b32(b8, p[i + 4 < 15 ? i + 4 : 14], p[i]);
#pragma unroll
    for (int i = 0; i < 9; ++i) t2[i] = sel_b32(b4, t1[i + 2], t1[i]);
#pragma unroll
    for (int i = 0; i < 8; ++i) t3[i] = sel_b32(b2, i + 1 < 9 ? t2[i + 1] : 0u, t2[i]);
    const uint32_t amt = (uint32_t)(s & 1) * 16u;
#pragma unroll
    for (int i = 0; i < 7; ++i) out[i] = __funnelshift_r(t3[i], t3[i + 1], amt);
}

__global__ void __launch_bounds__(W_THREADS, 1)
window_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmK16, const __grid_constant__ CUtensorMap tmR, int heads, int n_items,
                 float scale, __half* __restrict__ out, int* __restrict__ sched_counter, int skew_clocks, long long* __restrict__ trace,
                 int un_g, int un_h, int un_w) {
    extern __shared__ uint8_t w_smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(w_smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = w_smem_raw + (smem_base - ptx::smem_u32(w_smem_raw));
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int D = heads * W_HD;
    const int n_work = n_items * heads;
    // debug timeline (tools/trace_window.py): CTA 0 stamps clock64() per event for its first 8 items
#define W_TR(ev) do { if (trace != nullptr && blockIdx.x == 0 && it < 8 && lane == 0) trace[it * 64 + (ev)] = clock64(); } while (0)

    const uint32_t sQ0 = smem_base;              // [group] Q columns 0..63
    const uint32_t sQG = sQ0 + 2 * W_QB;          // [group] Q columns 64..79 (TMA) | Gsel_h 14 + 2 | Gsel_w 14 + 2 (threads) | unused
    const uint32_t sK0 = sQG + 2 * W_QB;          // K columns 0..63, 208 rows
    const uint32_t sKt = sK0 + W_KB;              // K columns 16..79 (the fifth k-step reads its columns 48..63)
    const uint32_t sSel = sKt + W_KB;             // constant selection matrix
    const uint32_t sV0 = sSel + W_KB;             // V columns 0..63, 208 keys (MN-major B operand)
    const uint32_t sV1 = sV0 + W_KB;              // V columns 64..127 (64..79 used)
    const uint32_t sR0 = sV1 + W_KB;              // Rcat columns 0..63
    const uint32_t sRt = sR0 + W_RB;              // Rcat columns 16..79
    const uint32_t sOnes = sRt + W_RB;
    const uint32_t sX = sOnes + W_ONES;           // row-maximum exchange: float [group][half][128]
    const uint32_t bar = sX + 2 * 2 * 128 * 4;
    const uint32_t const_full = bar, qk_full = bar + 8, qk_free = bar + 16, v_full = bar + 24, v_free = bar + 32;
    auto g_full = [&](int g) { return bar + 8u * (5 + g); };
    auto qg_ready = [&](int g) { return bar + 8u * (7 + g); };
    auto s_full = [&](int g) { return bar + 8u * (9 + g); };
    auto p_full = [&](int g) { return bar + 8u * (11 + g); };
    auto o_full = [&](int g) { return bar + 8u * (13 + g); };
    const uint32_t tmem_slot = bar + 8u * 17;
    // item sequence: claimed dynamically by warp 3 (SchedRing, common.cuh) when a counter is given, else blockIdx.x, + gridDim.x, ...
    ptx::SchedRing sched;
    sched.carve(bar + 8u * 18);
    const bool dyn = sched_counter != nullptr;
    auto next_item = [&](int k) -> int {
        if (dyn) return sched.consume(k, 0, warp == 0);
        const int t = (int)blockIdx.x + k * (int)gridDim.x;
        return t < n_work ? t : -1;
    };

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmK16); ptx::prefetch_tmap(&tmR);
    }
    if (warp == 1 && lane == 0) {
        ptx::mbar_init(const_full, 1); ptx::mbar_init(qk_full, 1); ptx::mbar_init(qk_free, 2);
        ptx::mbar_init(v_full, 1); ptx::mbar_init(v_free, 2);
        for (int g = 0; g < 2; ++g) {
            ptx::mbar_init(g_full(g), 1); ptx::mbar_init(qg_ready(g), 256); ptx::mbar_init(s_full(g), 1);
            ptx::mbar_init(p_full(g), 256); ptx::mbar_init(o_full(g), 1);
        }
        sched.init(19, false);   // consumers: TMA warp, two issuer warps, sixteen softmax warps
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    }
    {
        // constant operands, written once per CTA (generic proxy, 128B swizzle: 16-byte chunk j of row r sits at j ^ (r & 7))
        // Sel [208 keys][64]: ones at columns kh(k) and 16 + kw(k)
        for (int i = threadIdx.x; i < W_KP * 8; i += W_THREADS) {
            const int k = i >> 3, j = i & 7;
            const int kh = k / W_G, kw = k - kh * W_G;
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c0 = j * 8 + 2 * e, c1 = c0 + 1;
                const uint32_t lo = (k < W_S && (c0 == kh || c0 == 16 + kw)) ? 0x3C00u : 0u;
                const uint32_t hi = (k < W_S && (c1 == kh || c1 == 16 + kw)) ? 0x3C00u : 0u;
                w[e] = lo | (hi << 16);
            }
            *reinterpret_cast<uint4*>(smem_gen + (sSel - smem_base) + k * 128 + ((j ^ (k & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        // ones operand [4 key blocks][16 rows][64 keys]: row 0 = 1 for keys < 196, every other row 0 (row 0: r & 7 == 0, no swizzle)
        for (int i = threadIdx.x; i < (int)(W_ONES / 16); i += W_THREADS) {
            const int blk = i >> 7, r = (i >> 3) & 15, j = i & 7;
            uint32_t w[4] = {0u, 0u, 0u, 0u};
            if (r == 0) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int k0 = blk * 64 + j * 8 + 2 * e;
                    w[e] = (k0 < W_S ? 0x3C00u : 0u) | ((k0 + 1 < W_S ? 0x3C00u : 0u) << 16);
                }
            }
            *reinterpret_cast<uint4*>(smem_gen + (sOnes - smem_base) + i * 16) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        ptx::fence_proxy_async();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
    auto tS = [&](int g) { return tmem_base + (uint32_t)(g * 256); };

    if (warp == 0) {
        // ===================================================== TMA producer
        if (ptx::elect_one()) {
            ptx::mbar_expect_tx(const_full, 2 * W_RB);
            ptx::tma_load_2d(sR0, &tmR, const_full, 0, 0);
            ptx::tma_load_2d(sRt, &tmR, const_full, 16, 0);
        }
        int w = next_item(0);
        for (int it = 0; w >= 0; ++it) {
            const int w_next = next_item(it + 1);
            const int item = w / heads, head = w - item * heads;
            const int row0 = item * W_S;
            const uint32_t par = (uint32_t)(it & 1);
            ptx::mbar_wait(qk_free, par ^ 1u);     // both groups' S (and G) MMAs of the previous item have completed
            W_TR(0);
            if (ptx::elect_one()) {
                ptx::mbar_expect_tx(qk_full, 4 * W_QB + 2 * W_KB);
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    ptx::tma_load_2d(sQ0 + g * W_QB, &tmQ, qk_full, head * W_HD, row0 + g * W_BQ);
                    ptx::tma_load_2d(sQG + g * W_QB, &tmQ, qk_full, head * W_HD + 64, row0 + g * W_BQ);
                }
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    ptx::tma_load_2d(sK0 + t * 8192, &tmK, qk_full, D + head * W_HD, row0 + t * 64);
                    ptx::tma_load_2d(sKt + t * 8192, &tmK, qk_full, D + head * W_HD + 16, row0 + t * 64);
                }
                ptx::tma_load_2d(sK0 + 3 * 8192, &tmK16, qk_full, D + head * W_HD, row0 + 192);
                ptx::tma_load_2d(sKt + 3 * 8192, &tmK16, qk_full, D + head * W_HD + 16, row0 + 192);
            }
            ptx::mbar_wait(v_free, par ^ 1u);      // both groups' P V of the previous item have completed
            W_TR(1);
            if (ptx::elect_one()) {
                ptx::mbar_expect_tx(v_full, 2 * W_KB);
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    ptx::tma_load_2d(sV0 + t * 8192, &tmK, v_full, 2 * D + head * W_HD, row0 + t * 64);
                    ptx::tma_load_2d(sV1 + t * 8192, &tmK, v_full, 2 * D + head * W_HD + 64, row0 + t * 64);
                }
                ptx::tma_load_2d(sV0 + 3 * 8192, &tmK16, v_full, 2 * D + head * W_HD, row0 + 192);
                ptx::tma_load_2d(sV1 + 3 * 8192, &tmK16, v_full, 2 * D + head * W_HD + 64, row0 + 192);
            }
            // L2 prefetch of the next item's operands: all CTAs reload at about the same time, and a DRAM-latency burst of
            // 148 x 168 KB would otherwise sit between this item's S MMAs and the next item's G
            if (w_next >= 0 && ptx::elect_one()) {
                const int w2 = w_next, item2 = w2 / heads, head2 = w2 - item2 * heads, r2 = item2 * W_S;
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    ptx::tma_prefetch_2d(&tmQ, head2 * W_HD, r2 + g * W_BQ);
                    ptx::tma_prefetch_2d(&tmQ, head2 * W_HD + 64, r2 + g * W_BQ);
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const CUtensorMap* tm = t < 3 ? &tmK : &tmK16;
                    ptx::tma_prefetch_2d(tm, D + head2 * W_HD, r2 + t * 64);
                    ptx::tma_prefetch_2d(tm, D + head2 * W_HD + 64, r2 + t * 64);
                    ptx::tma_prefetch_2d(tm, 2 * D + head2 * W_HD, r2 + t * 64);
                    ptx::tma_prefetch_2d(tm, 2 * D + head2 * W_HD + 64, r2 + t * 64);
                }
            }
            w = w_next;
        }
    } else if (warp == 1 || warp == 2) {
        // ===================================================== MMA issuer of query group g
        const int g = warp - 1;
        // instruction descriptor: D = f32 (bit 4), A = B = f16, N >> 3 at bit 17, M >> 4 at bit 24; bit 16 = B is MN-major
        auto idesc = [](int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(W_BQ >> 4) << 24); };
        const uint32_t idesc_pv = idesc(W_HD) | (1u << 16);
        const uint64_t desc_hi = (2ull << 61) | (1ull << 46) | ((uint64_t)(1024 >> 4) << 32);  // SWIZZLE_128B, SBO 1024
        auto desc = [&](uint32_t addr) { return desc_hi | (uint64_t)((addr >> 4) & 0x3FFF); };
        // V as MN-major operand: 64 head-dim columns (128 B) contiguous per key, 8-key groups 1024 B apart (SBO), the
        // second 64-column block (columns 64..79 used) LBO = sV1 - sV0 further on
        const uint64_t descv_hi = desc_hi | ((uint64_t)((W_KB >> 4) & 0x3FFF) << 16);
        auto descv = [&](uint32_t addr) { return descv_hi | (uint64_t)((addr >> 4) & 0x3FFF); };
        ptx::mbar_wait(const_full, 0);
        const uint64_t a0 = desc(sQ0 + g * W_QB), aq = desc(sQG + g * W_QB);
        // G = Q Rcat^T of the item whose Q has just landed, into the S columns 0..63 of the group
        auto issue_g = [&]() {
            if (ptx::elect_one()) {
                const uint64_t r0 = desc(sR0), rt = desc(sRt);
#pragma unroll
                for (int k = 0; k < 4; ++k) ptx::umma_f16(tS(g), a0 + 2u * k, r0 + 2u * k, idesc(64), k != 0 ? 1u : 0u);
                ptx::umma_f16(tS(g), aq, rt + 6u, idesc(64), 1u);               // head-dim columns 64..79
                ptx::umma_commit(g_full(g));
            }
            __syncwarp();
        };
        int w = next_item(0);
        if (w >= 0) {
            const int it = 0;
            ptx::mbar_wait(qk_full, 0);
            ptx::tc_fence_after();
            W_TR(9 + 8 * g);
            issue_g();
        }
        for (int it = 0; w >= 0; ++it) {
            const int w_next = next_item(it + 1);
            const uint32_t par = (uint32_t)(it & 1);
            // Gsel written -- which also means that every thread of the group has read G and the previous item's O / l
            ptx::mbar_wait(qg_ready(g), par);
            if (g == 1 && it == 0 && skew_clocks > 0) {
                // Put the two query groups in ANTI-PHASE, once: left alone they run in lockstep (same barriers, same item), so both
                // are in the MUFU-bound exponential pass at the same time (3.3 k clocks instead of 1.7 k) while the tensor pipe
                // idles, and then both wait for their MMAs. Nothing re-synchronises them afterwards: the shared gates (K / Q reload
                // after both S MMAs, V reload after both P V) have several thousand clocks of slack.
                const long long t0 = clock64();
                while (clock64() - t0 < (long long)skew_clocks) {}
            }
            ptx::tc_fence_after();
            W_TR(10 + 8 * g);
            if (ptx::elect_one()) {
                const uint64_t b0 = desc(sK0), bt = desc(sKt), bs = desc(sSel);
#pragma unroll
                for (int k = 0; k < 4; ++k) ptx::umma_f16(tS(g), a0 + 2u * k, b0 + 2u * k, idesc(W_KP), k != 0 ? 1u : 0u);
                ptx::umma_f16(tS(g), aq, bt + 6u, idesc(W_KP), 1u);             // Q / K columns 64..79
                ptx::umma_f16(tS(g), aq + 2u, bs, idesc(W_KP), 1u);             // + Gsel_h Sel^T
                ptx::umma_f16(tS(g), aq + 4u, bs + 2u, idesc(W_KP), 1u);        // + Gsel_w Sel^T
                ptx::umma_commit(s_full(g));
                ptx::umma_commit(qk_free);
            }
            __syncwarp();
            ptx::mbar_wait(v_full, par);
            W_TR(11 + 8 * g);
            ptx::mbar_wait(p_full(g), par);
            ptx::tc_fence_after();
            W_TR(12 + 8 * g);
            if (ptx::elect_one()) {
#pragma unroll
                for (int k = 0; k < W_KP / 16; ++k) {   // 13 k-steps of 16 keys; P chunk = 8 packed TMEM columns
                    const uint32_t pa = tS(g) + (k < W_KA / 16 ? 8u * k : W_COL_PB + 8u * (k - W_KA / 16));
                    ptx::umma_f16_ts(tS(g) + W_COL_O, pa, descv(sV0 + (uint32_t)k * 2048u), idesc_pv, k != 0 ? 1u : 0u);
                    ptx::umma_f16_ts(tS(g) + W_COL_L, pa, desc(sOnes + (uint32_t)(k >> 2) * 2048u) + 2u * (k & 3), idesc(16), k != 0 ? 1u : 0u);
                }
                ptx::umma_commit(o_full(g));
                ptx::umma_commit(v_free);
            }
            __syncwarp();
            if (w_next >= 0) {
                // next item's G while the softmax threads store this item's output: Q of the next item landed long ago (its
                // load started when this item's S MMAs completed); P V must have finished reading P out of columns 0..47
                ptx::mbar_wait(qk_full, par ^ 1u);
                W_TR(8 + 8 * g);
                ptx::mbar_wait(o_full(g), par);
                ptx::tc_fence_after();
                W_TR(13 + 8 * g);
                issue_g();
            }
            w = w_next;
        }
    } else if (warp == 3) {
        if (dyn) sched.produce<false>(sched_counter, n_work);
    } else if (warp >= 4) {
        // ===================================================== softmax / output: two threads per query row
        const int idx = warp - 4, quad = warp & 3, grp = (idx >> 2) & 1, half = idx >> 3;
        const int r = quad * 32 + lane;
        const int qi = grp * W_BQ + r;
        const bool row_ok = qi < W_S;
        const bool warp_ok = grp * W_BQ + quad * 32 < W_S;   // warps whose 32 rows are all padding only keep the barrier counts
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        const float sl2 = scale * W_L2E, inv_scale = 1.0f / scale;
        const int qh = (qi < W_S ? qi : 0) / W_G, qw = (qi < W_S ? qi : 0) - qh * W_G;
        const int shift = 13 - (half == 0 ? qh : qw);
        const uint32_t qg_row = sQG + grp * W_QB + (uint32_t)r * 128u;
        const uint32_t sw = (uint32_t)(r & 7);
        float* xch = reinterpret_cast<float*>(smem_gen + (sX - smem_base));
        float* x_mine = xch + (grp * 2 + half) * 128 + r;
        const float* x_peer = xch + (grp * 2 + (half ^ 1)) * 128 + r;
        const int pair_bar = 1 + grp * 4 + quad;
        auto fl = [](uint32_t u) { return __uint_as_float(u); };
        for (int it = 0;; ++it) {
            const int w = next_item(it);
            if (w < 0) break;
            const int item = w / heads, head = w - item * heads;
            const uint32_t par = (uint32_t)(it & 1);
            const uint32_t ts = tS(grp) + lane_off;
            // ---- Gsel: this thread's table half (half 0: rel_h, G columns 0..26; half 1: rel_w, G columns 32..58)
            const int trb = 24 + 8 * grp + 16 * half;
            const bool trw = quad == 0;
            ptx::mbar_wait(g_full(grp), par);
            if (trw) W_TR(trb);
            if (warp_ok) {
                ptx::tc_fence_after();
                uint32_t gv[32];
                ptx::tmem_ld32(ts + 32u * half, gv);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                // in[i] = G[26 - i] / scale, i = 0..26 (in[27] = 0): Gsel[k] = G[qpos + 13 - k] = in[13 - qpos + k]
                uint32_t p[15];
#pragma unroll
                for (int i = 0; i < 13; ++i) p[i] = pack_h2(fl(gv[26 - 2 * i]) * inv_scale, fl(gv[25 - 2 * i]) * inv_scale);
                p[13] = pack_h2(fl(gv[0]) * inv_scale, 0.0f);
                p[14] = 0u;
                uint32_t o7[7];
                barrel14(p, shift, o7);
                // 14 values + 2 zeros = two 16-byte chunks: chunks 2, 3 (half 0) or 4, 5 (half 1) of the QG row
                const uint32_t c0 = 2u + 2u * half;
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(qg_row + (((c0) ^ sw) << 4)), "r"(o7[0]), "r"(o7[1]),
                             "r"(o7[2]), "r"(o7[3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(qg_row + (((c0 + 1u) ^ sw) << 4)), "r"(o7[4]), "r"(o7[5]),
                             "r"(o7[6]), "r"(0u) : "memory");
                ptx::fence_proxy_async();
            }
            ptx::mbar_arrive(qg_ready(grp));
            if (trw) W_TR(trb + 1);
            // ---- softmax over this thread's keys
            ptx::mbar_wait(s_full(grp), par);
            if (trw) W_TR(trb + 2);
            if (warp_ok) {
                ptx::tc_fence_after();
                const uint32_t tk = ts + (half == 0 ? 0u : (uint32_t)W_KA);     // first score column of this half
                const uint32_t tp = ts + (half == 0 ? 0u : W_COL_PB);           // first packed-P column of this half
                float mx = -INFINITY;
#pragma unroll 1
                for (int c = 0; c < 3; ++c) {
                    uint32_t v[32];
                    ptx::tmem_ld32(tk + 32u * c, v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j += 2) mx = fmaxf(mx, fmaxf(fl(v[j]), fl(v[j + 1])));
                }
                if (half == 1) {
                    uint32_t v[16];
                    ptx::tmem_ld16(tk + 96u, v);
                    ptx::tmem_ld_wait();
                    mx = fmaxf(mx, fmaxf(fmaxf(fl(v[0]), fl(v[1])), fmaxf(fl(v[2]), fl(v[3]))));   // keys 192..195; 196..207 are padding
                }
                *x_mine = mx;
                ptx::named_bar_sync(pair_bar, 64);
                mx = fmaxf(mx, *x_peer);
                if (trw) W_TR(trb + 3);
                const float mneg = fmaf(-mx, sl2, -104.0f);   // exponent rebias of exp2_pair_f16 (-112) folded in, and the weights scaled by 2^8
#pragma unroll 1
                for (int c = 0; c < 3; ++c) {
                    uint32_t v[32], pk[16];
                    ptx::tmem_ld32(tk + 32u * c, v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) pk[j] = exp2_pair_f16(fmaf(fl(v[2 * j]), sl2, mneg), fmaf(fl(v[2 * j + 1]), sl2, mneg));
                    ptx::tmem_st16(tp + 16u * c, pk);
                }
                if (half == 1) {
                    uint32_t v[16], pk[16];
                    ptx::tmem_ld16(tk + 96u, v);
                    ptx::tmem_ld_wait();
                    pk[0] = exp2_pair_f16(fmaf(fl(v[0]), sl2, mneg), fmaf(fl(v[1]), sl2, mneg));
                    pk[1] = exp2_pair_f16(fmaf(fl(v[2]), sl2, mneg), fmaf(fl(v[3]), sl2, mneg));
#pragma unroll
                    for (int j = 2; j < 16; ++j) pk[j] = 0u;       // keys 196..207: weight 0
                    ptx::tmem_st16(tp + 48u, pk);
                }
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
            }
            ptx::mbar_arrive(p_full(grp));
            if (trw) W_TR(trb + 4);
            // ---- output: half 0 stores head-dim columns 0..47, half 1 columns 48..79
            ptx::mbar_wait(o_full(grp), par);
            if (trw) W_TR(trb + 5);
            if (warp_ok) {
                ptx::tc_fence_after();
                const uint32_t to = ts + W_COL_O;
                uint32_t l16[16], d0[32], d1[16];
                ptx::tmem_ld16(ts + W_COL_L, l16);     // column 0 = sum of the row's (fp16-rounded) weights
                ptx::tmem_ld32(to + (half == 0 ? 0u : 48u), d0);
                if (half == 0) ptx::tmem_ld16(to + 32u, d1);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                const float inv = 1.0f / fl(l16[0]);
                auto f = [&](uint32_t u) { return __uint_as_float(u) * inv; };
                // output row: window order (item, token), or -- un_g > 0 -- the token's raster row in the un_h x un_w grid of its
                // image (window_unpartition, image_encoder.py:291-318), padding tokens dropped
                long long orow = (long long)item * W_S + qi;
                bool st_ok = row_ok;
                if (un_g > 0) {
                    const int per = un_g * un_g, b = item / per, wi = item - b * per;
                    const int y = (wi / un_g) * W_G + qi / W_G, x = (wi % un_g) * W_G + qi % W_G;
                    st_ok = row_ok && y < un_h && x < un_w;
                    orow = ((long long)b * un_h + y) * un_w + x;
                }
                if (st_ok) {
                    // each thread owns 96 / 64 contiguous bytes of its row: 32-byte stores (one full sector per request)
                    __half* dst = out + orow * D + head * W_HD + (half == 0 ? 0 : 48);
                    auto st32 = [&](__half* ptr, const uint32_t* d) {
                        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(pack_h2(f(d[0]), f(d[1]))),
                                     "r"(pack_h2(f(d[2]), f(d[3]))), "r"(pack_h2(f(d[4]), f(d[5]))), "r"(pack_h2(f(d[6]), f(d[7]))),
                                     "r"(pack_h2(f(d[8]), f(d[9]))), "r"(pack_h2(f(d[10]), f(d[11]))), "r"(pack_h2(f(d[12]), f(d[13]))),
                                     "r"(pack_h2(f(d[14]), f(d[15]))) : "memory");
                    };
                    st32(dst, d0);
                    st32(dst + 16, d0 + 16);
                    if (half == 0) st32(dst + 32, d1);
                }
            }
            if (trw) W_TR(trb + 6);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace

// ------------------------------------------------------------------------------------------ host side
static int g_window_skew = 3000;               // clocks by which query group 1 trails group 0 (window_tc_kernel); debug setter below
static long long* g_window_trace = nullptr;   // debug: device buffer of 8 x 64 clock stamps (cellvit_b200_debug.h)
extern "C" __attribute__((visibility("default"))) void cvb_debug_window_skew(int clocks) { g_window_skew = clocks; }
extern "C" __attribute__((visibility("default"))) void cvb_debug_window_trace(void* dev_buf) { g_window_trace = reinterpret_cast<long long*>(dev_buf); }
bool op_window_attention_tc_supported(int S, int hd, int gh, int gw) { return hd == W_HD && S == W_S && gh == W_G && gw == W_G; }

int op_window_attention_tc(const __half* qkv, int n_items, int heads, int hd, float scale, const __half* relcat, __half* out,
                           int* sched_counter, cudaStream_t stream, int un_g, int un_h, int un_w) {
    CVB_CHECK(qkv && out && relcat, CVB_EARG, "window_attention_tc: null operand");
    CVB_CHECK(un_g == 0 || (un_g > 0 && n_items % (un_g * un_g) == 0 && un_h > 0 && un_w > 0 && un_h <= un_g * W_G && un_w <= un_g * W_G),
              CVB_EARG, "window_attention_tc: bad un-partition geometry");
    CVB_CHECK(hd == W_HD && n_items > 0 && heads > 0, CVB_ESHAPE, "window_attention_tc: needs head dim 80");
    const int D = heads * hd;
    static unsigned long long configured = 0;  // one bit per device: function attributes are per device
    const int cfg_dev = cvb_current_device();
    if (!((configured >> cfg_dev) & 1ull)) {
        CVB_CUDA(cudaFuncSetAttribute(window_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W_SMEM));
        configured |= 1ull << cfg_dev;
    }
    CUtensorMap tq, tk, tk16, tr;
    const uint64_t rows = (uint64_t)n_items * W_S;
    CVB_TRY(cvb_tmap_2d_f16(&tq, qkv, (uint64_t)3 * D, rows, (uint64_t)3 * D * 2, 64, W_BQ));
    CVB_TRY(cvb_tmap_2d_f16(&tk, qkv, (uint64_t)3 * D, rows, (uint64_t)3 * D * 2, 64, 64));
    CVB_TRY(cvb_tmap_2d_f16(&tk16, qkv, (uint64_t)3 * D, rows, (uint64_t)3 * D * 2, 64, 16));
    CVB_TRY(cvb_tmap_2d_f16(&tr, relcat, (uint64_t)W_HD, 64, (uint64_t)W_HD * 2, 64, 64));
    const int n_work = n_items * heads;
    const int grid = n_work < cvb_num_sms() ? n_work : cvb_num_sms();
    window_tc_kernel<<<grid, W_THREADS, W_SMEM, stream>>>(tq, tk, tk16, tr, heads, n_items, scale, out, sched_counter, g_window_skew, g_window_trace, un_g, un_h, un_w);
    cvb_note_launches(1);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}
