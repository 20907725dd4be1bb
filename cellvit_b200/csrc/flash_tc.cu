// flash_tc.cu -- global (non-windowed) attention on the tcgen05 tensor cores, for both encoders:
//
//   SAM (head dim 80, 64-wide token grid):  out = softmax(scale * q k^T + rel_h[q, kh(k)] + rel_w[q, kw(k)]) v
//                                           (image_encoder.py:235-260, 354-392)
//   ViT-S of CellViT-256 (head dim 64, no bias, S = 4097 incl. the cls token):  out = softmax(scale * q k^T) v
//                                           (vits_histo.py:172-188)
//
// per (image, head). Warp-specialised in the style of the tile engine:
//
//   CTA = 2 x 128 queries of one (image, head) sharing every K / V tile of 64 keys. With the SAM bias a key tile is ONE key
//   row of the token grid, so rel_h is a per-query scalar for the whole tile and rel_w[q, 0..63] is the same vector for every
//   tile (kept in registers).
//   warp 0   TMA producer: Q once, then per key tile K and V through a 3-stage ring. Head dim 80 = a 64-column box plus
//            the last k-step out of a second box shifted by 16 columns, so only the K-major 128B-swizzle layout is used for
//            Q and K. V is NOT transposed: the TMA box [64 keys][64 head-dim columns] of the qkv rows is the canonical
//            MN-major 128B-swizzled B operand (8 keys x 64 columns per atom; head dim 80: a second box LBO bytes further on).
//   warps 1, 3   MMA issuers, one per query group: S[t&1] = Q K_t^T (fp32 in TMEM), then O += P_t V_t and l += P_t 1 (row sums from the very fp16
//            weights, via a constant ones operand) once the softmax warps have published P_t; QK of tile t+1 is issued
//            before PV of tile t so the tensor pipe works while tile t is in the softmax.
//   warp 2   TMEM allocator.   warps 4-19  softmax: TWO threads per query row (TMEM lane; each takes 32 of a tile's 64 key
//            columns, the tile maximum is exchanged through shared memory + a 64-thread named barrier), exact online softmax in
//            the log2 domain, P written back over the first 32 columns of its S buffer (packed fp16 pairs) and consumed by
//            P V as a TMEM A operand (TS-mode UMMA: no shared-memory round trip, no A read per instruction). O accumulates
//            in TMEM across all key tiles; the softmax reference m only moves when a score exceeds it by more than 2^8
//            (P stays <= 256 in fp16, sums in fp32 -- mathematically the same softmax), and only then is O rescaled in
//            place (tcgen05.ld / st) -- after the first tile practically never, so the softmax warps never wait for PV.
// Ragged sequences (S = 4097): the qkv rows of consecutive images are contiguous, so boxes that run past an image read the
// next image's rows (zero fill past the tensor); the keys past S of the last tile get weight 0 (score -inf), query rows
// past S are computed and never stored.
// The decomposed rel-pos bias tables rel_h / rel_w [q, 64] come from relpos_tables_kernel (attention.cu: G = Q R^T through
// the MMA path, UNSCALED q, pre-multiplied by log2(e), fp16 as in the mma.sync kernel).
#include "ops.h"

namespace {

constexpr int FT_BQ = 128, FT_BK = 64, FT_STAGES = 3;
constexpr int FT_NG = 2;                                  // query groups (of 128 rows) per CTA
constexpr float FT_L2E = 1.4426950408889634f;
constexpr uint32_t FT_ONES_BYTES = 16 * 128;              // ones operand of the row-sum MMA: 16 rows x 64 keys, row 0 = 1

// NG query groups of 128 rows per CTA share every K / V tile; each group has its own S double buffer, O accumulator
// and eight softmax warps (NG = 2: 16 softmax warps, four per scheduler).
template <int NG, int HD>
struct FtCfg {
    static constexpr int BOXES = HD > 64 ? 2 : 1;          // 64-column boxes per Q / K / V row block
    static constexpr int THREADS = 128 + NG * 256;          // 4 control warps + 8 softmax warps per query group (two threads per row)
    static constexpr uint32_t Q_BYTES = NG * BOXES * FT_BQ * 128;
    static constexpr uint32_t K_BYTES = BOXES * FT_BK * 128;
    static constexpr uint32_t V_BYTES = BOXES * FT_BK * 128;
    static constexpr uint32_t X_BYTES = NG * 2 * 128 * 4;   // tile-maximum exchange between the two threads of a row
    static constexpr uint32_t SMEM = 1024 + Q_BYTES + FT_STAGES * (K_BYTES + V_BYTES) + FT_ONES_BYTES + X_BYTES + 512;
    static constexpr uint32_t O_STRIDE = 96;               // TMEM columns per group's O: HD output columns, then the row sum
};

template <int NG, int HD, bool BIAS>
__global__ void __launch_bounds__(FtCfg<NG, HD>::THREADS, 1)
flash_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __half* __restrict__ bias_h, const __half* __restrict__ bias_w,
                int S, int heads, float scale, __half* __restrict__ out) {
    using Cfg = FtCfg<NG, HD>;
    static_assert(HD == 64 || HD == 80, "head dim 64 (ViT-S) or 80 (SAM)");
    constexpr int KS = HD / 16;                            // k-steps of Q K^T
    extern __shared__ uint8_t ft_smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(ft_smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = ft_smem_raw + (smem_base - ptx::smem_u32(ft_smem_raw));
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int ghd = blockIdx.y, g = ghd / heads, head = ghd - g * heads;
    const int q0 = blockIdx.x * (NG * FT_BQ);
    const int D = heads * HD;
    const int n_t = (S + FT_BK - 1) / FT_BK;

    const uint32_t sQ = smem_base;
    const uint32_t sK = sQ + Cfg::Q_BYTES;
    const uint32_t sV = sK + FT_STAGES * Cfg::K_BYTES;
    const uint32_t sOnes = sV + FT_STAGES * Cfg::V_BYTES;
    const uint32_t sX = sOnes + FT_ONES_BYTES;
    const uint32_t bar = sX + Cfg::X_BYTES;
    // barriers (8 B each); per-group ones are indexed by gb = group * 2 + buffer
    const uint32_t q_full = bar;
    auto kv_full = [&](int st) { return bar + 8u * (1 + st); };
    auto kv_empty = [&](int st) { return bar + 8u * (4 + st); };
    auto s_full = [&](int gb) { return bar + 8u * (7 + gb); };
    auto s_empty = [&](int gb) { return bar + 8u * (7 + 2 * NG + gb); };
    auto p_full = [&](int gb) { return bar + 8u * (7 + 4 * NG + gb); };
    auto o_full = [&](int gb) { return bar + 8u * (7 + 6 * NG + gb); };
    const uint32_t tmem_slot = bar + 8u * (7 + 8 * NG);

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmQ);
        ptx::prefetch_tmap(&tmK);
    }
    if (warp == 1 && lane == 0) {
        ptx::mbar_init(q_full, 1);
        for (int st = 0; st < FT_STAGES; ++st) { ptx::mbar_init(kv_full(st), 1); ptx::mbar_init(kv_empty(st), NG); }   // one P V commit per group's issuer
        for (int gb = 0; gb < 2 * NG; ++gb) {
            ptx::mbar_init(s_full(gb), 1);
            ptx::mbar_init(s_empty(gb), 256);
            ptx::mbar_init(p_full(gb), 256);
            ptx::mbar_init(o_full(gb), 1);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    }
    if (warp == 3) {
        // ones operand [16 rows][64 keys], K-major 128B swizzle: row 0 (r & 7 == 0: no swizzle) = 1.0, rows 1..15 = 0
        for (int i = lane; i < (int)(FT_ONES_BYTES / 16); i += 32) {
            const uint32_t w = (i >> 3) == 0 ? 0x3C003C00u : 0u;
            *reinterpret_cast<uint4*>(smem_gen + (sOnes - smem_base) + i * 16) = make_uint4(w, w, w, w);
        }
        ptx::fence_proxy_async();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
    // TMEM columns: S[group][buffer] at (group*2 + buffer)*64, O[group] at NG*128 + group*96 (HD used, then the row sum)
    auto tS = [&](int gb) { return tmem_base + (uint32_t)(gb * 64); };
    auto tO = [&](int grp) { return tmem_base + (uint32_t)(NG * 128 + grp * Cfg::O_STRIDE); };

    if (warp == 0) {
        // ===================================================== TMA producer
        const int row_q = g * S + q0;
        if (ptx::elect_one()) {
            ptx::mbar_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
            for (int grp = 0; grp < NG; ++grp) {
                ptx::tma_load_2d(sQ + grp * Cfg::BOXES * FT_BQ * 128, &tmQ, q_full, head * HD, row_q + grp * FT_BQ);
                if (Cfg::BOXES == 2)
                    ptx::tma_load_2d(sQ + grp * 2 * FT_BQ * 128 + FT_BQ * 128, &tmQ, q_full, head * HD + 16, row_q + grp * FT_BQ);
            }
        }
        int stage = 0;
        uint32_t phase = 0;
        for (int t = 0; t < n_t; ++t) {
            ptx::mbar_wait(kv_empty(stage), phase ^ 1u);
            if (ptx::elect_one()) {
                ptx::mbar_expect_tx(kv_full(stage), Cfg::K_BYTES + Cfg::V_BYTES);
                const int row_k = g * S + t * FT_BK;
                ptx::tma_load_2d(sK + stage * Cfg::K_BYTES, &tmK, kv_full(stage), D + head * HD, row_k);
                ptx::tma_load_2d(sV + stage * Cfg::V_BYTES, &tmK, kv_full(stage), 2 * D + head * HD, row_k);
                if (Cfg::BOXES == 2) {
                    ptx::tma_load_2d(sK + stage * Cfg::K_BYTES + FT_BK * 128, &tmK, kv_full(stage), D + head * HD + 16, row_k);
                    ptx::tma_load_2d(sV + stage * Cfg::V_BYTES + FT_BK * 128, &tmK, kv_full(stage), 2 * D + head * HD + 64, row_k);
                }
            }
            if (++stage == FT_STAGES) { stage = 0; phase ^= 1u; }
        }
    } else if (warp == 1 || (warp == 3 && NG == 2)) {
        // ===================================================== MMA issuers: one warp per query group (warp 1: group 0, warp 3: group 1).
        // With ONE issuer the in-order waits of the two groups' barriers (s_empty, p_full) serialised them; each group is now
        // its own pipeline, the tensor pipe interleaves them.
        const int grp = warp == 1 ? 0 : 1;
        // instruction descriptors: D=f32, A=B=f16; N>>3 @17, M>>4 @24; bit 16: B is MN-major (V read in place)
        const uint32_t idesc_qk = (1u << 4) | ((uint32_t)(FT_BK >> 3) << 17) | ((uint32_t)(FT_BQ >> 4) << 24);
        const uint32_t idesc_pv = (1u << 4) | (1u << 16) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(FT_BQ >> 4) << 24);
        const uint32_t idesc_l = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(FT_BQ >> 4) << 24);
        const uint64_t desc_hi = (2ull << 61) | (1ull << 46) | ((uint64_t)(1024 >> 4) << 32);  // SWIZZLE_128B, SBO 1024
        auto desc = [&](uint32_t addr) { return desc_hi | (uint64_t)((addr >> 4) & 0x3FFF); };
        // V: 64 head-dim columns (128 B) per key, 8-key groups 1024 B apart (SBO), second 64-column block LBO = 8 KB further on
        const uint64_t descv_hi = desc_hi | ((uint64_t)(((FT_BK * 128) >> 4) & 0x3FFF) << 16);
        auto descv = [&](uint32_t addr) { return descv_hi | (uint64_t)((addr >> 4) & 0x3FFF); };
        ptx::mbar_wait(q_full, 0);
        ptx::tc_fence_after();
        const uint32_t q = sQ + grp * Cfg::BOXES * FT_BQ * 128;
        auto issue_qk = [&](int t) {
            const int stage = t % FT_STAGES, gb = grp * 2 + (t & 1);
            ptx::mbar_wait(kv_full(stage), (uint32_t)((t / FT_STAGES) & 1));
            ptx::mbar_wait(s_empty(gb), (uint32_t)(((t >> 1) & 1) ^ 1));
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint64_t a0 = desc(q), b0 = desc(sK + stage * Cfg::K_BYTES);
#pragma unroll
                for (int k = 0; k < 4; ++k) ptx::umma_f16(tS(gb), a0 + 2u * k, b0 + 2u * k, idesc_qk, k != 0 ? 1u : 0u);
                if (KS == 5) {  // hd columns 64-79 = columns 48-63 of the second (shifted) box
                    const uint64_t a1 = desc(q + FT_BQ * 128), b1 = desc(sK + stage * Cfg::K_BYTES + FT_BK * 128);
                    ptx::umma_f16(tS(gb), a1 + 6u, b1 + 6u, idesc_qk, 1u);
                }
                ptx::umma_commit(s_full(gb));
            }
            __syncwarp();
        };
        auto issue_pv = [&](int t) {
            const int stage = t % FT_STAGES, gb = grp * 2 + (t & 1);
            ptx::mbar_wait(p_full(gb), (uint32_t)((t >> 1) & 1));
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint32_t v = sV + stage * Cfg::V_BYTES;
                const uint64_t ones = desc(sOnes);
#pragma unroll
                for (int k = 0; k < 4; ++k) {  // 16 keys per k-step: 2 KB of V rows; P chunk = 8 packed TMEM columns
                    ptx::umma_f16_ts(tO(grp), tS(gb) + 8u * k, descv(v + (uint32_t)k * 2048u), idesc_pv, (t | k) != 0 ? 1u : 0u);
                    ptx::umma_f16_ts(tO(grp) + HD, tS(gb) + 8u * k, ones + 2u * k, idesc_l, (t | k) != 0 ? 1u : 0u);
                }
                ptx::umma_commit(o_full(gb));
                ptx::umma_commit(kv_empty(stage));
            }
            __syncwarp();
        };
        issue_qk(0);
        for (int t = 0; t < n_t; ++t) {
            if (t + 1 < n_t) issue_qk(t + 1);
            issue_pv(t);
        }
    } else if (warp >= 4) {
        // ===================================================== softmax / output: TWO threads per query row (column-split)
        // warps 4 .. 4 + 8 NG - 1: idx = warp - 4 -> half = idx / (4 NG) takes key columns 32 half .. 32 half + 31 of every tile,
        // grp = (idx / 4) % NG, TMEM lane quadrant = warp & 3. One thread per row ran at 40 % issue utilisation (2 warps per
        // scheduler, ~390 dependent instructions per tile); two threads per row halve the chain and double the warps.
        const int idx = warp - 4, quad = warp & 3, grp = (idx >> 2) % NG, half = idx / (4 * NG), r = quad * 32 + lane;
        const int qrow = q0 + grp * FT_BQ + r;                         // query index inside the image
        const bool row_ok = qrow < S;
        const long long row = (long long)ghd * S + (row_ok ? qrow : S - 1);
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        const float sl2 = scale * FT_L2E;
        float* xch = reinterpret_cast<float*>(smem_gen + (sX - smem_base));
        float* x_mine = xch + (grp * 2 + half) * 128 + r;
        const float* x_peer = xch + (grp * 2 + (half ^ 1)) * 128 + r;
        const int pair_bar = 1 + grp * 4 + quad;
        // rel_w[q, 32 half .. 32 half + 31] (log2 domain) packed as 16 half2 registers
        uint32_t bw[BIAS ? 16 : 1];
        const __half* bh_row = nullptr;
        unsigned short bh_raw = 0;
        if (BIAS) {
            const uint4* p = reinterpret_cast<const uint4*>(bias_w + row * 64 + 32 * half);
#pragma unroll
            for (int i = 0; i < (BIAS ? 4 : 0); ++i) {
                const uint4 v = __ldg(p + i);
                bw[4 * i] = v.x; bw[4 * i + 1] = v.y; bw[4 * i + 2] = v.z; bw[4 * i + 3] = v.w;
            }
            bh_row = bias_h + row * 64;
            // the per-tile rel_h scalar is fetched one tile ahead and left untouched (raw fp16) until the next iteration, so the
            // L2 round trip never sits on the critical path (converting it right away stalled every tile on the load)
            bh_raw = __ldg(reinterpret_cast<const unsigned short*>(bh_row));
        }
        float m_ref = -INFINITY;
        for (int t = 0; t < n_t; ++t) {
            const int gb = grp * 2 + (t & 1);
            float bh = 0.0f;
            if (BIAS) {
                bh = __half2float(__ushort_as_half(bh_raw));
                if (t + 1 < n_t) bh_raw = __ldg(reinterpret_cast<const unsigned short*>(bh_row) + t + 1);
            }
            ptx::mbar_wait(s_full(gb), (uint32_t)((t >> 1) & 1));
            ptx::tc_fence_after();
            uint32_t v0[32];
            ptx::tmem_ld32(tS(gb) + lane_off + 32u * half, v0);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            // scores in the log2 domain (without the per-tile scalar bh)
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float a0, a1;
                if (BIAS) {
                    const float2 w0 = __half22float2(*reinterpret_cast<const __half2*>(&bw[BIAS ? j : 0]));
                    a0 = fmaf(__uint_as_float(v0[2 * j]), sl2, w0.x); a1 = fmaf(__uint_as_float(v0[2 * j + 1]), sl2, w0.y);
                } else {
                    a0 = __uint_as_float(v0[2 * j]) * sl2; a1 = __uint_as_float(v0[2 * j + 1]) * sl2;
                }
                v0[2 * j] = __float_as_uint(a0); v0[2 * j + 1] = __float_as_uint(a1);
            }
            if (t == n_t - 1 && S - t * FT_BK < FT_BK) {
                // ragged last tile: keys past the end of the sequence are the next image's rows -> weight 0
                const int valid = S - t * FT_BK - 32 * half;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j >= valid) v0[j] = __float_as_uint(-INFINITY);
            }
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; j += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(v0[j]), __uint_as_float(v0[j + 1])));
            // the two threads of a row agree on the tile maximum (and thereby on every decision below); the barrier also orders
            // BOTH threads' score loads before either one's P store, which lands in the partner's score columns
            *x_mine = mx;
            ptx::named_bar_sync(pair_bar, 64);
            mx = fmaxf(mx, *x_peer);
            ptx::mbar_arrive(s_empty(gb));
            // lazy reference update: move m only when this tile exceeds it by more than 8 (a factor 256)
            const float cand = mx + bh;
            const bool move = cand > m_ref + 8.0f;  // always true on the first tile (m_ref = -inf)
            if (__any_sync(0xffffffffu, move) && t > 0) {
                // rare: rescale the accumulated O (TMEM) and l of the rows that moved (each thread of the pair its half of the
                // columns); all PV issued so far must be done
                const float alpha = move ? ptx::ex2(m_ref - cand) : 1.0f;
                ptx::mbar_wait(o_full(grp * 2 + ((t - 1) & 1)), (uint32_t)(((t - 1) >> 1) & 1));
                ptx::tc_fence_after();
#pragma unroll 1
                for (int c0 = 48 * half; c0 < 48 * half + 48; c0 += 16) {  // HD output columns + the row-sum column: 96 in all
                    uint32_t d[16];
                    ptx::tmem_ld16(tO(grp) + lane_off + (uint32_t)c0, d);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) d[i] = __float_as_uint(__uint_as_float(d[i]) * alpha);
                    ptx::tmem_st16(tO(grp) + lane_off + (uint32_t)c0, d);
                }
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
            }
            if (move) m_ref = cand;
            // + the exponent rebias of exp2_pair_f16, - 6: the row's largest weight lies in [2^6, 2^14], so the fp32 flush of the
            // MUFU (weights below 2^-14) only drops what is below 2^-20 of it
            const float mref = m_ref - bh + (112.0f - 6.0f);
            // P = 2^(t - m) as packed fp16 pairs (fp32 MUFU + integer packing, common.cuh). The row sum is not accumulated here:
            // the ones operand makes it output column HD of the P V pass, from exactly these rounded weights.
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = exp2_pair_f16(__uint_as_float(v0[2 * j]) - mref, __uint_as_float(v0[2 * j + 1]) - mref);
            // P row (2 x 16 packed fp16 pairs) over the first 32 columns of this S buffer: the A operand of the TS-mode P V MMA
            ptx::tmem_st16(tS(gb) + lane_off + 16u * half, pk);
            ptx::tmem_st_wait();
            ptx::tc_fence_before();
            ptx::mbar_arrive(p_full(gb));
        }
        {   // all key tiles accumulated: O / l -- half 0 stores head-dim columns [0, HD/2 rounded to 16), half 1 the rest
            ptx::mbar_wait(o_full(grp * 2 + ((n_t - 1) & 1)), (uint32_t)(((n_t - 1) >> 1) & 1));
            ptx::tc_fence_after();
            float inv;
            {
                uint32_t d[16];
                ptx::tmem_ld16(tO(grp) + lane_off + (uint32_t)HD, d);  // column HD = sum of the row's weights
                ptx::tmem_ld_wait();
                inv = 1.0f / __uint_as_float(d[0]);
            }
            __half* dst = out + ((long long)g * S + (row_ok ? qrow : 0)) * D + head * HD;
            auto f = [&](uint32_t u) { return __uint_as_float(u) * inv; };
            constexpr int SPLIT = HD == 80 ? 48 : 32;
            const int c_begin = half == 0 ? 0 : SPLIT, c_end = half == 0 ? SPLIT : HD;
#pragma unroll 1
            for (int c0 = c_begin; c0 < c_end; c0 += 16) {
                uint32_t d[16];
                ptx::tmem_ld16(tO(grp) + lane_off + (uint32_t)c0, d);
                ptx::tmem_ld_wait();
                if (row_ok) {
#pragma unroll
                    for (int i = 0; i < 16; i += 8)
                        *reinterpret_cast<uint4*>(dst + c0 + i) = make_uint4(pack_h2(f(d[i]), f(d[i + 1])), pack_h2(f(d[i + 2]), f(d[i + 3])),
                                                                             pack_h2(f(d[i + 4]), f(d[i + 5])), pack_h2(f(d[i + 6]), f(d[i + 7])));
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

template <int HD, bool BIAS>
int launch_flash_tc(const CUtensorMap& tq, const CUtensorMap& tk, const __half* bias_h, const __half* bias_w, int Gb, int S, int heads,
                    float scale, __half* out, cudaStream_t stream) {
    using Cfg = FtCfg<FT_NG, HD>;
    static unsigned long long configured = 0;  // one bit per device: function attributes are per device
    const int cfg_dev = cvb_current_device();
    if (!((configured >> cfg_dev) & 1ull)) {
        CVB_CUDA(cudaFuncSetAttribute(flash_tc_kernel<FT_NG, HD, BIAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
        configured |= 1ull << cfg_dev;
    }
    const dim3 grid((S + FT_NG * FT_BQ - 1) / (FT_NG * FT_BQ), Gb * heads);
    flash_tc_kernel<FT_NG, HD, BIAS><<<grid, Cfg::THREADS, Cfg::SMEM, stream>>>(tq, tk, bias_h, bias_w, S, heads, scale, out);
    cvb_note_launches(1);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}

}  // namespace

bool op_attention_tc_supported(int S, int hd, const __half* Rh, int gh, int gw) {
    if (hd == 64 && Rh == nullptr) return S >= FT_BK;                                   // ViT-S: no bias, any length
    return hd == 80 && Rh != nullptr && gw == FT_BK && gh <= 64 && gh * gw == S;        // SAM: a key tile = one row of the token grid
}

size_t op_attention_tc_workspace_bytes(int Gb, int S, int heads) {
    const size_t bias = align_up((size_t)Gb * heads * S * 64 * 2, 1024);  // the two rel-pos bias tables of a SAM call
    return 2 * bias + 1024;
}

int op_attention_tc(const __half* qkv, int Gb, int S, int heads, int hd, float scale, const __half* Rh, const __half* Rw, int gh,
                    int gw, __half* out, void* workspace, size_t ws_bytes, cudaStream_t stream) {
    CVB_CHECK(qkv && out, CVB_EARG, "attention_tc: null operand");
    CVB_CHECK((Rh == nullptr) == (Rw == nullptr), CVB_EARG, "attention_tc: Rh and Rw must both be set or both null");
    CVB_CHECK(op_attention_tc_supported(S, hd, Rh, gh, gw), CVB_ESHAPE,
              "attention_tc: needs head dim 64 without bias, or head dim 80 with rel-pos tables on a 64-wide token grid (S=%d hd=%d grid %dx%d)",
              S, hd, gh, gw);
    const int D = heads * hd;
    CUtensorMap tq, tk;
    CVB_TRY(cvb_tmap_2d_f16(&tq, qkv, (uint64_t)3 * D, (uint64_t)Gb * S, (uint64_t)3 * D * 2, 64, FT_BQ));
    CVB_TRY(cvb_tmap_2d_f16(&tk, qkv, (uint64_t)3 * D, (uint64_t)Gb * S, (uint64_t)3 * D * 2, 64, FT_BK));
    if (!Rh) return launch_flash_tc<64, false>(tq, tk, nullptr, nullptr, Gb, S, heads, scale, out, stream);
    CVB_CHECK(workspace != nullptr && ws_bytes >= op_attention_tc_workspace_bytes(Gb, S, heads), CVB_EWORKSPACE, "attention_tc: workspace too small");
    CVB_CHECK(((uintptr_t)workspace & 1023) == 0, CVB_EARG, "attention_tc: workspace must be 1024-byte aligned");
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    const size_t bias_b = align_up((size_t)Gb * heads * S * 64 * 2, 1024);
    __half* bias_h = reinterpret_cast<__half*>(ws);
    __half* bias_w = reinterpret_cast<__half*>(ws + bias_b);
    CVB_TRY(op_relpos_tables(qkv, Gb, S, heads, hd, Rh, Rw, gh, gw, bias_h, bias_w, stream));
    return launch_flash_tc<80, true>(tq, tk, bias_h, bias_w, Gb, S, heads, scale, out, stream);
}
