// flash_tc.cu -- global (non-windowed) attention of the SAM encoder on the tcgen05 tensor cores.
//
//   out = softmax(scale * q k^T + rel_h[q, kh(k)] + rel_w[q, kw(k)]) v     (image_encoder.py:235-260, 354-392)
//
// per (image, head): S = 4096 keys (64 x 64 token grid), head dim 80. Replaces the mma.sync flash kernel for this
// shape (246 TFLOP/s, 44 % of the legacy-HMMA ceiling) with a warp-specialised kernel in the style of the tile engine:
//
//   CTA = 2 x 128 queries of one (image, head) sharing every K / V tile; key tiles of 64 keys = ONE key row of the token grid, so rel_h is a per-query
//   scalar for the whole tile and rel_w[q, 0..63] is the same vector for every tile (kept in registers).
//   warp 0   TMA producer: Q once (two 64-column boxes: hd columns 0-63 and 16-79 -- the fifth k-step reads columns
//            64-79 out of the second box, so only the proven 128B-swizzle K-major layout is used), then per key
//            tile K (same two boxes) and V^T (80 x 64 keys, K-major) through a 3-stage ring.
//   warp 1   MMA issuer: S[t&1] = Q K_t^T (5 x UMMA 128x64x16, fp32 in TMEM), then O[t&1] = P_t V_t (4 x UMMA 128x80x16)
//            once the softmax warps have published P_t; QK of tile t+1 is issued before PV of tile t so the tensor
//            pipe works while tile t is in the softmax.
//   warp 2   TMEM allocator.   warps 4-7  softmax: one query row per thread (TMEM lane), exact online softmax in
//            the log2 domain, P written back over the first 32 columns of its S buffer (packed fp16 pairs) and consumed by
//            P V as a TMEM A operand (TS-mode UMMA: no shared-memory round trip, no A read per instruction). O accumulates
//            in TMEM across all key tiles; the softmax reference m only moves when a score exceeds it by more than 2^8
//            (P stays <= 256 in fp16, sums in fp32 -- mathematically the same softmax), and only then is O rescaled in
//            place (tcgen05.ld / st) -- after the first tile practically never, so the softmax warps never wait for PV.
// V must be K-major for the B operand of P V, i.e. transposed to [hd, keys]: v_transpose_kernel does that once per
// block (42 MB). The decomposed rel-pos bias tables rel_h / rel_w [q, 64] come from relpos_tables_kernel (attention.cu:
// G = Q R^T through the MMA path, UNSCALED q, pre-multiplied by log2(e), fp16 as in the mma.sync kernel).
#include <cudaTypedefs.h>

#include "ops.h"

namespace {

constexpr int FT_BQ = 128, FT_BK = 64, FT_HD = 80, FT_STAGES = 3;
constexpr uint32_t FT_K_BYTES = 2 * FT_BK * 128;          // two boxes of 64 rows x 128 B
constexpr uint32_t FT_V_BYTES = FT_HD * 128;              // 80 rows (hd) x 64 keys (window kernel)
constexpr int FT_VR = 96;                                 // global kernel: V^T rows per head = 80 + a row of ones (row sums of P
                                                          // come out of the P V MMA as output column 80) + 15 zero rows
constexpr uint32_t FTG_V_BYTES = FT_VR * 128;
constexpr bool FT_P_TMEM = true;                          // P V reads P from tensor memory (TS-mode UMMA) instead of shared memory
constexpr int FT_NG = 2;                                  // query groups (of 128 rows) per CTA
constexpr float FT_L2E = 1.4426950408889634f;

// ------------------------------------------------------------------------------------------ V^T
// v rows [Gb*S, 3*D] (columns 2*D + head*hd + d) -> vt [(g*heads + head)*96 + d][S]; row 80 = ones, rows 81..95 = zeros
__global__ void __launch_bounds__(256)
v_transpose_kernel(const __half* __restrict__ qkv, int S, int heads, __half* __restrict__ vt) {
    __shared__ __half tile[64][FT_HD + 2];
    const int gh = blockIdx.y, g = gh / heads, head = gh - g * heads;
    const int D = heads * FT_HD, k0 = blockIdx.x * 64;
    const __half* src = qkv + ((long long)g * S + k0) * 3 * D + 2 * D + head * FT_HD;
    for (int i = threadIdx.x; i < 64 * (FT_HD / 8); i += 256) {
        const int r = i / (FT_HD / 8), c = i - r * (FT_HD / 8);
        const uint4 v = *reinterpret_cast<const uint4*>(src + (long long)r * 3 * D + c * 8);
        const __half* h = reinterpret_cast<const __half*>(&v);
#pragma unroll
        for (int j = 0; j < 8; ++j) tile[r][c * 8 + j] = h[j];
    }
    __syncthreads();
    __half* dst = vt + (long long)gh * FT_VR * S + k0;
    for (int i = threadIdx.x; i < FT_VR * 32; i += 256) {
        const int d = i >> 5, kp = i & 31;
        __half2 v = __floats2half2_rn(0.f, 0.f);
        if (d < FT_HD) v = __halves2half2(tile[2 * kp][d], tile[2 * kp + 1][d]);
        else if (d == FT_HD) v = __floats2half2_rn(1.f, 1.f);
        *reinterpret_cast<__half2*>(dst + (long long)d * S + 2 * kp) = v;
    }
}

// ------------------------------------------------------------------------------------------ main kernel
// NG query groups of 128 rows per CTA share every K / V^T tile; each group has its own S double buffer, O accumulator,
// P double buffer and four softmax warps (NG = 2: 8 softmax warps, two per scheduler).
template <int NG>
struct FtCfg {
    static constexpr int THREADS = 128 + NG * 128;
    static constexpr uint32_t Q_BYTES = NG * 2 * FT_BQ * 128;
    static constexpr uint32_t P_BYTES = FT_BQ * 128;
    static constexpr uint32_t SMEM = 1024 + Q_BYTES + FT_STAGES * (FT_K_BYTES + FTG_V_BYTES) + NG * 2 * P_BYTES + 512;
};

template <int NG>
__global__ void __launch_bounds__(FtCfg<NG>::THREADS, 1)
flash_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __half* __restrict__ bias_h, const __half* __restrict__ bias_w,
                int S, int heads, float scale, __half* __restrict__ out) {
    using Cfg = FtCfg<NG>;
    extern __shared__ uint8_t ft_smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(ft_smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = ft_smem_raw + (smem_base - ptx::smem_u32(ft_smem_raw));
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int ghd = blockIdx.y, g = ghd / heads, head = ghd - g * heads;
    const int q0 = blockIdx.x * (NG * FT_BQ);
    const int D = heads * FT_HD;
    const int n_t = S / FT_BK;

    const uint32_t sQ = smem_base;
    const uint32_t sK = sQ + Cfg::Q_BYTES;
    const uint32_t sV = sK + FT_STAGES * FT_K_BYTES;
    const uint32_t sP = sV + FT_STAGES * FTG_V_BYTES;
    const uint32_t bar = sP + NG * 2 * Cfg::P_BYTES;
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
        ptx::prefetch_tmap(&tmV);
    }
    if (warp == 1 && lane == 0) {
        ptx::mbar_init(q_full, 1);
        for (int st = 0; st < FT_STAGES; ++st) { ptx::mbar_init(kv_full(st), 1); ptx::mbar_init(kv_empty(st), 1); }
        for (int gb = 0; gb < 2 * NG; ++gb) {
            ptx::mbar_init(s_full(gb), 1);
            ptx::mbar_init(s_empty(gb), 128);
            ptx::mbar_init(p_full(gb), 128);
            ptx::mbar_init(o_full(gb), 1);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
    // TMEM columns: S[group][buffer] at (group*2 + buffer)*64, O[group] at NG*128 + group*96 (80 used)
    auto tS = [&](int gb) { return tmem_base + (uint32_t)(gb * 64); };
    auto tO = [&](int grp) { return tmem_base + (uint32_t)(NG * 128 + grp * 96); };

    if (warp == 0) {
        // ===================================================== TMA producer
        const int row_q = g * S + q0;
        if (ptx::elect_one()) {
            ptx::mbar_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
            for (int grp = 0; grp < NG; ++grp) {
                ptx::tma_load_2d(sQ + grp * 2 * FT_BQ * 128, &tmQ, q_full, head * FT_HD, row_q + grp * FT_BQ);
                ptx::tma_load_2d(sQ + grp * 2 * FT_BQ * 128 + FT_BQ * 128, &tmQ, q_full, head * FT_HD + 16, row_q + grp * FT_BQ);
            }
        }
        int stage = 0;
        uint32_t phase = 0;
        for (int t = 0; t < n_t; ++t) {
            ptx::mbar_wait(kv_empty(stage), phase ^ 1u);
            if (ptx::elect_one()) {
                ptx::mbar_expect_tx(kv_full(stage), FT_K_BYTES + FTG_V_BYTES);
                const int row_k = g * S + t * FT_BK;
                ptx::tma_load_2d(sK + stage * FT_K_BYTES, &tmK, kv_full(stage), D + head * FT_HD, row_k);
                ptx::tma_load_2d(sK + stage * FT_K_BYTES + FT_BK * 128, &tmK, kv_full(stage), D + head * FT_HD + 16, row_k);
                ptx::tma_load_2d(sV + stage * FTG_V_BYTES, &tmV, kv_full(stage), t * FT_BK, ghd * FT_VR);
            }
            if (++stage == FT_STAGES) { stage = 0; phase ^= 1u; }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        // instruction descriptors: D=f32, A=B=f16, both K-major; N>>3 @17, M>>4 @24
        const uint32_t idesc_qk = (1u << 4) | ((uint32_t)(FT_BK >> 3) << 17) | ((uint32_t)(FT_BQ >> 4) << 24);
        const uint32_t idesc_pv = (1u << 4) | ((uint32_t)(FT_VR >> 3) << 17) | ((uint32_t)(FT_BQ >> 4) << 24);
        const uint64_t desc_hi = (2ull << 61) | (1ull << 46) | ((uint64_t)(1024 >> 4) << 32);  // SWIZZLE_128B, SBO 1024
        auto desc = [&](uint32_t addr) { return desc_hi | (uint64_t)((addr >> 4) & 0x3FFF); };
        ptx::mbar_wait(q_full, 0);
        ptx::tc_fence_after();
        auto issue_qk = [&](int t) {
            const int stage = t % FT_STAGES, b = t & 1;
            ptx::mbar_wait(kv_full(stage), (uint32_t)((t / FT_STAGES) & 1));
#pragma unroll
            for (int grp = 0; grp < NG; ++grp) {
                const int gb = grp * 2 + b;
                ptx::mbar_wait(s_empty(gb), (uint32_t)(((t >> 1) & 1) ^ 1));
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    const uint32_t q = sQ + grp * 2 * FT_BQ * 128;
                    const uint64_t a0 = desc(q), a1 = desc(q + FT_BQ * 128);
                    const uint64_t b0 = desc(sK + stage * FT_K_BYTES), b1 = desc(sK + stage * FT_K_BYTES + FT_BK * 128);
#pragma unroll
                    for (int k = 0; k < 4; ++k) ptx::umma_f16(tS(gb), a0 + 2u * k, b0 + 2u * k, idesc_qk, k != 0 ? 1u : 0u);
                    ptx::umma_f16(tS(gb), a1 + 6u, b1 + 6u, idesc_qk, 1u);  // hd columns 64-79 = columns 48-63 of the second box
                    ptx::umma_commit(s_full(gb));
                }
                __syncwarp();
            }
        };
        auto issue_pv = [&](int t) {
            const int stage = t % FT_STAGES, b = t & 1;
#pragma unroll
            for (int grp = 0; grp < NG; ++grp) {
                const int gb = grp * 2 + b;
                ptx::mbar_wait(p_full(gb), (uint32_t)((t >> 1) & 1));
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    const uint64_t a0 = desc(sP + gb * Cfg::P_BYTES), b0 = desc(sV + stage * FTG_V_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (FT_P_TMEM) ptx::umma_f16_ts(tO(grp), tS(gb) + 8u * k, b0 + 2u * k, idesc_pv, (t | k) != 0 ? 1u : 0u);
                        else ptx::umma_f16(tO(grp), a0 + 2u * k, b0 + 2u * k, idesc_pv, (t | k) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(o_full(gb));
                    if (grp == NG - 1) ptx::umma_commit(kv_empty(stage));
                }
                __syncwarp();
            }
        };
        issue_qk(0);
        for (int t = 0; t < n_t; ++t) {
            if (t + 1 < n_t) issue_qk(t + 1);
            issue_pv(t);
        }
    } else if (warp >= 4) {
        // ===================================================== softmax / output: one query row per thread
        const int grp = (warp - 4) >> 2, quad = warp & 3, r = quad * 32 + lane;
        const long long row = (long long)ghd * S + q0 + grp * FT_BQ + r;
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        const float sl2 = scale * FT_L2E;
        // rel_w[q, 0..63] (log2 domain) packed as 32 half2 registers
        uint32_t bw[32];
        {
            const uint4* p = reinterpret_cast<const uint4*>(bias_w + row * 64);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint4 v = __ldg(p + i);
                bw[4 * i] = v.x; bw[4 * i + 1] = v.y; bw[4 * i + 2] = v.z; bw[4 * i + 3] = v.w;
            }
        }
        const __half* bh_row = bias_h + row * 64;
        float m_ref = -INFINITY;
        // the per-tile rel_h scalar is fetched one tile ahead and left untouched (raw fp16) until the next iteration, so the
        // L2 round trip never sits on the critical path (converting it right away stalled every tile on the load)
        unsigned short bh_raw = __ldg(reinterpret_cast<const unsigned short*>(bh_row));
        const uint32_t p_row = (uint32_t)r * 128u;
        const uint32_t sw = (uint32_t)(r & 7);
        for (int t = 0; t < n_t; ++t) {
            const int gb = grp * 2 + (t & 1);
            const float bh = __half2float(__ushort_as_half(bh_raw));
            if (t + 1 < n_t) bh_raw = __ldg(reinterpret_cast<const unsigned short*>(bh_row) + t + 1);
            ptx::mbar_wait(s_full(gb), (uint32_t)((t >> 1) & 1));
            ptx::tc_fence_after();
            uint32_t v0[32], v1[32];
            ptx::tmem_ld32(tS(gb) + lane_off, v0);
            ptx::tmem_ld32(tS(gb) + lane_off + 32u, v1);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            ptx::mbar_arrive(s_empty(gb));
            // scores in the log2 domain (without the per-tile scalar bh), and their maximum
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float2 w0 = __half22float2(*reinterpret_cast<const __half2*>(&bw[j]));
                const float2 w1 = __half22float2(*reinterpret_cast<const __half2*>(&bw[16 + j]));
                const float a0 = fmaf(__uint_as_float(v0[2 * j]), sl2, w0.x), a1 = fmaf(__uint_as_float(v0[2 * j + 1]), sl2, w0.y);
                const float c0 = fmaf(__uint_as_float(v1[2 * j]), sl2, w1.x), c1 = fmaf(__uint_as_float(v1[2 * j + 1]), sl2, w1.y);
                v0[2 * j] = __float_as_uint(a0); v0[2 * j + 1] = __float_as_uint(a1);
                v1[2 * j] = __float_as_uint(c0); v1[2 * j + 1] = __float_as_uint(c1);
                mx = fmaxf(mx, fmaxf(fmaxf(a0, a1), fmaxf(c0, c1)));
            }
            // lazy reference update: move m only when this tile exceeds it by more than 8 (a factor 256)
            const float cand = mx + bh;
            const bool move = cand > m_ref + 8.0f;  // always true on the first tile (m_ref = -inf)
            if (__any_sync(0xffffffffu, move) && t > 0) {
                // rare: rescale the accumulated O (TMEM) and l of the rows that moved; all PV issued so far must be done
                const float alpha = move ? ptx::ex2(m_ref - cand) : 1.0f;
                ptx::mbar_wait(o_full(grp * 2 + ((t - 1) & 1)), (uint32_t)(((t - 1) >> 1) & 1));
                ptx::tc_fence_after();
#pragma unroll 1
                for (int c0 = 0; c0 < FT_VR; c0 += 16) {  // 80 output columns + the row-sum column
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
            const float mref = m_ref - bh;
            // P = 2^(t - m) straight in fp16 pairs (one MUFU op per two weights; they are rounded to fp16 for the MMA anyway,
            // and the fp16 rounding of the exponent only matters for weights that are negligible). The row sum is not
            // accumulated here: the ones row of V^T makes it output column 80 of P V, from exactly these rounded weights.
            uint32_t pk[32];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                pk[j] = ptx::ex2_f16x2(pack_h2(__uint_as_float(v0[2 * j]) - mref, __uint_as_float(v0[2 * j + 1]) - mref));
                pk[16 + j] = ptx::ex2_f16x2(pack_h2(__uint_as_float(v1[2 * j]) - mref, __uint_as_float(v1[2 * j + 1]) - mref));
            }
            if (FT_P_TMEM) {
                // P row (32 packed fp16 pairs) over the first 32 columns of this S buffer: the A operand of the TS-mode P V MMA
                ptx::tmem_st32(tS(gb) + lane_off, pk);
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
            } else {
                // P row (64 keys fp16 = 8 chunks of 16 B) into the K-major SWIZZLE_128B A operand: chunk c -> c ^ (row & 7)
                const uint32_t base = sP + gb * Cfg::P_BYTES + p_row;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t addr = base + (((uint32_t)c ^ sw) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * c]), "r"(pk[4 * c + 1]),
                                 "r"(pk[4 * c + 2]), "r"(pk[4 * c + 3]) : "memory");
                }
                ptx::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
            }
            ptx::mbar_arrive(p_full(gb));
        }
        {   // all key tiles accumulated: O / l
            ptx::mbar_wait(o_full(grp * 2 + ((n_t - 1) & 1)), (uint32_t)(((n_t - 1) >> 1) & 1));
            ptx::tc_fence_after();
            float inv;
            {
                uint32_t d[16];
                ptx::tmem_ld16(tO(grp) + lane_off + (uint32_t)FT_HD, d);  // column 80 = sum of the row's weights
                ptx::tmem_ld_wait();
                inv = 1.0f / __uint_as_float(d[0]);
            }
            __half* dst = out + ((long long)g * S + q0 + grp * FT_BQ + r) * D + head * FT_HD;
            auto f = [&](uint32_t u) { return __uint_as_float(u) * inv; };
#pragma unroll 1
            for (int c0 = 0; c0 < FT_HD; c0 += 16) {
                uint32_t d[16];
                ptx::tmem_ld16(tO(grp) + lane_off + (uint32_t)c0, d);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i += 8)
                    *reinterpret_cast<uint4*>(dst + c0 + i) = make_uint4(pack_h2(f(d[i]), f(d[i + 1])), pack_h2(f(d[i + 2]), f(d[i + 3])),
                                                                         pack_h2(f(d[i + 4]), f(d[i + 5])), pack_h2(f(d[i + 6]), f(d[i + 7])));
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------ windowed attention (14 x 14)
// Same machinery for the SAM windows: CTA = one (window, head): 196 queries as two groups of 128 rows (rows past 196 are
// the next window's tokens: computed, never stored), 196 keys as tiles of 64 + 64 + 64 + 16 (the last tile is a
// 128x16x16 UMMA; keys 196..207 get P = 0). The decomposed rel-pos bias is produced in the kernel: G = Q Rcat^T is one
// more UMMA (Rcat = [rel_h rows | rel_w rows], 64 x 80, K-major, parked in the O columns of TMEM before P V starts);
// each thread keeps bh[kh] = G[qh + 13 - kh] and bw[kw] = G[32 + qw + 13 - kw] in registers, and because the key
// tile loop is fully unrolled (kh, kw) of every score are compile-time constants. V^T comes from
// v_transpose_win_kernel in a per-window layout padded to 208 keys (16-byte aligned TMA boxes).
constexpr int WT_S = 196, WT_G = 14, WT_NT = 4, WT_VLD = 208, WT_STAGES = 2;
constexpr uint32_t WT_Q_BYTES = 2 * 2 * FT_BQ * 128;
constexpr uint32_t WT_R_BYTES = 2 * 64 * 128;
constexpr uint32_t WT_P_BYTES = FT_BQ * 128;
constexpr uint32_t WT_SMEM = 1024 + WT_Q_BYTES + WT_R_BYTES + WT_STAGES * (FT_K_BYTES + FT_V_BYTES) + 4 * WT_P_BYTES + 512;
constexpr int WT_THREADS = 384;

// v rows of window `item` [196, 3*D] -> vt [(head*hd + d)][item*208 + key], keys 196..207 zero
__global__ void __launch_bounds__(256)
v_transpose_win_kernel(const __half* __restrict__ qkv, int heads, int n_items, __half* __restrict__ vt) {
    __shared__ __half tile[WT_VLD][FT_HD + 2];
    const int item = blockIdx.x, head = blockIdx.y;
    const int D = heads * FT_HD;
    const __half* src = qkv + (long long)item * WT_S * 3 * D + 2 * D + head * FT_HD;
    for (int i = threadIdx.x; i < WT_VLD * (FT_HD / 8); i += 256) {
        const int r = i / (FT_HD / 8), c = i - r * (FT_HD / 8);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r < WT_S) v = *reinterpret_cast<const uint4*>(src + (long long)r * 3 * D + c * 8);
        const __half* h = reinterpret_cast<const __half*>(&v);
#pragma unroll
        for (int j = 0; j < 8; ++j) tile[r][c * 8 + j] = h[j];
    }
    __syncthreads();
    const long long ld = (long long)n_items * WT_VLD;
    __half* dst = vt + (long long)head * FT_HD * ld + (long long)item * WT_VLD;
    for (int i = threadIdx.x; i < FT_HD * (WT_VLD / 2); i += 256) {
        const int d = i / (WT_VLD / 2), kp = i - d * (WT_VLD / 2);
        *reinterpret_cast<__half2*>(dst + (long long)d * ld + 2 * kp) = __halves2half2(tile[2 * kp][d], tile[2 * kp + 1][d]);
    }
}

__global__ void __launch_bounds__(WT_THREADS, 1)
window_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmR, int heads, int n_items, float scale,
                 __half* __restrict__ out) {
    extern __shared__ uint8_t ft_smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(ft_smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = ft_smem_raw + (smem_base - ptx::smem_u32(ft_smem_raw));
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    // persistent: CTA c works on (window, head) pairs c, c + gridDim.x, ... -- TMEM, barriers and the rel-pos operand are set
    // up once, and the K / V ring runs ahead into the next pair. Every ring / buffer barrier completes an even number
    // of phases per pair (4 key tiles, 2 buffers), so their parities are functions of the tile index alone; the
    // once-per-pair barriers (q_full, q_empty, g_full, o_read) use the parity of the pair counter `it`.
    const int D = heads * FT_HD;
    const int n_work = n_items * heads;

    const uint32_t sQ = smem_base;
    const uint32_t sR = sQ + WT_Q_BYTES;
    const uint32_t sK = sR + WT_R_BYTES;
    const uint32_t sV = sK + WT_STAGES * FT_K_BYTES;
    const uint32_t sP = sV + WT_STAGES * FT_V_BYTES;
    const uint32_t bar = sP + 4 * WT_P_BYTES;
    const uint32_t q_full = bar;  // Q (both groups) + Rcat
    auto kv_full = [&](int st) { return bar + 8u * (1 + st); };
    auto kv_empty = [&](int st) { return bar + 8u * (3 + st); };
    auto s_full = [&](int gb) { return bar + 8u * (5 + gb); };
    auto s_empty = [&](int gb) { return bar + 8u * (9 + gb); };
    auto p_full = [&](int gb) { return bar + 8u * (13 + gb); };
    auto o_full = [&](int gb) { return bar + 8u * (17 + gb); };
    auto g_full = [&](int grp) { return bar + 8u * (21 + grp); };
    auto o_read = [&](int grp) { return bar + 8u * (23 + grp); };
    const uint32_t q_empty = bar + 8u * 25, r_full = bar + 8u * 26;
    const uint32_t tmem_slot = bar + 8u * 27;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmQ);
        ptx::prefetch_tmap(&tmK);
        ptx::prefetch_tmap(&tmV);
        ptx::prefetch_tmap(&tmR);
    }
    if (warp == 1 && lane == 0) {
        ptx::mbar_init(q_full, 1);
        for (int st = 0; st < WT_STAGES; ++st) { ptx::mbar_init(kv_full(st), 1); ptx::mbar_init(kv_empty(st), 1); }
        for (int gb = 0; gb < 4; ++gb) {
            ptx::mbar_init(s_full(gb), 1);
            ptx::mbar_init(s_empty(gb), 128);
            ptx::mbar_init(p_full(gb), 128);
            ptx::mbar_init(o_full(gb), 1);
        }
        ptx::mbar_init(g_full(0), 1);
        ptx::mbar_init(g_full(1), 1);
        ptx::mbar_init(o_read(0), 128);
        ptx::mbar_init(o_read(1), 128);
        ptx::mbar_init(q_empty, 1);
        ptx::mbar_init(r_full, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
    auto tS = [&](int gb) { return tmem_base + (uint32_t)(gb * 64); };
    auto tO = [&](int grp) { return tmem_base + (uint32_t)(256 + grp * 96); };

    if (warp == 0) {
        // ===================================================== TMA producer
        if (ptx::elect_one()) {
            ptx::mbar_expect_tx(r_full, WT_R_BYTES);
            ptx::tma_load_2d(sR, &tmR, r_full, 0, 0);
            ptx::tma_load_2d(sR + 64 * 128, &tmR, r_full, 16, 0);
        }
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
            const int item = w / heads, head = w - item * heads;
            const int row0 = item * WT_S;
            ptx::mbar_wait(q_empty, (uint32_t)((it & 1) ^ 1));  // all Q K^T / G MMAs of the previous pair have read Q
            if (ptx::elect_one()) {
                ptx::mbar_expect_tx(q_full, WT_Q_BYTES);
#pragma unroll
                for (int grp = 0; grp < 2; ++grp) {
                    ptx::tma_load_2d(sQ + grp * 2 * FT_BQ * 128, &tmQ, q_full, head * FT_HD, row0 + grp * FT_BQ);
                    ptx::tma_load_2d(sQ + grp * 2 * FT_BQ * 128 + FT_BQ * 128, &tmQ, q_full, head * FT_HD + 16, row0 + grp * FT_BQ);
                }
            }
            for (int t = 0; t < WT_NT; ++t) {
                ptx::mbar_wait(kv_empty(stage), phase ^ 1u);
                if (ptx::elect_one()) {
                    ptx::mbar_expect_tx(kv_full(stage), FT_K_BYTES + FT_V_BYTES);
                    const int row_k = row0 + t * FT_BK;
                    ptx::tma_load_2d(sK + stage * FT_K_BYTES, &tmK, kv_full(stage), D + head * FT_HD, row_k);
                    ptx::tma_load_2d(sK + stage * FT_K_BYTES + FT_BK * 128, &tmK, kv_full(stage), D + head * FT_HD + 16, row_k);
                    ptx::tma_load_2d(sV + stage * FT_V_BYTES, &tmV, kv_full(stage), item * WT_VLD + t * FT_BK, head * FT_HD);
                }
                if (++stage == WT_STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        auto idesc = [](int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(FT_BQ >> 4) << 24); };
        const uint64_t desc_hi = (2ull << 61) | (1ull << 46) | ((uint64_t)(1024 >> 4) << 32);
        auto desc = [&](uint32_t addr) { return desc_hi | (uint64_t)((addr >> 4) & 0x3FFF); };
        ptx::mbar_wait(r_full, 0);
        auto issue_qk = [&](int t) {
            const int stage = t % WT_STAGES, b = t & 1;
            const int n = t == WT_NT - 1 ? 16 : 64;
            ptx::mbar_wait(kv_full(stage), (uint32_t)((t / WT_STAGES) & 1));
#pragma unroll
            for (int grp = 0; grp < 2; ++grp) {
                const int gb = grp * 2 + b;
                ptx::mbar_wait(s_empty(gb), (uint32_t)(((t >> 1) & 1) ^ 1));
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    const uint32_t q = sQ + grp * 2 * FT_BQ * 128;
                    const uint64_t a0 = desc(q), a1 = desc(q + FT_BQ * 128);
                    const uint64_t b0 = desc(sK + stage * FT_K_BYTES), b1 = desc(sK + stage * FT_K_BYTES + FT_BK * 128);
#pragma unroll
                    for (int k = 0; k < 4; ++k) ptx::umma_f16(tS(gb), a0 + 2u * k, b0 + 2u * k, idesc(n), k != 0 ? 1u : 0u);
                    ptx::umma_f16(tS(gb), a1 + 6u, b1 + 6u, idesc(n), 1u);
                    ptx::umma_commit(s_full(gb));
                }
                __syncwarp();
            }
        };
        auto issue_pv = [&](int t) {
            const int stage = t % WT_STAGES, b = t & 1;
            const int ksteps = t == WT_NT - 1 ? 1 : 4;
#pragma unroll
            for (int grp = 0; grp < 2; ++grp) {
                const int gb = grp * 2 + b;
                ptx::mbar_wait(p_full(gb), (uint32_t)((t >> 1) & 1));
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    const uint64_t b0 = desc(sV + stage * FT_V_BYTES);  // P: TMEM A operand, first columns of the S buffer
                    for (int k = 0; k < ksteps; ++k) ptx::umma_f16_ts(tO(grp), tS(gb) + 8u * k, b0 + 2u * k, idesc(FT_HD), (t | k) != 0 ? 1u : 0u);
                    ptx::umma_commit(o_full(gb));
                    if (grp == 1) ptx::umma_commit(kv_empty(stage));
                }
                __syncwarp();
            }
        };
        int it = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
            ptx::mbar_wait(q_full, (uint32_t)(it & 1));
            // G = Q Rcat^T into the O columns of each group, once the previous pair's output has been read out of them
            ptx::mbar_wait(o_read(0), (uint32_t)((it & 1) ^ 1));
            ptx::mbar_wait(o_read(1), (uint32_t)((it & 1) ^ 1));
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
#pragma unroll
                for (int grp = 0; grp < 2; ++grp) {
                    const uint32_t q = sQ + grp * 2 * FT_BQ * 128;
                    const uint64_t a0 = desc(q), a1 = desc(q + FT_BQ * 128), b0 = desc(sR), b1 = desc(sR + 64 * 128);
#pragma unroll
                    for (int k = 0; k < 4; ++k) ptx::umma_f16(tO(grp), a0 + 2u * k, b0 + 2u * k, idesc(64), k != 0 ? 1u : 0u);
                    ptx::umma_f16(tO(grp), a1 + 6u, b1 + 6u, idesc(64), 1u);
                    ptx::umma_commit(g_full(grp));
                }
            }
            __syncwarp();
            issue_qk(0);
            for (int t = 0; t < WT_NT; ++t) {
                if (t + 1 < WT_NT) issue_qk(t + 1);
                if (t + 1 == WT_NT - 1) {  // the last Q K^T of this pair has been issued: Q may be overwritten once it completes
                    if (ptx::elect_one()) ptx::umma_commit(q_empty);
                    __syncwarp();
                }
                issue_pv(t);
            }
        }
    } else if (warp >= 4) {
        // ===================================================== softmax / output: one query row per thread
        const int grp = (warp - 4) >> 2, quad = warp & 3, r = quad * 32 + lane;
        const int qi = grp * FT_BQ + r;                 // query index inside the window (rows >= 196: next window, not stored)
        const int qc = qi < WT_S ? qi : WT_S - 1;
        const int qh = qc / WT_G, qw = qc - qh * WT_G;
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        const float sl2 = scale * FT_L2E;
        int it = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
        const int item = w / heads, head = w - item * heads;
        const int row0 = item * WT_S;
        float bh[WT_G], bw[WT_G];
        {
            // this thread's G row: stash it in shared memory (aliasing the group's two P buffers, 256 B per row), then
            // gather the 14 + 14 values this query needs
            ptx::mbar_wait(g_full(grp), (uint32_t)(it & 1));
            ptx::tc_fence_after();
            uint32_t g0[32], g1[32];
            ptx::tmem_ld32(tO(grp) + lane_off, g0);
            ptx::tmem_ld32(tO(grp) + lane_off + 32u, g1);
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            float* gs = reinterpret_cast<float*>(smem_gen + (sP - smem_base) + grp * 2 * WT_P_BYTES) + r * 64;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                *reinterpret_cast<uint4*>(gs + i) = make_uint4(g0[i], g0[i + 1], g0[i + 2], g0[i + 3]);
                *reinterpret_cast<uint4*>(gs + 32 + i) = make_uint4(g1[i], g1[i + 1], g1[i + 2], g1[i + 3]);
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < WT_G; ++k) {
                bh[k] = gs[qh + WT_G - 1 - k] * FT_L2E;
                bw[k] = gs[32 + qw + WT_G - 1 - k] * FT_L2E;
            }
            ptx::named_bar_sync(1 + grp, 128);  // every row of the group is read before the first P row is written
        }
        float m_ref = -INFINITY, l_run = 0.f;
        const uint32_t p_row = (uint32_t)r * 128u;
        const uint32_t sw = (uint32_t)(r & 7);
#pragma unroll
        for (int t = 0; t < WT_NT; ++t) {
            const int gb = grp * 2 + (t & 1);
            constexpr int NC_FULL = 64;
            const int ncol = t == WT_NT - 1 ? 16 : NC_FULL;
            ptx::mbar_wait(s_full(gb), (uint32_t)((t >> 1) & 1));
            ptx::tc_fence_after();
            uint32_t v[64];
            if (t == WT_NT - 1) {
                uint32_t d[16];
                ptx::tmem_ld16(tS(gb) + lane_off, d);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = d[j];
            } else {
                uint32_t a[32], c[32];
                ptx::tmem_ld32(tS(gb) + lane_off, a);
                ptx::tmem_ld32(tS(gb) + lane_off + 32u, c);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) { v[j] = a[j]; v[32 + j] = c[j]; }
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(s_empty(gb));
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                const int k = t * 64 + j;
                if (j < ncol && k < WT_S) {
                    const float a = fmaf(__uint_as_float(v[j]), sl2, bh[k / WT_G] + bw[k % WT_G]);
                    v[j] = __float_as_uint(a);
                    mx = fmaxf(mx, a);
                }
            }
            const bool move = mx > m_ref + 8.0f;
            if (__any_sync(0xffffffffu, move) && t > 0) {
                const float alpha = move ? ptx::ex2(m_ref - mx) : 1.0f;
                ptx::mbar_wait(o_full(grp * 2 + ((t - 1) & 1)), (uint32_t)(((t - 1) >> 1) & 1));
                ptx::tc_fence_after();
#pragma unroll 1
                for (int c0 = 0; c0 < FT_HD; c0 += 16) {
                    uint32_t d[16];
                    ptx::tmem_ld16(tO(grp) + lane_off + (uint32_t)c0, d);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) d[i] = __float_as_uint(__uint_as_float(d[i]) * alpha);
                    ptx::tmem_st16(tO(grp) + lane_off + (uint32_t)c0, d);
                }
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
                l_run *= alpha;
            }
            if (move) m_ref = mx;
            float rs = 0.f;
            uint32_t pk[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int k0 = t * 64 + 2 * j;
                float p0 = 0.f, p1 = 0.f;
                if (2 * j < ncol && k0 < WT_S) p0 = ptx::ex2(__uint_as_float(v[2 * j]) - m_ref);
                if (2 * j + 1 < ncol && k0 + 1 < WT_S) p1 = ptx::ex2(__uint_as_float(v[2 * j + 1]) - m_ref);
                rs += p0 + p1;
                pk[j] = pack_h2(p0, p1);
            }
            l_run += rs;
            // P (packed fp16 pairs) over the first columns of this S buffer: TMEM A operand of the P V MMA
            if (t == WT_NT - 1) {
                uint32_t d[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) d[j] = pk[j];
                ptx::tmem_st16(tS(gb) + lane_off, d);
            } else {
                ptx::tmem_st32(tS(gb) + lane_off, pk);
            }
            ptx::tmem_st_wait();
            ptx::tc_fence_before();
            ptx::mbar_arrive(p_full(gb));
        }
        {
            ptx::mbar_wait(o_full(grp * 2 + ((WT_NT - 1) & 1)), (uint32_t)(((WT_NT - 1) >> 1) & 1));
            ptx::tc_fence_after();
            const float inv = 1.0f / l_run;
            __half* dst = out + ((long long)row0 + qi) * D + head * FT_HD;
            auto f = [&](uint32_t u) { return __uint_as_float(u) * inv; };
#pragma unroll 1
            for (int c0 = 0; c0 < FT_HD; c0 += 16) {
                uint32_t d[16];
                ptx::tmem_ld16(tO(grp) + lane_off + (uint32_t)c0, d);
                ptx::tmem_ld_wait();
                if (qi < WT_S) {
#pragma unroll
                    for (int i = 0; i < 16; i += 8)
                        *reinterpret_cast<uint4*>(dst + c0 + i) = make_uint4(pack_h2(f(d[i]), f(d[i + 1])), pack_h2(f(d[i + 2]), f(d[i + 3])),
                                                                             pack_h2(f(d[i + 4]), f(d[i + 5])), pack_h2(f(d[i + 6]), f(d[i + 7])));
                }
            }
            ptx::tc_fence_before();
            ptx::mbar_arrive(o_read(grp));  // the O columns may take the next pair's G
        }
        }  // work items
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------ windowed attention, single-shot
// Second design for the windows: the whole 196-key score row is ONE UMMA per k-step (S = Q K^T with N = 208), so the
// softmax warps run through the keys without waiting for the tensor pipe between key tiles (the four-tile loop above
// exposes a ~4.5 k-clock Q K^T -> softmax -> P V latency per tile). Per (window, head) and query group:
//   G = Q Rcat^T (N = 64) into S columns 0..63 -> threads stash their row, gather bh / bw -> S = Q K^T over them ->
//   exact row maximum over the 196 biased scores -> key chunks 192..207, 128..191, 64..127, 0..63 (in that order): P
//   chunk to shared memory (2-buffer ring) -> P V for the chunk. O lives in S columns 128..207, which are dead once
//   the first two chunks have been read, so S (208) + O fit in 256 TMEM columns per group.
// K (208 rows, two 64-column boxes) and V^T (4 x 64 keys) stay resident per pair; the next pair's Q / K load as soon as
// this pair's Q K^T has completed.
constexpr uint32_t W2_KA_BYTES = WT_VLD * 128;                                  // 208 key rows x 128 B per box
constexpr uint32_t W2_V_BYTES = 4 * FT_V_BYTES;
constexpr uint32_t W2_SMEM = 1024 + WT_Q_BYTES + 2 * W2_KA_BYTES + W2_V_BYTES + 4 * WT_P_BYTES + 512;

__global__ void __launch_bounds__(WT_THREADS, 1)
window_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmK16, const __grid_constant__ CUtensorMap tmV,
                  const __grid_constant__ CUtensorMap tmR, int heads, int n_items, float scale, __half* __restrict__ out) {
    extern __shared__ uint8_t ft_smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(ft_smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = ft_smem_raw + (smem_base - ptx::smem_u32(ft_smem_raw));
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int D = heads * FT_HD;
    const int n_work = n_items * heads;

    const uint32_t sQ = smem_base;
    const uint32_t sKa = sQ + WT_Q_BYTES;
    const uint32_t sKb = sKa + W2_KA_BYTES;
    const uint32_t sV = sKb + W2_KA_BYTES;
    const uint32_t sP = sV + W2_V_BYTES;   // [group][buffer] 16 KB each; the first 16 KB double as the Rcat operand
    const uint32_t sR = sP;
    const uint32_t bar = sP + 4 * WT_P_BYTES;
    // once-per-pair barriers (parity = pair counter & 1)
    const uint32_t q_full = bar, k_full = bar + 8, v_full = bar + 16, qk_done = bar + 24, v_done = bar + 32, r_full = bar + 40,
                   r_done = bar + 48;
    auto g_full = [&](int g) { return bar + 8u * (8 + g); };
    auto g_read = [&](int g) { return bar + 8u * (10 + g); };
    auto s_full = [&](int g) { return bar + 8u * (12 + g); };
    auto o_full = [&](int g) { return bar + 8u * (14 + g); };
    auto o_read = [&](int g) { return bar + 8u * (16 + g); };
    auto p_full = [&](int g, int c) { return bar + 8u * (18 + g * 4 + c); };   // chunk c of group g published
    auto p_free = [&](int g, int b) { return bar + 8u * (26 + g * 2 + b); };   // two completions per pair
    const uint32_t tmem_slot = bar + 8u * 30;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmK16); ptx::prefetch_tmap(&tmV); ptx::prefetch_tmap(&tmR);
    }
    if (warp == 1 && lane == 0) {
        ptx::mbar_init(q_full, 1); ptx::mbar_init(k_full, 1); ptx::mbar_init(v_full, 1); ptx::mbar_init(qk_done, 1);
        ptx::mbar_init(v_done, 1); ptx::mbar_init(r_full, 1); ptx::mbar_init(r_done, 1);
        for (int g = 0; g < 2; ++g) {
            ptx::mbar_init(g_full(g), 1); ptx::mbar_init(g_read(g), 128); ptx::mbar_init(s_full(g), 1);
            ptx::mbar_init(o_full(g), 1); ptx::mbar_init(o_read(g), 128);
            for (int c = 0; c < 4; ++c) ptx::mbar_init(p_full(g, c), 128);
            for (int b = 0; b < 2; ++b) ptx::mbar_init(p_free(g, b), 1);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
    auto tS = [&](int g) { return tmem_base + (uint32_t)(g * 256); };
    auto tO = [&](int g) { return tmem_base + (uint32_t)(g * 256 + 128); };

    if (warp == 0) {
        // ===================================================== TMA producer
        int it = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
            const int item = w / heads, head = w - item * heads;
            const int row0 = item * WT_S;
            const uint32_t par = (uint32_t)(it & 1);
            // Rcat lives in the P region: reload it for every pair once the previous pair's P V has finished with P
            ptx::mbar_wait(v_done, par ^ 1u);
            if (ptx::elect_one()) {
                ptx::mbar_expect_tx(r_full, WT_R_BYTES);
                ptx::tma_load_2d(sR, &tmR, r_full, 0, 0);
                ptx::tma_load_2d(sR + 64 * 128, &tmR, r_full, 16, 0);
                ptx::mbar_expect_tx(v_full, W2_V_BYTES);
#pragma unroll
                for (int t = 0; t < 4; ++t) ptx::tma_load_2d(sV + t * FT_V_BYTES, &tmV, v_full, item * WT_VLD + t * FT_BK, head * FT_HD);
            }
            ptx::mbar_wait(qk_done, par ^ 1u);  // Q / K of the previous pair are no longer read
            if (ptx::elect_one()) {
                ptx::mbar_expect_tx(q_full, WT_Q_BYTES);
#pragma unroll
                for (int grp = 0; grp < 2; ++grp) {
                    ptx::tma_load_2d(sQ + grp * 2 * FT_BQ * 128, &tmQ, q_full, head * FT_HD, row0 + grp * FT_BQ);
                    ptx::tma_load_2d(sQ + grp * 2 * FT_BQ * 128 + FT_BQ * 128, &tmQ, q_full, head * FT_HD + 16, row0 + grp * FT_BQ);
                }
                ptx::mbar_expect_tx(k_full, 2 * W2_KA_BYTES);
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    ptx::tma_load_2d(sKa + t * 8192, &tmK, k_full, D + head * FT_HD, row0 + t * 64);
                    ptx::tma_load_2d(sKb + t * 8192, &tmK, k_full, D + head * FT_HD + 16, row0 + t * 64);
                }
                ptx::tma_load_2d(sKa + 3 * 8192, &tmK16, k_full, D + head * FT_HD, row0 + 192);
                ptx::tma_load_2d(sKb + 3 * 8192, &tmK16, k_full, D + head * FT_HD + 16, row0 + 192);
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        auto idesc = [](int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(FT_BQ >> 4) << 24); };
        const uint64_t desc_hi = (2ull << 61) | (1ull << 46) | ((uint64_t)(1024 >> 4) << 32);
        auto desc = [&](uint32_t addr) { return desc_hi | (uint64_t)((addr >> 4) & 0x3FFF); };
        int it = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
            const uint32_t par = (uint32_t)(it & 1);
            ptx::mbar_wait(q_full, par);
            ptx::mbar_wait(r_full, par);
            // G = Q Rcat^T into S columns 0..63 of each group (the previous pair's O, in the same columns region, has been read)
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                ptx::mbar_wait(o_read(g), par ^ 1u);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    const uint32_t q = sQ + g * 2 * FT_BQ * 128;
                    const uint64_t a0 = desc(q), a1 = desc(q + FT_BQ * 128), b0 = desc(sR), b1 = desc(sR + 64 * 128);
#pragma unroll
                    for (int k = 0; k < 4; ++k) ptx::umma_f16(tS(g), a0 + 2u * k, b0 + 2u * k, idesc(64), k != 0 ? 1u : 0u);
                    ptx::umma_f16(tS(g), a1 + 6u, b1 + 6u, idesc(64), 1u);
                    ptx::umma_commit(g_full(g));
                    if (g == 1) ptx::umma_commit(r_done);
                }
                __syncwarp();
            }
            ptx::mbar_wait(k_full, par);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                ptx::mbar_wait(g_read(g), par);  // every thread of the group has its G row
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    const uint32_t q = sQ + g * 2 * FT_BQ * 128;
                    const uint64_t a0 = desc(q), a1 = desc(q + FT_BQ * 128), b0 = desc(sKa), b1 = desc(sKb);
#pragma unroll
                    for (int k = 0; k < 4; ++k) ptx::umma_f16(tS(g), a0 + 2u * k, b0 + 2u * k, idesc(WT_VLD), k != 0 ? 1u : 0u);
                    ptx::umma_f16(tS(g), a1 + 6u, b1 + 6u, idesc(WT_VLD), 1u);
                    ptx::umma_commit(s_full(g));
                    if (g == 1) ptx::umma_commit(qk_done);
                }
                __syncwarp();
            }
            ptx::mbar_wait(v_full, par);
            // P V per key chunk, chunks 3, 2, 1, 0; O (S columns 128..207) may only be written once chunks 3 AND 2 were read
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                ptx::mbar_wait(p_full(g, 3), par);
                ptx::mbar_wait(p_full(g, 2), par);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    ptx::umma_f16(tO(g), desc(sP + (g * 2 + 1) * WT_P_BYTES), desc(sV + 3 * FT_V_BYTES), idesc(FT_HD), 0u);  // keys 192..207
                    ptx::umma_commit(p_free(g, 1));
                    const uint64_t a0 = desc(sP + (g * 2 + 0) * WT_P_BYTES), b0 = desc(sV + 2 * FT_V_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k) ptx::umma_f16(tO(g), a0 + 2u * k, b0 + 2u * k, idesc(FT_HD), 1u);
                    ptx::umma_commit(p_free(g, 0));
                }
                __syncwarp();
            }
#pragma unroll
            for (int c = 1; c >= 0; --c) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    ptx::mbar_wait(p_full(g, c), par);
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {
                        const uint64_t a0 = desc(sP + (g * 2 + (c & 1)) * WT_P_BYTES), b0 = desc(sV + c * FT_V_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; ++k) ptx::umma_f16(tO(g), a0 + 2u * k, b0 + 2u * k, idesc(FT_HD), 1u);
                        if (c == 1) ptx::umma_commit(p_free(g, 1));
                        else { ptx::umma_commit(p_free(g, 0)); ptx::umma_commit(o_full(g)); if (g == 1) ptx::umma_commit(v_done); }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 4) {
        // ===================================================== softmax / output: one query row per thread
        const int grp = (warp - 4) >> 2, quad = warp & 3, r = quad * 32 + lane;
        const int qi = grp * FT_BQ + r;
        const int qc = qi < WT_S ? qi : WT_S - 1;
        const int qh = qc / WT_G, qw = qc - qh * WT_G;
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        const float sl2 = scale * FT_L2E;
        const uint32_t p_row = (uint32_t)r * 128u;
        const uint32_t sw = (uint32_t)(r & 7);
        int it = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
            const int item = w / heads, head = w - item * heads;
            const int row0 = item * WT_S;
            const uint32_t par = (uint32_t)(it & 1);
            float bh[WT_G], bw[WT_G];
            {
                ptx::mbar_wait(g_full(grp), par);
                ptx::tc_fence_after();
                uint32_t g0[32], g1[32];
                ptx::tmem_ld32(tS(grp) + lane_off, g0);
                ptx::tmem_ld32(tS(grp) + lane_off + 32u, g1);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                ptx::mbar_arrive(g_read(grp));  // S columns 0..63 may take Q K^T now
                // stash the row (group 1 uses the upper half of the P region; group 0 must not touch the Rcat operand in the
                // first 16 KB before the G MMAs of BOTH groups have completed)
                ptx::mbar_wait(r_done, par);
                float* gs = reinterpret_cast<float*>(smem_gen + (sP - smem_base) + grp * 2 * WT_P_BYTES) + r * 64;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    *reinterpret_cast<uint4*>(gs + i) = make_uint4(g0[i], g0[i + 1], g0[i + 2], g0[i + 3]);
                    *reinterpret_cast<uint4*>(gs + 32 + i) = make_uint4(g1[i], g1[i + 1], g1[i + 2], g1[i + 3]);
                }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < WT_G; ++k) {
                    bh[k] = gs[qh + WT_G - 1 - k] * FT_L2E;
                    bw[k] = gs[32 + qw + WT_G - 1 - k] * FT_L2E;
                }
                ptx::named_bar_sync(1 + grp, 128);  // every row of the group is read before the first P row is written
            }
            ptx::mbar_wait(s_full(grp), par);
            ptx::tc_fence_after();
            // ---- exact row maximum over the 196 biased scores
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (c < 3) {
                    uint32_t a[32], b2[32];
                    ptx::tmem_ld32(tS(grp) + lane_off + (uint32_t)(c * 64), a);
                    ptx::tmem_ld32(tS(grp) + lane_off + (uint32_t)(c * 64 + 32), b2);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int k0 = c * 64 + j, k1 = k0 + 32;
                        mx = fmaxf(mx, fmaf(__uint_as_float(a[j]), sl2, bh[k0 / WT_G] + bw[k0 % WT_G]));
                        mx = fmaxf(mx, fmaf(__uint_as_float(b2[j]), sl2, bh[k1 / WT_G] + bw[k1 % WT_G]));
                    }
                } else {
                    uint32_t d[16];
                    ptx::tmem_ld16(tS(grp) + lane_off + 192u, d);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int k0 = 192 + j;
                        mx = fmaxf(mx, fmaf(__uint_as_float(d[j]), sl2, bh[k0 / WT_G] + bw[k0 % WT_G]));
                    }
                }
            }
            float l_run = 0.f;
            // ---- P chunks in the order 3, 2, 1, 0
#pragma unroll
            for (int c = 3; c >= 0; --c) {
                const int buf = c & 1;
                uint32_t pk[32];
                if (c == 3) {
                    uint32_t d[16];
                    ptx::tmem_ld16(tS(grp) + lane_off + 192u, d);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int k0 = 192 + 2 * j;
                        float p0 = 0.f, p1 = 0.f;
                        if (k0 < WT_S) p0 = ptx::ex2(fmaf(__uint_as_float(d[2 * j]), sl2, bh[k0 / WT_G] + bw[k0 % WT_G]) - mx);
                        if (k0 + 1 < WT_S) p1 = ptx::ex2(fmaf(__uint_as_float(d[2 * j + 1]), sl2, bh[(k0 + 1) / WT_G] + bw[(k0 + 1) % WT_G]) - mx);
                        l_run += p0 + p1;
                        pk[j] = pack_h2(p0, p1);
                    }
                } else {
                    uint32_t a[32], b2[32];
                    ptx::tmem_ld32(tS(grp) + lane_off + (uint32_t)(c * 64), a);
                    ptx::tmem_ld32(tS(grp) + lane_off + (uint32_t)(c * 64 + 32), b2);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int k0 = c * 64 + 2 * j, k2 = k0 + 32;
                        const float p0 = ptx::ex2(fmaf(__uint_as_float(a[2 * j]), sl2, bh[k0 / WT_G] + bw[k0 % WT_G]) - mx);
                        const float p1 = ptx::ex2(fmaf(__uint_as_float(a[2 * j + 1]), sl2, bh[(k0 + 1) / WT_G] + bw[(k0 + 1) % WT_G]) - mx);
                        const float p2 = ptx::ex2(fmaf(__uint_as_float(b2[2 * j]), sl2, bh[k2 / WT_G] + bw[k2 % WT_G]) - mx);
                        const float p3 = ptx::ex2(fmaf(__uint_as_float(b2[2 * j + 1]), sl2, bh[(k2 + 1) / WT_G] + bw[(k2 + 1) % WT_G]) - mx);
                        l_run += (p0 + p1) + (p2 + p3);
                        pk[j] = pack_h2(p0, p1);
                        pk[16 + j] = pack_h2(p2, p3);
                    }
                }
                ptx::tc_fence_before();
                // the buffer was last read by the P V of: chunk c + 2 of this pair (c = 1, 0) or chunk c - 2 of the previous pair
                ptx::mbar_wait(p_free(grp, buf), c >= 2 ? 1u : 0u);  // two completions per pair: parity 1 = last P V of the previous pair
                {
                    const uint32_t base = sP + (grp * 2 + buf) * WT_P_BYTES + p_row;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if (c < 3 || q < 2) {
                            const uint32_t addr = base + (((uint32_t)q ^ sw) << 4);
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * q]), "r"(pk[4 * q + 1]),
                                         "r"(pk[4 * q + 2]), "r"(pk[4 * q + 3]) : "memory");
                        }
                    }
                }
                ptx::fence_proxy_async();
                ptx::mbar_arrive(p_full(grp, c));
            }
            {
                ptx::mbar_wait(o_full(grp), par);
                ptx::tc_fence_after();
                const float inv = 1.0f / l_run;
                __half* dst = out + ((long long)row0 + qi) * D + head * FT_HD;
                auto f = [&](uint32_t u) { return __uint_as_float(u) * inv; };
#pragma unroll 1
                for (int c0 = 0; c0 < FT_HD; c0 += 16) {
                    uint32_t d[16];
                    ptx::tmem_ld16(tO(grp) + lane_off + (uint32_t)c0, d);
                    ptx::tmem_ld_wait();
                    if (qi < WT_S) {
#pragma unroll
                        for (int i = 0; i < 16; i += 8)
                            *reinterpret_cast<uint4*>(dst + c0 + i) = make_uint4(pack_h2(f(d[i]), f(d[i + 1])), pack_h2(f(d[i + 2]), f(d[i + 3])),
                                                                                 pack_h2(f(d[i + 4]), f(d[i + 5])), pack_h2(f(d[i + 6]), f(d[i + 7])));
                    }
                }
                ptx::tc_fence_before();
                ptx::mbar_arrive(o_read(grp));
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------ windowed attention, third design
// EXPERIMENTAL -- written after the GPU budget of round 1 was spent: it compiles for sm_100a but has NOT been run yet (variant
// 3 of cvb_set_window_tc_variant, off by default; tests/test_gpu_forward.py covers it only with CVB_EXPERIMENTAL=1).
//
// A window item is treated like ONE key tile of flash_tc_kernel. The scores leave the tensor core fully biased:
//     S = [Q | Gsel] [K | Sel]^T    (5 + 2 k-steps of one UMMA 128 x 208 x 16 per query group)
// with Gsel = [rel_h(q, kh) | rel_w(q, kw)] / scale from the pre-pass window_qg_kernel (attention.cu; its 64-column row also
// carries Q[.][64..79], so one shared-memory box serves the fifth Q k-step and both bias k-steps) and Sel the constant 0/1
// selection matrix (Sel[k][kh(k)] = Sel[k][14 + kw(k)] = 1), resident in shared memory. The softmax threads (one query row each)
// make two passes over their S row in tensor memory -- running maximum, then P = 2^((s - m) scale log2 e) as packed fp16 pairs
// (ex2.approx.f16x2) stored back IN PLACE in ascending key order: the packed chunk c (columns 16c..16c+15) lies below the
// score chunk c (columns 32c..32c+31) that produced it -- so ~64 registers are live and no bias arithmetic is left in the
// loop. P V is a TS-mode UMMA (A = P from tensor memory) over V^T with a ones row (row sums = output column 80); O goes to
// columns 128..223 of the group's 256 (the score columns there are dead by then, P sits in 0..103). Shared memory is single-buffered (190 KB): Q /
// QG / K are released as soon as both S MMAs have completed, V^T after the second P V, and the producer refills them for the
// next item while this item's softmax and P V run.
constexpr int W3_VR = 96;                                        // V^T rows per head: 80 + ones + 15 zero rows
constexpr uint32_t W3_QB = FT_BQ * 128;                          // one 128-row box (16 KB)
constexpr uint32_t W3_KB = WT_VLD * 128;                         // one 208-row box (26 KB)
constexpr uint32_t W3_VB = W3_VR * 128;                          // one V^T box: 96 rows x 64 keys
constexpr uint32_t W3_SMEM = 1024 + 4 * W3_QB + 3 * W3_KB + 4 * W3_VB + 512;
constexpr uint32_t W3_O_COL = 128;                              // O columns of a group: 128..223 (32-column aligned, as in the second design)

// v rows of window `item` [196, 3*D] -> vt [(head*96 + d)][item*208 + key]; row 80 = ones (keys < 196), rows 81..95 and keys
// 196..207 zero
__global__ void __launch_bounds__(256)
v_transpose_win96_kernel(const __half* __restrict__ qkv, int heads, int n_items, __half* __restrict__ vt) {
    __shared__ __half tile[WT_VLD][FT_HD + 2];
    const int item = blockIdx.x, head = blockIdx.y;
    const int D = heads * FT_HD;
    const __half* src = qkv + (long long)item * WT_S * 3 * D + 2 * D + head * FT_HD;
    for (int i = threadIdx.x; i < WT_VLD * (FT_HD / 8); i += 256) {
        const int r = i / (FT_HD / 8), c = i - r * (FT_HD / 8);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r < WT_S) v = *reinterpret_cast<const uint4*>(src + (long long)r * 3 * D + c * 8);
        const __half* h = reinterpret_cast<const __half*>(&v);
#pragma unroll
        for (int j = 0; j < 8; ++j) tile[r][c * 8 + j] = h[j];
    }
    __syncthreads();
    const long long ld = (long long)n_items * WT_VLD;
    __half* dst = vt + (long long)head * W3_VR * ld + (long long)item * WT_VLD;
    for (int i = threadIdx.x; i < W3_VR * (WT_VLD / 2); i += 256) {
        const int d = i / (WT_VLD / 2), kp = i - d * (WT_VLD / 2);
        __half2 v = __floats2half2_rn(0.f, 0.f);
        if (d < FT_HD) v = __halves2half2(tile[2 * kp][d], tile[2 * kp + 1][d]);
        else if (d == FT_HD && 2 * kp < WT_S) v = __floats2half2_rn(1.f, 1.f);   // 196 is even: a pair is valid or padding as a whole
        *reinterpret_cast<__half2*>(dst + (long long)d * ld + 2 * kp) = v;
    }
}

// sel fp16 [208][64]: row k (a key of the 14 x 14 window) has ones at columns kh(k) and 14 + kw(k)
__global__ void window_sel_kernel(__half* __restrict__ sel) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= WT_VLD * 64) return;
    const int k = i >> 6, c = i & 63;
    const int kh = k / WT_G, kw = k - kh * WT_G;
    sel[i] = __float2half((k < WT_S && (c == kh || c == WT_G + kw)) ? 1.0f : 0.0f);
}

__global__ void __launch_bounds__(WT_THREADS, 1)
window_tc3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQG,
                  const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmK16,
                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmSel, int heads, int n_items,
                  float scale, __half* __restrict__ out) {
    extern __shared__ uint8_t ft_smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(ft_smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = ft_smem_raw + (smem_base - ptx::smem_u32(ft_smem_raw));
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const int D = heads * FT_HD;
    const int n_work = n_items * heads;

    const uint32_t sQ0 = smem_base;               // [group] Q columns 0..63
    const uint32_t sQG = sQ0 + 2 * W3_QB;          // [group] Q columns 64..79 | Gsel | 0
    const uint32_t sK0 = sQG + 2 * W3_QB;          // K columns 0..63, 208 rows
    const uint32_t sKt = sK0 + W3_KB;              // K columns 16..79 (the last k-step reads its columns 48..63)
    const uint32_t sSel = sKt + W3_KB;
    const uint32_t sV = sSel + W3_KB;              // 4 boxes of 64 keys
    const uint32_t bar = sV + 4 * W3_VB;
    const uint32_t sel_full = bar, qk_full = bar + 8, qk_free = bar + 16, v_full = bar + 24, v_free = bar + 32;
    auto s_full = [&](int g) { return bar + 8u * (5 + g); };
    auto p_full = [&](int g) { return bar + 8u * (7 + g); };
    auto o_full = [&](int g) { return bar + 8u * (9 + g); };
    auto o_free = [&](int g) { return bar + 8u * (11 + g); };
    const uint32_t tmem_slot = bar + 8u * 13;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmQ); ptx::prefetch_tmap(&tmQG); ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmK16);
        ptx::prefetch_tmap(&tmV); ptx::prefetch_tmap(&tmSel);
    }
    if (warp == 1 && lane == 0) {
        ptx::mbar_init(sel_full, 1); ptx::mbar_init(qk_full, 1); ptx::mbar_init(qk_free, 1); ptx::mbar_init(v_full, 1);
        ptx::mbar_init(v_free, 1);
        for (int g = 0; g < 2; ++g) {
            ptx::mbar_init(s_full(g), 1); ptx::mbar_init(p_full(g), 128); ptx::mbar_init(o_full(g), 1); ptx::mbar_init(o_free(g), 128);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_slot, 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
    auto tS = [&](int g) { return tmem_base + (uint32_t)(g * 256); };
    auto tO = [&](int g) { return tmem_base + (uint32_t)(g * 256) + W3_O_COL; };

    if (warp == 0) {
        // ===================================================== TMA producer
        if (ptx::elect_one()) {
            ptx::mbar_expect_tx(sel_full, W3_KB);
            ptx::tma_load_2d(sSel, &tmSel, sel_full, 0, 0);
        }
        int it = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
            const int item = w / heads, head = w - item * heads;
            const int row0 = item * WT_S;
            const uint32_t par = (uint32_t)(it & 1);
            ptx::mbar_wait(qk_free, par ^ 1u);     // both S MMAs of the previous item have completed
            if (ptx::elect_one()) {
                ptx::mbar_expect_tx(qk_full, 4 * W3_QB + 2 * W3_KB);
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    ptx::tma_load_2d(sQ0 + g * W3_QB, &tmQ, qk_full, head * FT_HD, row0 + g * FT_BQ);
                    ptx::tma_load_2d(sQG + g * W3_QB, &tmQG, qk_full, head * 64, row0 + g * FT_BQ);
                }
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    ptx::tma_load_2d(sK0 + t * 8192, &tmK, qk_full, D + head * FT_HD, row0 + t * 64);
                    ptx::tma_load_2d(sKt + t * 8192, &tmK, qk_full, D + head * FT_HD + 16, row0 + t * 64);
                }
                ptx::tma_load_2d(sK0 + 3 * 8192, &tmK16, qk_full, D + head * FT_HD, row0 + 192);
                ptx::tma_load_2d(sKt + 3 * 8192, &tmK16, qk_full, D + head * FT_HD + 16, row0 + 192);
            }
            ptx::mbar_wait(v_free, par ^ 1u);      // both P V of the previous item have completed
            if (ptx::elect_one()) {
                ptx::mbar_expect_tx(v_full, 4 * W3_VB);
#pragma unroll
                for (int t = 0; t < 4; ++t) ptx::tma_load_2d(sV + t * W3_VB, &tmV, v_full, item * WT_VLD + t * FT_BK, head * W3_VR);
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        auto idesc = [](int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(FT_BQ >> 4) << 24); };
        const uint64_t desc_hi = (2ull << 61) | (1ull << 46) | ((uint64_t)(1024 >> 4) << 32);  // SWIZZLE_128B, SBO 1024
        auto desc = [&](uint32_t addr) { return desc_hi | (uint64_t)((addr >> 4) & 0x3FFF); };
        ptx::mbar_wait(sel_full, 0);
        int it = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
            const uint32_t par = (uint32_t)(it & 1);
            ptx::mbar_wait(qk_full, par);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                ptx::mbar_wait(o_free(g), par ^ 1u);   // the previous item's O (same TMEM columns) has been read out
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    const uint64_t a0 = desc(sQ0 + g * W3_QB), aq = desc(sQG + g * W3_QB);
                    const uint64_t b0 = desc(sK0), bt = desc(sKt), bs = desc(sSel);
#pragma unroll
                    for (int k = 0; k < 4; ++k) ptx::umma_f16(tS(g), a0 + 2u * k, b0 + 2u * k, idesc(WT_VLD), k != 0 ? 1u : 0u);
                    ptx::umma_f16(tS(g), aq, bt + 6u, idesc(WT_VLD), 1u);           // Q / K columns 64..79
                    ptx::umma_f16(tS(g), aq + 2u, bs, idesc(WT_VLD), 1u);           // + Gsel Sel^T (two k-steps of 16 bias columns)
                    ptx::umma_f16(tS(g), aq + 4u, bs + 2u, idesc(WT_VLD), 1u);
                    ptx::umma_commit(s_full(g));
                    if (g == 1) ptx::umma_commit(qk_free);
                }
                __syncwarp();
            }
            ptx::mbar_wait(v_full, par);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                ptx::mbar_wait(p_full(g), par);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
#pragma unroll
                    for (int k = 0; k < WT_VLD / 16; ++k)   // 13 k-steps of 16 keys; P chunk k = 8 packed TMEM columns
                        ptx::umma_f16_ts(tO(g), tS(g) + 8u * k, desc(sV + (uint32_t)(k >> 2) * W3_VB) + 2u * (k & 3), idesc(W3_VR),
                                         k != 0 ? 1u : 0u);
                    ptx::umma_commit(o_full(g));
                    if (g == 1) ptx::umma_commit(v_free);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===================================================== softmax / output: one query row per thread
        const int grp = (warp - 4) >> 2, quad = warp & 3, r = quad * 32 + lane;
        const int qi = grp * FT_BQ + r;
        const bool row_ok = qi < WT_S;
        const bool warp_ok = grp * FT_BQ + quad * 32 < WT_S;   // warps whose 32 rows are all padding only keep the barrier counts
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        const float sl2 = scale * FT_L2E;
        auto fl = [](uint32_t u) { return __uint_as_float(u); };
        int it = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
            const int item = w / heads, head = w - item * heads;
            const uint32_t par = (uint32_t)(it & 1);
            // (every thread waits and arrives, also in the all-padding warps: an arrival that ran ahead of its phase would
            //  be counted towards the previous one)
            ptx::mbar_wait(s_full(grp), par);
            if (warp_ok) {
                ptx::tc_fence_after();
                const uint32_t ts = tS(grp) + lane_off;
                // pass 1: exact row maximum over the 196 keys (raw accumulator units)
                float mx = -INFINITY;
#pragma unroll 1
                for (int c = 0; c < 6; ++c) {
                    uint32_t v[32];
                    ptx::tmem_ld32(ts + 32u * c, v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j += 2) mx = fmaxf(mx, fmaxf(fl(v[j]), fl(v[j + 1])));
                }
                {
                    uint32_t v[16];
                    ptx::tmem_ld16(ts + 192u, v);
                    ptx::tmem_ld_wait();
                    mx = fmaxf(mx, fmaxf(fmaxf(fl(v[0]), fl(v[1])), fmaxf(fl(v[2]), fl(v[3]))));   // keys 192..195; 196..207 are padding
                }
                const float mneg = -mx * sl2;
                // pass 2: P in place, ascending key order
#pragma unroll 1
                for (int c = 0; c < 6; ++c) {
                    uint32_t v[32], pk[16];
                    ptx::tmem_ld32(ts + 32u * c, v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        pk[j] = ptx::ex2_f16x2(pack_h2(fmaf(fl(v[2 * j]), sl2, mneg), fmaf(fl(v[2 * j + 1]), sl2, mneg)));
                    ptx::tmem_st16(ts + 16u * c, pk);
                }
                {
                    uint32_t v[16], pk[16];
                    ptx::tmem_ld16(ts + 192u, v);
                    ptx::tmem_ld_wait();
                    pk[0] = ptx::ex2_f16x2(pack_h2(fmaf(fl(v[0]), sl2, mneg), fmaf(fl(v[1]), sl2, mneg)));
                    pk[1] = ptx::ex2_f16x2(pack_h2(fmaf(fl(v[2]), sl2, mneg), fmaf(fl(v[3]), sl2, mneg)));
#pragma unroll
                    for (int j = 2; j < 16; ++j) pk[j] = 0u;       // keys 196..207: weight 0 (columns 104..111 are unused padding)
                    ptx::tmem_st16(ts + 96u, pk);
                }
                ptx::tmem_st_wait();
                ptx::tc_fence_before();
            }
            ptx::mbar_arrive(p_full(grp));
            ptx::mbar_wait(o_full(grp), par);
            if (warp_ok) {
                ptx::tc_fence_after();
                const uint32_t to = tO(grp) + lane_off;
                float inv;
                {
                    uint32_t d[16];
                    ptx::tmem_ld16(to + (uint32_t)FT_HD, d);   // output column 80 = sum of the row's (fp16-rounded) weights
                    ptx::tmem_ld_wait();
                    inv = 1.0f / fl(d[0]);
                }
                __half* dst = out + ((long long)item * WT_S + qi) * D + head * FT_HD;
                auto f = [&](uint32_t u) { return __uint_as_float(u) * inv; };
#pragma unroll 1
                for (int c0 = 0; c0 < FT_HD; c0 += 16) {
                    uint32_t d[16];
                    ptx::tmem_ld16(to + (uint32_t)c0, d);
                    ptx::tmem_ld_wait();
                    if (row_ok) {
#pragma unroll
                        for (int i = 0; i < 16; i += 8)
                            *reinterpret_cast<uint4*>(dst + c0 + i) = make_uint4(pack_h2(f(d[i]), f(d[i + 1])), pack_h2(f(d[i + 2]), f(d[i + 3])),
                                                                                 pack_h2(f(d[i + 4]), f(d[i + 5])), pack_h2(f(d[i + 6]), f(d[i + 7])));
                    }
                }
                ptx::tc_fence_before();
            }
            ptx::mbar_arrive(o_free(grp));
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

PFN_cuTensorMapEncodeTiled_v12000 ft_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

int ft_tmap_2d(CUtensorMap* tm, const void* base, uint64_t cols, uint64_t rows, uint64_t row_stride_bytes, uint32_t box_cols,
               uint32_t box_rows) {
    auto fn = ft_encode_fn();
    CVB_CHECK(fn != nullptr, CVB_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    uint64_t dims[2] = {cols, rows};
    uint64_t str[1] = {row_stride_bytes};
    uint32_t box[2] = {box_cols, box_rows};
    uint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, str, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CVB_CHECK(r == CUDA_SUCCESS, CVB_ECUDA, "flash_tc: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return CVB_OK;
}

}  // namespace

bool op_attention_tc_supported(int S, int hd, const __half* Rh, int gh, int gw) {
    return hd == FT_HD && Rh != nullptr && gw == FT_BK && gh <= 64 && gh * gw == S && S % (FT_NG * FT_BQ) == 0;
}

size_t op_attention_tc_workspace_bytes(int Gb, int S, int heads) {
    const size_t vt = align_up((size_t)Gb * heads * FT_VR * S * 2, 1024);
    const size_t bias = align_up((size_t)Gb * heads * S * 64 * 2, 1024);
    return vt + 2 * bias + 1024;
}

int op_attention_tc(const __half* qkv, int Gb, int S, int heads, int hd, float scale, const __half* Rh, const __half* Rw, int gh,
                    int gw, __half* out, void* workspace, size_t ws_bytes, cudaStream_t stream) {
    CVB_CHECK(qkv && out && workspace, CVB_EARG, "attention_tc: null operand");
    CVB_CHECK(op_attention_tc_supported(S, hd, Rh, gh, gw) && Rw != nullptr, CVB_ESHAPE,
              "attention_tc: needs head dim 80, a 64-wide token grid and S %% 256 == 0 (S=%d hd=%d grid %dx%d)", S, hd, gh, gw);
    CVB_CHECK(ws_bytes >= op_attention_tc_workspace_bytes(Gb, S, heads), CVB_EWORKSPACE, "attention_tc: workspace too small");
    CVB_CHECK(((uintptr_t)workspace & 1023) == 0, CVB_EARG, "attention_tc: workspace must be 1024-byte aligned");
    const int D = heads * hd;
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    __half* vt = reinterpret_cast<__half*>(ws);
    const size_t vt_b = align_up((size_t)Gb * heads * FT_VR * S * 2, 1024);
    const size_t bias_b = align_up((size_t)Gb * heads * S * 64 * 2, 1024);
    __half* bias_h = reinterpret_cast<__half*>(ws + vt_b);
    __half* bias_w = reinterpret_cast<__half*>(ws + vt_b + bias_b);

    static unsigned long long configured = 0;  // one bit per device: function attributes are per device
    const int cfg_dev = cvb_current_device();
    if (!((configured >> cfg_dev) & 1ull)) {
        CVB_CUDA(cudaFuncSetAttribute(flash_tc_kernel<FT_NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FtCfg<FT_NG>::SMEM));
        configured |= 1ull << cfg_dev;
    }
    const dim3 grid64(S / 64, Gb * heads);
    v_transpose_kernel<<<grid64, 256, 0, stream>>>(qkv, S, heads, vt);
    CVB_TRY(op_relpos_tables(qkv, Gb, S, heads, hd, Rh, Rw, gh, gw, bias_h, bias_w, stream));
    CUtensorMap tq, tk, tv;
    CVB_TRY(ft_tmap_2d(&tq, qkv, (uint64_t)3 * D, (uint64_t)Gb * S, (uint64_t)3 * D * 2, 64, FT_BQ));
    CVB_TRY(ft_tmap_2d(&tk, qkv, (uint64_t)3 * D, (uint64_t)Gb * S, (uint64_t)3 * D * 2, 64, FT_BK));
    CVB_TRY(ft_tmap_2d(&tv, vt, (uint64_t)S, (uint64_t)Gb * heads * FT_VR, (uint64_t)S * 2, 64, FT_VR));
    flash_tc_kernel<FT_NG><<<dim3(S / (FT_NG * FT_BQ), Gb * heads), FtCfg<FT_NG>::THREADS, FtCfg<FT_NG>::SMEM, stream>>>(tq, tk, tv, bias_h, bias_w, S, heads,
                                                                                                               scale, out);
    cvb_note_launches(2);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}

// ------------------------------------------------------------------------------------------ windows: host side
bool op_window_attention_tc_supported(int S, int hd, int gh, int gw) { return hd == FT_HD && S == WT_S && gh == WT_G && gw == WT_G; }

static size_t w3_vt_bytes(int n_items, int heads) { return align_up((size_t)heads * W3_VR * n_items * WT_VLD * 2, 1024); }
static size_t w3_qg_bytes(int n_items, int heads) { return align_up((size_t)n_items * WT_S * heads * 64 * 2, 1024); }

size_t op_window_attention_tc_workspace_bytes(int n_items, int heads) {
    // variants 1 / 2: V^T (80 rows per head); variant 3: V^T (96 rows per head) + QG + Sel
    const size_t v12 = align_up((size_t)heads * FT_HD * n_items * WT_VLD * 2, 1024) + 1024;
    const size_t v3 = w3_vt_bytes(n_items, heads) + w3_qg_bytes(n_items, heads) + align_up((size_t)W3_KB, 1024) + 1024;
    return v12 > v3 ? v12 : v3;
}

// 1: four-key-tile loop (window_tc_kernel, 128 us per SAM-H block at B = 4), 2: single-shot N = 208 (window_tc2_kernel, 152 us:
// one thread per 196-score row is ~1,800 serial instructions per pair, and only 8 softmax warps fit beside the TMEM budget)
static int g_window_tc_variant = 1;
// 3: scores fully biased by the tensor core, softmax in place in tensor memory (window_tc3_kernel) -- EXPERIMENTAL, never run yet
extern "C" __attribute__((visibility("default"))) void cvb_set_window_tc_variant(int v) { g_window_tc_variant = (v == 2 || v == 3) ? v : 1; }

int op_window_attention_tc(const __half* qkv, int n_items, int heads, int hd, float scale, const __half* relcat, __half* out,
                           void* workspace, size_t ws_bytes, cudaStream_t stream) {
    CVB_CHECK(qkv && out && relcat && workspace, CVB_EARG, "window_attention_tc: null operand");
    CVB_CHECK(hd == FT_HD && n_items > 0 && heads > 0, CVB_ESHAPE, "window_attention_tc: needs head dim 80");
    CVB_CHECK(ws_bytes >= op_window_attention_tc_workspace_bytes(n_items, heads), CVB_EWORKSPACE, "window_attention_tc: workspace too small");
    CVB_CHECK(((uintptr_t)workspace & 1023) == 0, CVB_EARG, "window_attention_tc: workspace must be 1024-byte aligned");
    const int D = heads * hd;
    __half* vt = reinterpret_cast<__half*>(workspace);
    static unsigned long long configured = 0;  // one bit per device: function attributes are per device
    const int cfg_dev = cvb_current_device();
    if (!((configured >> cfg_dev) & 1ull)) {
        CVB_CUDA(cudaFuncSetAttribute(window_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WT_SMEM));
        CVB_CUDA(cudaFuncSetAttribute(window_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W2_SMEM));
        CVB_CUDA(cudaFuncSetAttribute(window_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W3_SMEM));
        configured |= 1ull << cfg_dev;
    }
    if (g_window_tc_variant == 3) {
        __half* vt96 = reinterpret_cast<__half*>(workspace);
        __half* qg = reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(workspace) + w3_vt_bytes(n_items, heads));
        __half* sel = reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(qg) + w3_qg_bytes(n_items, heads));
        v_transpose_win96_kernel<<<dim3(n_items, heads), 256, 0, stream>>>(qkv, heads, n_items, vt96);
        window_sel_kernel<<<(WT_VLD * 64 + 255) / 256, 256, 0, stream>>>(sel);
        cvb_note_launches(2);
        CVB_TRY(op_window_qg(qkv, n_items, WT_S, heads, hd, scale, relcat, relcat + 32 * FT_HD, WT_G, WT_G, qg, stream));
        CUtensorMap tq, tqg, tk, tk16, tv, ts;
        const uint64_t rows = (uint64_t)n_items * WT_S;
        CVB_TRY(ft_tmap_2d(&tq, qkv, (uint64_t)3 * D, rows, (uint64_t)3 * D * 2, 64, FT_BQ));
        CVB_TRY(ft_tmap_2d(&tqg, qg, (uint64_t)heads * 64, rows, (uint64_t)heads * 64 * 2, 64, FT_BQ));
        CVB_TRY(ft_tmap_2d(&tk, qkv, (uint64_t)3 * D, rows, (uint64_t)3 * D * 2, 64, FT_BK));
        CVB_TRY(ft_tmap_2d(&tk16, qkv, (uint64_t)3 * D, rows, (uint64_t)3 * D * 2, 64, 16));
        CVB_TRY(ft_tmap_2d(&tv, vt96, (uint64_t)n_items * WT_VLD, (uint64_t)heads * W3_VR, (uint64_t)n_items * WT_VLD * 2, 64, W3_VR));
        CVB_TRY(ft_tmap_2d(&ts, sel, 64, WT_VLD, 64 * 2, 64, WT_VLD));
        const int n_work3 = n_items * heads;
        const int grid3 = n_work3 < cvb_num_sms() ? n_work3 : cvb_num_sms();
        window_tc3_kernel<<<grid3, WT_THREADS, W3_SMEM, stream>>>(tq, tqg, tk, tk16, tv, ts, heads, n_items, scale, out);
        cvb_note_launches(1);
        CVB_CUDA(cudaGetLastError());
        return CVB_OK;
    }
    v_transpose_win_kernel<<<dim3(n_items, heads), 256, 0, stream>>>(qkv, heads, n_items, vt);
    CUtensorMap tq, tk, tv, tr;
    const uint64_t rows = (uint64_t)n_items * WT_S;
    CVB_TRY(ft_tmap_2d(&tq, qkv, (uint64_t)3 * D, rows, (uint64_t)3 * D * 2, 64, FT_BQ));
    CVB_TRY(ft_tmap_2d(&tk, qkv, (uint64_t)3 * D, rows, (uint64_t)3 * D * 2, 64, FT_BK));
    CVB_TRY(ft_tmap_2d(&tv, vt, (uint64_t)n_items * WT_VLD, (uint64_t)heads * FT_HD, (uint64_t)n_items * WT_VLD * 2, 64, FT_HD));
    CVB_TRY(ft_tmap_2d(&tr, relcat, (uint64_t)FT_HD, 64, (uint64_t)FT_HD * 2, 64, 64));
    const int n_work = n_items * heads;
    const int grid = n_work < cvb_num_sms() ? n_work : cvb_num_sms();
    if (g_window_tc_variant == 2) {
        CUtensorMap tk16;
        CVB_TRY(ft_tmap_2d(&tk16, qkv, (uint64_t)3 * D, rows, (uint64_t)3 * D * 2, 64, 16));
        window_tc2_kernel<<<grid, WT_THREADS, W2_SMEM, stream>>>(tq, tk, tk16, tv, tr, heads, n_items, scale, out);
    } else {
        window_tc_kernel<<<grid, WT_THREADS, WT_SMEM, stream>>>(tq, tk, tv, tr, heads, n_items, scale, out);
    }
    cvb_note_launches(2);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}
