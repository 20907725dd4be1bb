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
#include "ops.h"

namespace {

constexpr int FT_BQ = 128, FT_BK = 64, FT_HD = 80, FT_STAGES = 3;
constexpr uint32_t FT_K_BYTES = 2 * FT_BK * 128;          // two boxes of 64 rows x 128 B
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
    CVB_TRY(cvb_tmap_2d_f16(&tq, qkv, (uint64_t)3 * D, (uint64_t)Gb * S, (uint64_t)3 * D * 2, 64, FT_BQ));
    CVB_TRY(cvb_tmap_2d_f16(&tk, qkv, (uint64_t)3 * D, (uint64_t)Gb * S, (uint64_t)3 * D * 2, 64, FT_BK));
    CVB_TRY(cvb_tmap_2d_f16(&tv, vt, (uint64_t)S, (uint64_t)Gb * heads * FT_VR, (uint64_t)S * 2, 64, FT_VR));
    flash_tc_kernel<FT_NG><<<dim3(S / (FT_NG * FT_BQ), Gb * heads), FtCfg<FT_NG>::THREADS, FtCfg<FT_NG>::SMEM, stream>>>(tq, tk, tv, bias_h, bias_w, S, heads,
                                                                                                               scale, out);
    cvb_note_launches(2);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}
