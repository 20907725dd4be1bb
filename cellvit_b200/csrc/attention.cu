// attention.cu -- multi-head self-attention of the ViT encoders (windowed and global, SAM and ViT-S).
//
// flash_kernel<HD, BIAS>: S = scale * q k^T (+ rel_h[q, k / gw] + rel_w[q, k % gw]); online softmax; O = P v.
// Scores never leave the SM (the reference materialises a 4.3 GB fp32 score tensor at B=4,
// image_encoder.py:244-251). Tensor-core path: mma.sync m16n8k16 fp16 -> fp32 with ldmatrix-fed fragments and
// cp.async double-buffered K/V tiles.
//
// Decomposed relative position (image_encoder.py:354-392): rel_h[q, kh] = q . Rh[qh - kh + gh - 1] with the
// UNSCALED q. The kernel computes G = Q Rh^T for all 2*gh-1 table rows with the same MMA path as Q K^T (the table
// is just another K-major operand tile) and scatters G[q, j] to rel_h[q, kh = qh + gh - 1 - j] in shared memory;
// same for rel_w. No separate rel-pos kernel, no HBM round trip of the bias tables.
//
// The SAM global-attention shape (64 x 64 tokens, head dim 80) runs on tcgen05 instead (flash_tc.cu); this file keeps the
// mma.sync kernels for the windows, ViT-S and the smaller token grids, and the rel-pos table kernel flash_tc.cu uses.
#include "ops.h"

namespace {

constexpr int FA_BQ = 64, FA_BK = 64, FA_THREADS = 128;
constexpr int REL_LD = 66;  // row stride (halves) of the bias tables: 33 words, rows land in different banks

template <int HD, bool BIAS>
struct FaSmem {
    static constexpr int LD = HD + 8;  // halves per smem row: (HD+8)*2 B is an odd multiple of 16 B -> conflict-free ldmatrix
    __half q[FA_BQ * LD];
    __half k[2][FA_BK * LD];
    __half v[2][FA_BK * LD];
    __half rel_h[BIAS ? FA_BQ * REL_LD : 2];  // fp16 keeps the CTA at 73 KB -> 3 CTAs (12 warps) per SM
    __half rel_w[BIAS ? FA_BQ * REL_LD : 2];
};

// 64-row x HD tile loader. Every thread owns the same NJ 16-byte chunks of every tile, so the (row, chunk) split and the
// source / destination offsets are computed once (the naive per-tile div/mod + 64-bit address math was 43 % of the
// global-attention kernel's instructions).
template <int HD>
struct TileLoader {
    static constexpr int LD = HD + 8;
    static constexpr int CH = HD / 8;                                      // 16-byte chunks per row
    static constexpr int NJ = (FA_BK * CH + FA_THREADS - 1) / FA_THREADS;  // chunks per thread (5 for HD=80, 4 for 64)
    int row[NJ];
    int src_off[NJ];       // in halves, relative to the tile's first row
    uint32_t dst_off[NJ];  // in bytes
    __device__ __forceinline__ void init(int tid, int row_stride) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int i = tid + j * FA_THREADS;
            const int r = i / CH, c = i - r * CH;
            row[j] = i < FA_BK * CH ? r : (1 << 30);
            src_off[j] = r * row_stride + c * 8;
            dst_off[j] = (uint32_t)(r * LD + c * 8) * 2u;
        }
    }
    // rows >= n_rows (or chunks past the tile) are zero-filled
    __device__ __forceinline__ void load(uint32_t dst_smem, const __half* tile_src, int row0, int n_rows) const {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            if ((FA_BK * CH) % FA_THREADS != 0 && row[j] >= FA_BK) continue;
            const bool ok = row0 + row[j] < n_rows;
            ptx::cp_async16(dst_smem + dst_off[j], tile_src + (ok ? src_off[j] : 0), ok);
        }
    }
};

// s_acc[16 x 64 per warp] = Q_frag (16 x HD) * tile^T, tile = [64 rows][HD] K-major in shared memory
template <int HD>
__device__ __forceinline__ void fa_qk(const __half* tile, const uint32_t (&q_frag)[HD / 16][4], float (&s_acc)[8][4], int lane,
                                      int n_pairs) {
    constexpr int LD = HD + 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s_acc[i][0] = s_acc[i][1] = s_acc[i][2] = s_acc[i][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {  // pairs of 8-row n-tiles
            if (np < n_pairs) {
                uint32_t b0, b1, b2, b3;
                const int row = np * 16 + (lane & 7) + ((lane >> 4) << 3);
                const int col = ks * 16 + ((lane >> 3) & 1) * 8;
                ptx::ldmatrix_x4(ptx::smem_u32(tile + row * LD + col), b0, b1, b2, b3);
                ptx::mma_16816(s_acc[2 * np], q_frag[ks], b0, b1);
                ptx::mma_16816(s_acc[2 * np + 1], q_frag[ks], b2, b3);
            }
        }
    }
}

template <int HD, bool BIAS>
__global__ void __launch_bounds__(FA_THREADS, 3)
flash_kernel(const __half* __restrict__ qkv, int S, int heads, float scale, const __half* __restrict__ Rh,
             const __half* __restrict__ Rw, int gh, int gw, __half* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t fa_smem_raw[];
    using Smem = FaSmem<HD, BIAS>;
    Smem& sm = *reinterpret_cast<Smem*>(fa_smem_raw);
    constexpr int LD = Smem::LD;
    constexpr int KSTEPS = HD / 16;   // k-steps of QK^T
    constexpr int NT_O = HD / 8;      // n-tiles of O
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.y, bp = g / heads, head = g - bp * heads;
    const int D = heads * HD;
    const long long row_stride = 3LL * D;
    const int q0 = blockIdx.x * FA_BQ;
    const __half* q_base = qkv + (long long)bp * S * row_stride + head * HD;
    const __half* k_base = q_base + D;
    const __half* v_base = q_base + 2 * D;
    const int n_kt = (S + FA_BK - 1) / FA_BK;
    const bool warp_active = q0 + warp * 16 < S;  // tail q-tiles: idle warps only help loading
    const int r_lo = warp * 16 + (lane >> 2);     // local query row of c0/c1; c2/c3 are r_lo + 8

    TileLoader<HD> ld;
    ld.init(tid, (int)row_stride);
    constexpr uint32_t KV_BYTES = FA_BK * LD * 2;
    const uint32_t sq = ptx::smem_u32(sm.q), sk0 = ptx::smem_u32(sm.k[0]), sv0 = ptx::smem_u32(sm.v[0]);
    ld.load(sq, q_base + (long long)q0 * row_stride, q0, S);
    ptx::cp_async_commit();
    ld.load(sk0, k_base, 0, S);
    ld.load(sv0, v_base, 0, S);
    ptx::cp_async_commit();

    uint32_t q_frag[KSTEPS][4];
    float s_acc[8][4];

    if (BIAS) {
        // ---- rel-pos tables through the MMA path; table tiles double-buffer through k[1] / v[1]
        const int Lh = 2 * gh - 1, Lw = 2 * gw - 1;
        const int ph = (Lh + 63) / 64, pw = (Lw + 63) / 64;
        const int n_pass = ph + pw;
        TileLoader<HD> ldt;
        ldt.init(tid, HD);
        auto issue = [&](int pass) {
            const bool is_h = pass < ph;
            const int p = is_h ? pass : pass - ph;
            ldt.load(((pass & 1) ? sv0 : sk0) + KV_BYTES, (is_h ? Rh : Rw) + (long long)p * 64 * HD, p * 64, is_h ? Lh : Lw);
            ptx::cp_async_commit();
        };
        issue(0);
        for (int pass = 0; pass < n_pass; ++pass) {
            if (pass + 1 < n_pass) { issue(pass + 1); ptx::cp_async_wait<1>(); }
            else ptx::cp_async_wait<0>();
            __syncthreads();
            if (pass == 0) {
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks) {
                    const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                    const int col = ks * 16 + (lane >> 4) * 8;
                    ptx::ldmatrix_x4(ptx::smem_u32(sm.q + row * LD + col), q_frag[ks][0], q_frag[ks][1], q_frag[ks][2], q_frag[ks][3]);
                }
            }
            const bool is_h = pass < ph;
            const int p = is_h ? pass : pass - ph;
            const int L = is_h ? Lh : Lw, gdim = is_h ? gh : gw;
            const int rows_here = min(64, L - p * 64);
            if (warp_active) {
                fa_qk<HD>((pass & 1) ? sm.v[1] : sm.k[1], q_frag, s_acc, lane, (rows_here + 15) >> 4);
                __half* dst = is_h ? sm.rel_h : sm.rel_w;
#pragma unroll
                for (int hrow = 0; hrow < 2; ++hrow) {
                    const int row = r_lo + 8 * hrow;
                    const int t = q0 + row;
                    if (t < S) {
                        const int qh = t / gw;
                        const int qpos = is_h ? qh : t - qh * gw;
#pragma unroll
                        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int j = p * 64 + nt * 8 + 2 * (lane & 3) + e;
                                const int kk = qpos + gdim - 1 - j;
                                if (kk >= 0 && kk < gdim && j < L) dst[row * REL_LD + kk] = __float2half(s_acc[nt][2 * hrow + e]);
                            }
                    }
                }
            }
            __syncthreads();
        }
    } else {
        ptx::cp_async_wait<0>();
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
            const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
            const int col = ks * 16 + (lane >> 4) * 8;
            ptx::ldmatrix_x4(ptx::smem_u32(sm.q + row * LD + col), q_frag[ks][0], q_frag[ks][1], q_frag[ks][2], q_frag[ks][3]);
        }
    }

    constexpr float L2E = 1.4426950408889634f;
    const float sl2 = scale * L2E;
    const bool fast_bias = BIAS && gw == FA_BK;  // one key tile == one key row: rel_h is a per-row scalar, rel_w loop-invariant
    float rw[BIAS ? 32 : 1];
    if (BIAS && fast_bias && warp_active) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = nt * 8 + 2 * (lane & 3) + e;
                rw[nt * 4 + e] = __half2float(sm.rel_w[r_lo * REL_LD + c]) * L2E;
                rw[nt * 4 + 2 + e] = __half2float(sm.rel_w[(r_lo + 8) * REL_LD + c]) * L2E;
            }
    }
    const float inv_gw = 1.0f / (float)gw;

    float o_acc[NT_O][4];
#pragma unroll
    for (int i = 0; i < NT_O; ++i) { o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

    for (int kt = 0; kt < n_kt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < n_kt) {
            const long long toff = (long long)(kt + 1) * FA_BK * row_stride;
            ld.load(sk0 + (buf ^ 1) * KV_BYTES, k_base + toff, (kt + 1) * FA_BK, S);
            ld.load(sv0 + (buf ^ 1) * KV_BYTES, v_base + toff, (kt + 1) * FA_BK, S);
            ptx::cp_async_commit();
            ptx::cp_async_wait<1>();
        } else {
            ptx::cp_async_wait<0>();
        }
        __syncthreads();
        if (warp_active) {
            fa_qk<HD>(sm.k[buf], q_frag, s_acc, lane, 4);
            // ---- scale, bias, mask (log2 domain). Fast path: the key tile is one key row, so rel_h is a per-row scalar
            // bh for the whole tile: it is folded into the softmax reference (p = 2^(s' - (m - bh))) instead of being
            // added to all 64 scores.
            const int kbase = kt * FA_BK;
            float bh[2] = {0.f, 0.f};
            if (BIAS && fast_bias) {
                bh[0] = __half2float(sm.rel_h[r_lo * REL_LD + kt]) * L2E;
                bh[1] = __half2float(sm.rel_h[(r_lo + 8) * REL_LD + kt]) * L2E;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        s_acc[nt][e] = fmaf(s_acc[nt][e], sl2, rw[nt * 4 + e]);
                        s_acc[nt][2 + e] = fmaf(s_acc[nt][2 + e], sl2, rw[nt * 4 + 2 + e]);
                    }
            } else {
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int kcol = kbase + nt * 8 + 2 * (lane & 3) + e;
                        const bool valid = kcol < S;
                        float b_lo = 0.f, b_hi = 0.f;
                        if (BIAS && valid) {
                            const int kh = (int)(((float)kcol + 0.5f) * inv_gw), kw = kcol - kh * gw;
                            b_lo = (__half2float(sm.rel_h[r_lo * REL_LD + kh]) + __half2float(sm.rel_w[r_lo * REL_LD + kw])) * L2E;
                            b_hi = (__half2float(sm.rel_h[(r_lo + 8) * REL_LD + kh]) + __half2float(sm.rel_w[(r_lo + 8) * REL_LD + kw])) * L2E;
                        }
                        s_acc[nt][e] = valid ? fmaf(s_acc[nt][e], sl2, b_lo) : -INFINITY;
                        s_acc[nt][2 + e] = valid ? fmaf(s_acc[nt][2 + e], sl2, b_hi) : -INFINITY;
                    }
            }
            // ---- online softmax
            float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                mx[0] = fmaxf(mx[0], fmaxf(s_acc[nt][0], s_acc[nt][1]));
                mx[1] = fmaxf(mx[1], fmaxf(s_acc[nt][2], s_acc[nt][3]));
            }
            float alpha[2], mref[2], rs[2] = {0.f, 0.f};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
                mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
                const float m_new = fmaxf(m_run[h], mx[h] + bh[h]);
                alpha[h] = ptx::ex2(m_run[h] - m_new);
                m_run[h] = m_new;
                mref[h] = m_new - bh[h];
            }
            uint32_t p_frag[4][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float p0 = ptx::ex2(s_acc[nt][0] - mref[0]), p1 = ptx::ex2(s_acc[nt][1] - mref[0]);
                const float p2 = ptx::ex2(s_acc[nt][2] - mref[1]), p3 = ptx::ex2(s_acc[nt][3] - mref[1]);
                rs[0] += p0 + p1;
                rs[1] += p2 + p3;
                p_frag[nt >> 1][(nt & 1) * 2 + 0] = pack_h2(p0, p1);
                p_frag[nt >> 1][(nt & 1) * 2 + 1] = pack_h2(p2, p3);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) l_run[h] = l_run[h] * alpha[h] + rs[h];
            if (__any_sync(0xffffffffu, alpha[0] != 1.0f || alpha[1] != 1.0f)) {  // the running max moves rarely after the first tiles
#pragma unroll
                for (int i = 0; i < NT_O; ++i) {
                    o_acc[i][0] *= alpha[0]; o_acc[i][1] *= alpha[0];
                    o_acc[i][2] *= alpha[1]; o_acc[i][3] *= alpha[1];
                }
            }
            // ---- O += P V
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {  // 16 keys per step
#pragma unroll
                for (int dp = 0; dp < NT_O / 2; ++dp) {
                    uint32_t b0, b1, b2, b3;
                    const int row = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                    const int col = dp * 16 + (lane >> 4) * 8;
                    ptx::ldmatrix_x4_trans(ptx::smem_u32(sm.v[buf] + row * LD + col), b0, b1, b2, b3);
                    ptx::mma_16816(o_acc[2 * dp], p_frag[kk], b0, b1);
                    ptx::mma_16816(o_acc[2 * dp + 1], p_frag[kk], b2, b3);
                }
            }
        }
        __syncthreads();
    }
    if (!warp_active) return;
    // ---- finalize
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
        l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
    }
    const float inv0 = 1.0f / l_run[0], inv1 = 1.0f / l_run[1];
    const int qr0 = q0 + r_lo, qr1 = qr0 + 8;
    __half* o0 = out + ((long long)bp * S + qr0) * D + head * HD + 2 * (lane & 3);
    __half* o1 = out + ((long long)bp * S + qr1) * D + head * HD + 2 * (lane & 3);
#pragma unroll
    for (int i = 0; i < NT_O; ++i) {
        if (qr0 < S) *reinterpret_cast<uint32_t*>(o0 + i * 8) = pack_h2(o_acc[i][0] * inv0, o_acc[i][1] * inv0);
        if (qr1 < S) *reinterpret_cast<uint32_t*>(o1 + i * 8) = pack_h2(o_acc[i][2] * inv1, o_acc[i][3] * inv1);
    }
}

// ------------------------------------------------------------------------------------------ rel-pos tables only
// The bias prologue of flash_kernel as a kernel of its own, for the tcgen05 attention (flash_tc.cu): per 64 queries of one
// (image, head), G = Q R^T through the MMA path, scattered to rel[q, k] = G[q, qpos + g - 1 - k], then written out
// (times log2 e, fp16) as bias_h / bias_w [(g*heads + head)*S + q][64].
template <int HD>
struct RtSmem {
    static constexpr int LD = HD + 8;
    __half q[FA_BQ * LD];
    __half t[2][FA_BK * LD];
    __half rel_h[FA_BQ * REL_LD];
    __half rel_w[FA_BQ * REL_LD];
};

template <int HD>
__global__ void __launch_bounds__(FA_THREADS)
relpos_tables_kernel(const __half* __restrict__ qkv, int S, int heads, const __half* __restrict__ Rh, const __half* __restrict__ Rw,
                     int gh, int gw, __half* __restrict__ bias_h, __half* __restrict__ bias_w) {
    extern __shared__ __align__(16) uint8_t rt_smem_raw[];
    using Smem = RtSmem<HD>;
    Smem& sm = *reinterpret_cast<Smem*>(rt_smem_raw);
    constexpr int LD = Smem::LD;
    constexpr int KSTEPS = HD / 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.y, bp = g / heads, head = g - bp * heads;
    const int D = heads * HD;
    const long long row_stride = 3LL * D;
    const int q0 = blockIdx.x * FA_BQ;
    const __half* q_base = qkv + (long long)bp * S * row_stride + head * HD;
    const int r_lo = warp * 16 + (lane >> 2);
    for (int i = tid; i < FA_BQ * REL_LD; i += FA_THREADS) { sm.rel_h[i] = __float2half(0.f); sm.rel_w[i] = __float2half(0.f); }
    TileLoader<HD> ld, ldt;
    ld.init(tid, (int)row_stride);
    ldt.init(tid, HD);
    constexpr uint32_t T_BYTES = FA_BK * LD * 2;
    const uint32_t sq = ptx::smem_u32(sm.q), st0 = ptx::smem_u32(sm.t[0]);
    ld.load(sq, q_base + (long long)q0 * row_stride, q0, S);
    ptx::cp_async_commit();
    const int Lh = 2 * gh - 1, Lw = 2 * gw - 1;
    const int ph = (Lh + 63) / 64, pw = (Lw + 63) / 64;
    const int n_pass = ph + pw;
    auto issue = [&](int pass) {
        const bool is_h = pass < ph;
        const int p = is_h ? pass : pass - ph;
        ldt.load(st0 + (pass & 1) * T_BYTES, (is_h ? Rh : Rw) + (long long)p * 64 * HD, p * 64, is_h ? Lh : Lw);
        ptx::cp_async_commit();
    };
    issue(0);
    uint32_t q_frag[KSTEPS][4];
    float s_acc[8][4];
    for (int pass = 0; pass < n_pass; ++pass) {
        if (pass + 1 < n_pass) { issue(pass + 1); ptx::cp_async_wait<1>(); }
        else ptx::cp_async_wait<0>();
        __syncthreads();
        if (pass == 0) {
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int col = ks * 16 + (lane >> 4) * 8;
                ptx::ldmatrix_x4(ptx::smem_u32(sm.q + row * LD + col), q_frag[ks][0], q_frag[ks][1], q_frag[ks][2], q_frag[ks][3]);
            }
        }
        const bool is_h = pass < ph;
        const int p = is_h ? pass : pass - ph;
        const int L = is_h ? Lh : Lw, gdim = is_h ? gh : gw;
        const int rows_here = min(64, L - p * 64);
        fa_qk<HD>(sm.t[pass & 1], q_frag, s_acc, lane, (rows_here + 15) >> 4);
        __half* dst = is_h ? sm.rel_h : sm.rel_w;
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
            const int row = r_lo + 8 * hrow;
            const int t = q0 + row;
            if (t < S) {
                const int qh = t / gw;
                const int qpos = is_h ? qh : t - qh * gw;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int j = p * 64 + nt * 8 + 2 * (lane & 3) + e;
                        const int kk = qpos + gdim - 1 - j;
                        if (kk >= 0 && kk < gdim && j < L) dst[row * REL_LD + kk] = __float2half(s_acc[nt][2 * hrow + e] * 1.4426950408889634f);
                    }
            }
        }
        __syncthreads();
    }
    // coalesced write-out: 64 rows x 64 halves per table
    for (int i = tid; i < FA_BQ * 32; i += FA_THREADS) {
        const int r = i >> 5, c = (i & 31) * 2;
        if (q0 + r >= S) continue;
        const long long o = ((long long)g * S + q0 + r) * 64 + c;
        *reinterpret_cast<__half2*>(bias_h + o) = *reinterpret_cast<const __half2*>(sm.rel_h + r * REL_LD + c);
        *reinterpret_cast<__half2*>(bias_w + o) = *reinterpret_cast<const __half2*>(sm.rel_w + r * REL_LD + c);
    }
}

// ------------------------------------------------------------------------------------------ windowed attention
// One CTA per (window, head): the whole K and V of the window (S <= 208 keys, 14 x 14 = 196 for SAM) stay in shared
// memory, so the seven warps run their 16-query m-tiles without any block-level synchronisation after the load.
// Q fragments are read straight from global memory in the MMA A-fragment layout (31 KB per CTA, L1-resident).
constexpr int WA_WARPS = 7, WA_THREADS = WA_WARPS * 32, WA_MAXS = 208;

template <int HD>
struct WaSmem {
    static constexpr int LD = HD + 8;
    static constexpr int GLD = 40;           // row stride (halves) of the 32-column bias operands: 80 B, conflict-free ldmatrix
    __half k[WA_MAXS * LD];
    __half v[WA_MAXS * LD];
    __half th[32 * LD];                      // rel-pos tables (L = 2g-1 <= 29 rows for g <= 15), zero-padded to 32 rows
    __half tw[32 * LD];
    __half sel[WA_MAXS * GLD];               // Sel[k][j] = 1 where j == kh(k) or j == gh + kw(k): bias = Gsel * Sel^T
    __half gsel[WA_WARPS][16 * GLD];         // per m-tile: [rel_h(q, 0..gh-1) | rel_w(q, 0..gw-1)] / scale
};

// The decomposed rel-pos bias is added by the tensor cores: bias[q,k] = rel_h[q, kh(k)] + rel_w[q, kw(k)] is the product
// of the per-query row Gsel[q, :] = [rel_h | rel_w] (gh + gw <= 30 values) with a constant 0/1 selection matrix, i.e.
// two extra k-steps of the QK^T MMA instead of ~20 scalar instructions per score. Gsel is stored divided by `scale`
// so that one factor scale*log2(e) applies to the whole accumulator, FlashAttention-2 style: p = 2^(acc*c - m*c).
template <int HD>
__global__ void __launch_bounds__(WA_THREADS, 2)
window_attn_kernel(const __half* __restrict__ qkv, int S, int heads, float scale, const __half* __restrict__ Rh,
                   const __half* __restrict__ Rw, int gh, int gw, __half* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t wa_smem_raw[];
    using Smem = WaSmem<HD>;
    Smem& sm = *reinterpret_cast<Smem*>(wa_smem_raw);
    constexpr int LD = Smem::LD, GLD = Smem::GLD, KSTEPS = HD / 16, NT_O = HD / 8, CH = HD / 8;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x, bp = g / heads, head = g - bp * heads;
    const int D = heads * HD;
    const long long row_stride = 3LL * D;
    const __half* q_base = qkv + (long long)bp * S * row_stride + head * HD;
    const __half* k_base = q_base + D;
    const __half* v_base = q_base + 2 * D;
    const int Lh = 2 * gh - 1, Lw = 2 * gw - 1;
    const int s_pad = (S + 15) & ~15;

    for (int i = tid; i < s_pad * CH; i += WA_THREADS) {
        const int r = i / CH, c = i - r * CH;
        const bool ok = r < S;
        const long long off = (long long)(ok ? r : 0) * row_stride + c * 8;
        ptx::cp_async16(ptx::smem_u32(sm.k + r * LD + c * 8), k_base + off, ok);
        ptx::cp_async16(ptx::smem_u32(sm.v + r * LD + c * 8), v_base + off, ok);
    }
    for (int i = tid; i < 32 * CH; i += WA_THREADS) {
        const int r = i / CH, c = i - r * CH;
        ptx::cp_async16(ptx::smem_u32(sm.th + r * LD + c * 8), Rh + (long long)(r < Lh ? r : 0) * HD + c * 8, r < Lh);
        ptx::cp_async16(ptx::smem_u32(sm.tw + r * LD + c * 8), Rw + (long long)(r < Lw ? r : 0) * HD + c * 8, r < Lw);
    }
    ptx::cp_async_commit();
    // selection matrix and zeroed Gsel tiles (columns gh+gw..31 stay zero)
    for (int k = tid; k < s_pad; k += WA_THREADS) {  // one key row per thread: zero 32 columns, then set its two ones
        uint4* row = reinterpret_cast<uint4*>(sm.sel + k * GLD);
        row[0] = row[1] = row[2] = row[3] = make_uint4(0, 0, 0, 0);
        if (k < S) {
            const int kh = k / gw, kw = k - kh * gw;
            sm.sel[k * GLD + kh] = __float2half(1.0f);
            sm.sel[k * GLD + gh + kw] = __float2half(1.0f);
        }
    }
    for (int i = tid; i < WA_WARPS * 16 * GLD; i += WA_THREADS) (&sm.gsel[0][0])[i] = __float2half(0.0f);
    ptx::cp_async_wait<0>();
    __syncthreads();

    constexpr float L2E = 1.4426950408889634f;
    const float sl2 = scale * L2E, inv_scale = 1.0f / scale;
    const float inv_gw = 1.0f / (float)gw;
    const int n_mt = s_pad >> 4, n_kt = (S + 63) >> 6;
    __half* gs = sm.gsel[warp];
    const int rl = lane >> 2;  // local row of c0/c1; c2/c3 are rl + 8

    for (int mt = warp; mt < n_mt; mt += WA_WARPS) {
        const int q0 = mt * 16;
        // ---- Q fragments from global: a0=(rl, 2*(lane&3)), a1=(rl+8, .), a2=(rl, .+8), a3=(rl+8, .+8)
        uint32_t q_frag[KSTEPS][4];
        {
            const int r0 = q0 + rl, r1 = r0 + 8;
            const __half* p0 = q_base + (long long)min(r0, S - 1) * row_stride + 2 * (lane & 3);
            const __half* p1 = q_base + (long long)min(r1, S - 1) * row_stride + 2 * (lane & 3);
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                q_frag[ks][0] = r0 < S ? *reinterpret_cast<const uint32_t*>(p0 + ks * 16) : 0u;
                q_frag[ks][1] = r1 < S ? *reinterpret_cast<const uint32_t*>(p1 + ks * 16) : 0u;
                q_frag[ks][2] = r0 < S ? *reinterpret_cast<const uint32_t*>(p0 + ks * 16 + 8) : 0u;
                q_frag[ks][3] = r1 < S ? *reinterpret_cast<const uint32_t*>(p1 + ks * 16 + 8) : 0u;
            }
        }
        float s_acc[8][4];
        // ---- rel-pos: G = Q T^T (unscaled q), scattered to Gsel[row][kk] with kk = qpos + g - 1 - j
#pragma unroll
        for (int tbl = 0; tbl < 2; ++tbl) {
            const int L = tbl == 0 ? Lh : Lw, gdim = tbl == 0 ? gh : gw, col0 = tbl == 0 ? 0 : gh;
            fa_qk<HD>(tbl == 0 ? sm.th : sm.tw, q_frag, s_acc, lane, (L + 15) >> 4);
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
                const int t = q0 + rl + 8 * hrow;
                if (t < S) {
                    const int qh = (int)(((float)t + 0.5f) * inv_gw);
                    const int qpos = tbl == 0 ? qh : t - qh * gw;
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt)  // L <= 29: only the first 32 table rows exist
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int j = nt * 8 + 2 * (lane & 3) + e;
                            const int kk = qpos + gdim - 1 - j;
                            if (kk >= 0 && kk < gdim && j < L)
                                gs[(rl + 8 * hrow) * GLD + col0 + kk] = __float2half(s_acc[nt][2 * hrow + e] * inv_scale);
                        }
                }
            }
        }
        __syncwarp();
        uint32_t g_frag[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const int row = (lane & 7) + ((lane >> 3) & 1) * 8;
            const int col = ks * 16 + (lane >> 4) * 8;
            ptx::ldmatrix_x4(ptx::smem_u32(gs + row * GLD + col), g_frag[ks][0], g_frag[ks][1], g_frag[ks][2], g_frag[ks][3]);
        }

        float o_acc[NT_O][4];
#pragma unroll
        for (int i = 0; i < NT_O; ++i) { o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f; }
        float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};  // m_run in raw accumulator units
        for (int kt = 0; kt < n_kt; ++kt) {
            const int kbase = kt * 64;
            const int keys_here = min(64, s_pad - kbase);        // multiple of 16
            const int n_pairs = keys_here >> 4;
            fa_qk<HD>(sm.k + kbase * LD, q_frag, s_acc, lane, n_pairs);
            // ---- + Gsel * Sel^T (two k-steps over the 32 bias columns)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
                for (int np = 0; np < 4; ++np) {
                    if (np < n_pairs) {
                        uint32_t b0, b1, b2, b3;
                        const int row = kbase + np * 16 + (lane & 7) + ((lane >> 4) << 3);
                        const int col = ks * 16 + ((lane >> 3) & 1) * 8;
                        ptx::ldmatrix_x4(ptx::smem_u32(sm.sel + row * GLD + col), b0, b1, b2, b3);
                        ptx::mma_16816(s_acc[2 * np], g_frag[ks], b0, b1);
                        ptx::mma_16816(s_acc[2 * np + 1], g_frag[ks], b2, b3);
                    }
                }
            }
            if (kbase + 64 > S) {  // only the last key tile has columns past S
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        if (kbase + nt * 8 + 2 * (lane & 3) + e >= S) { s_acc[nt][e] = -INFINITY; s_acc[nt][2 + e] = -INFINITY; }
            }
            float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                mx[0] = fmaxf(mx[0], fmaxf(s_acc[nt][0], s_acc[nt][1]));
                mx[1] = fmaxf(mx[1], fmaxf(s_acc[nt][2], s_acc[nt][3]));
            }
            float alpha[2], mneg[2], rs[2] = {0.f, 0.f};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
                mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
                alpha[h] = ptx::ex2((m_run[h] - mx[h]) * sl2);
                m_run[h] = mx[h];
                mneg[h] = -mx[h] * sl2;
            }
            uint32_t p_frag[4][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const float p0 = ptx::ex2(fmaf(s_acc[nt][0], sl2, mneg[0])), p1 = ptx::ex2(fmaf(s_acc[nt][1], sl2, mneg[0]));
                const float p2 = ptx::ex2(fmaf(s_acc[nt][2], sl2, mneg[1])), p3 = ptx::ex2(fmaf(s_acc[nt][3], sl2, mneg[1]));
                rs[0] += p0 + p1;
                rs[1] += p2 + p3;
                p_frag[nt >> 1][(nt & 1) * 2 + 0] = pack_h2(p0, p1);
                p_frag[nt >> 1][(nt & 1) * 2 + 1] = pack_h2(p2, p3);
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) l_run[h] = l_run[h] * alpha[h] + rs[h];
#pragma unroll
            for (int i = 0; i < NT_O; ++i) {
                o_acc[i][0] *= alpha[0]; o_acc[i][1] *= alpha[0];
                o_acc[i][2] *= alpha[1]; o_acc[i][3] *= alpha[1];
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                if (kk * 16 < keys_here) {
#pragma unroll
                    for (int dp = 0; dp < NT_O / 2; ++dp) {
                        uint32_t b0, b1, b2, b3;
                        const int row = kbase + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                        const int col = dp * 16 + (lane >> 4) * 8;
                        ptx::ldmatrix_x4_trans(ptx::smem_u32(sm.v + row * LD + col), b0, b1, b2, b3);
                        ptx::mma_16816(o_acc[2 * dp], p_frag[kk], b0, b1);
                        ptx::mma_16816(o_acc[2 * dp + 1], p_frag[kk], b2, b3);
                    }
                }
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
            l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
        }
        const float inv0 = 1.0f / l_run[0], inv1 = 1.0f / l_run[1];
        const int qr0 = q0 + rl, qr1 = qr0 + 8;
        __half* o0 = out + ((long long)bp * S + qr0) * D + head * HD + 2 * (lane & 3);
        __half* o1 = out + ((long long)bp * S + qr1) * D + head * HD + 2 * (lane & 3);
#pragma unroll
        for (int i = 0; i < NT_O; ++i) {
            if (qr0 < S) *reinterpret_cast<uint32_t*>(o0 + i * 8) = pack_h2(o_acc[i][0] * inv0, o_acc[i][1] * inv0);
            if (qr1 < S) *reinterpret_cast<uint32_t*>(o1 + i * 8) = pack_h2(o_acc[i][2] * inv1, o_acc[i][3] * inv1);
        }
        __syncwarp();  // Gsel is rewritten by the warp's next m-tile
    }
}

template <int HD>
int launch_window(const __half* qkv, int Gb, int S, int heads, float scale, const __half* Rh, const __half* Rw, int gh, int gw,
                  __half* out, cudaStream_t stream) {
    static unsigned long long configured = 0;  // one bit per device: function attributes are per device
    const int cfg_dev = cvb_current_device();
    const int smem = (int)sizeof(WaSmem<HD>);
    if (!((configured >> cfg_dev) & 1ull)) {
        CVB_CUDA(cudaFuncSetAttribute(window_attn_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured |= 1ull << cfg_dev;
    }
    window_attn_kernel<HD><<<Gb * heads, WA_THREADS, smem, stream>>>(qkv, S, heads, scale, Rh, Rw, gh, gw, out);
    cvb_note_launches(1);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}

template <int HD, bool BIAS>
int launch_flash(const __half* qkv, int Gb, int S, int heads, float scale, const __half* Rh, const __half* Rw,
                 int gh, int gw, __half* out, cudaStream_t stream) {
    static unsigned long long configured = 0;  // one bit per device: function attributes are per device
    const int cfg_dev = cvb_current_device();
    const int smem = (int)sizeof(FaSmem<HD, BIAS>);
    if (!((configured >> cfg_dev) & 1ull)) {
        CVB_CUDA(cudaFuncSetAttribute(flash_kernel<HD, BIAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured |= 1ull << cfg_dev;
    }
    dim3 grid(cdiv(S, FA_BQ), Gb * heads);
    flash_kernel<HD, BIAS><<<grid, FA_THREADS, smem, stream>>>(qkv, S, heads, scale, Rh, Rw, gh, gw, out);
    cvb_note_launches(1);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}

}  // namespace

int op_relpos_tables(const __half* qkv, int Gb, int S, int heads, int hd, const __half* Rh, const __half* Rw, int gh, int gw,
                     __half* bias_h, __half* bias_w, cudaStream_t stream) {
    CVB_CHECK(qkv && Rh && Rw && bias_h && bias_w && hd == 80 && gh <= 64 && gw <= 64 && gh * gw == S, CVB_ESHAPE,
              "relpos_tables: needs head dim 80 and a token grid of at most 64 x 64");
    static unsigned long long configured = 0;  // one bit per device: function attributes are per device
    const int cfg_dev = cvb_current_device();
    const int smem = (int)sizeof(RtSmem<80>);
    if (!((configured >> cfg_dev) & 1ull)) {
        CVB_CUDA(cudaFuncSetAttribute(relpos_tables_kernel<80>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured |= 1ull << cfg_dev;
    }
    relpos_tables_kernel<80><<<dim3(cdiv(S, FA_BQ), Gb * heads), FA_THREADS, smem, stream>>>(qkv, S, heads, Rh, Rw, gh, gw, bias_h, bias_w);
    cvb_note_launches(1);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}

int op_attention(const __half* qkv, int Gb, int S, int heads, int hd, float scale, const __half* Rh, const __half* Rw,
                 int gh, int gw, __half* out, cudaStream_t stream) {
    CVB_CHECK(qkv && out && Gb > 0 && S > 0, CVB_EARG, "attention: null operand or empty shape");
    CVB_CHECK((Rh == nullptr) == (Rw == nullptr), CVB_EARG, "attention: Rh and Rw must both be set or both null");
    if (Rh) CVB_CHECK(gh * gw == S && gh <= 64 && gw <= 64, CVB_ESHAPE, "attention: bias grid %dx%d does not match S=%d", gh, gw, S);
    if (!Rh) { gh = 1; gw = S; }
    if (Rh && S <= WA_MAXS && gh <= 15 && gw <= 15 && gh + gw <= 32) {  // SAM windows: whole K/V resident in shared memory
        if (hd == 80) return launch_window<80>(qkv, Gb, S, heads, scale, Rh, Rw, gh, gw, out, stream);
        if (hd == 64) return launch_window<64>(qkv, Gb, S, heads, scale, Rh, Rw, gh, gw, out, stream);
    }
    if (hd == 80) return Rh ? launch_flash<80, true>(qkv, Gb, S, heads, scale, Rh, Rw, gh, gw, out, stream)
                            : launch_flash<80, false>(qkv, Gb, S, heads, scale, Rh, Rw, gh, gw, out, stream);
    if (hd == 64) return Rh ? launch_flash<64, true>(qkv, Gb, S, heads, scale, Rh, Rw, gh, gw, out, stream)
                            : launch_flash<64, false>(qkv, Gb, S, heads, scale, Rh, Rw, gh, gw, out, stream);
    cvb_set_error("attention: head dim %d not supported (64 or 80)", hd);
    return CVB_ESHAPE;
}
