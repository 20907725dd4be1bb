// tc_gemm.cu -- persistent, warp-specialised tcgen05 + TMA tile engine for sm_100a.
//
// One kernel serves every dense contraction of the CellViT forward:
//   * linear layers (patch-embed, QKV, attn-out, MLP)                 -> GEMM mode
//   * ConvTranspose2d k2 s2 (= GEMM with N = 4*Cout + scatter epilogue) -> GEMM mode, TC_EPI_CONVT
//   * Conv3x3 pad 1 over NHWC fp16 (implicit GEMM)                     -> CONV mode: the A tile of tap (dy,dx)
//     is one 4-D TMA box shifted by (dx-1, dy-1); out-of-image pixels are zero-filled by the TMA unit, so
//     there is no im2col buffer. torch.cat([skip, x], 1) is a K-split over two tensor maps.
//
// CTA = 12 warps: w0 / w3 TMA producers for A / B (1 lane each), w1 MMA issuer (1 lane), w2 TMEM allocator, w4..11 epilogue
// (TMEM lane quadrant = warp & 3; the two warps of a quadrant take alternating 32-column chunks). Tile = 128 x block_n fp32 accumulator in TMEM, double buffered so the
// epilogue of tile i overlaps the main loop of tile i+1. Operands are fp16 (11-bit mantissa: the reference's
// own AMP mode, cell_detection.py:314-318), accumulation fp32.
#include <cudaTypedefs.h>

#include <vector>

#include "tc_gemm.h"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                       // 64 fp16 = one 128-byte swizzle atom
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int MAX_STAGES = 8;
constexpr int NUM_THREADS = 384;  // 4 control warps + 8 epilogue warps
constexpr int SMEM_BUDGET = 200 * 1024;

struct TcParams {
    int M, N, K;
    int block_n, n_tiles_n, n_tiles, num_kb, stages;
    int conv;
    int chunks0, chunks_per_tap;  // conv: 64-channel chunks of source 0 / of both sources
    int H, W, TW;                 // conv geometry: image size, tile width (tile height = 128 / TW)
    int l2_prefetch;              // GEMM mode: A k-blocks prefetched into L2 ahead of the smem ring (0 = off)
    uint32_t tmem_cols;
    TcEpilogue epi;
};

// shared-memory block of the fused 1x1 head (floats): weights transposed to [channel][8 classes], bias, BN
// (scale, shift) pairs, and the partial-sum exchange buffer of the two epilogue warp groups
constexpr int HEAD_S_BIAS = 64 * 8;
constexpr int HEAD_S_SS = HEAD_S_BIAS + 8;
constexpr int HEAD_S_PART = HEAD_S_SS + 128;
constexpr int HEAD_S_FLOATS = HEAD_S_PART + 8 * 128;
constexpr int HEAD_S_BYTES = HEAD_S_FLOATS * 4;

__device__ __forceinline__ void head_smem_fill(float* hs, const TcEpilogue& e, int t, int nthreads) {
    for (int i = t; i < 64 * 8; i += nthreads) {
        const int n = i >> 3, k = i & 7;
        hs[i] = k < e.head_nc ? e.head_w[k * 64 + n] : 0.0f;
    }
    if (t < 8) hs[HEAD_S_BIAS + t] = t < e.head_nc ? e.head_b[t] : 0.0f;
    for (int i = t; i < 64; i += nthreads) {
        hs[HEAD_S_SS + 2 * i] = e.scale[i];
        hs[HEAD_S_SS + 2 * i + 1] = e.shift[i];
    }
}

__device__ __forceinline__ int map_out_row(const TcEpilogue& e, int m) {
    if (e.row_map == TC_ROW_IDENTITY) return m;
    if (e.row_map == TC_ROW_SEQ) return m + (m / e.row_seq) * e.row_pad + e.row_off;
    const int ws = e.win_size, g = e.win_grid;
    const int per_img = g * g * ws * ws;
    if (e.row_map == TC_ROW_TO_WINDOW) {  // image_encoder.py:263-288 window_partition of raster row m
        const int hw = e.tok_h * e.tok_w;
        const int b = m / hw;
        const int rem = m - b * hw;
        const int y = rem / e.tok_w, x = rem - y * e.tok_w;
        const int wy = y / ws, wx = x / ws;
        return b * per_img + (wy * g + wx) * ws * ws + (y - wy * ws) * ws + (x - wx * ws);
    }
    // TC_ROW_WINDOW (image_encoder.py:291-318 window_unpartition)
    const int b = m / per_img;
    int rem = m - b * per_img;
    const int win = rem / (ws * ws);
    const int t = rem - win * ws * ws;
    const int y = (win / g) * ws + t / ws;
    const int x = (win % g) * ws + t % ws;
    if (y >= e.tok_h || x >= e.tok_w) return -1;
    return (b * e.tok_h + y) * e.tok_w + x;
}

template <int ACT>
__device__ __forceinline__ float act_t(float v) {
    if (ACT == TC_ACT_RELU) return fmaxf(v, 0.0f);
    if (ACT == TC_ACT_GELU) return gelu_erf_fast(v);
    return v;
}

// F16 epilogue of one accumulator, specialised at compile time on the activation and on the presence of a per-column
// scale (folded BatchNorm): the generic path below re-tests `act` / `scale` per element, which doubled the instruction
// count of the GELU epilogue (ncu: 32 instead of 16 issue slots per element, making fc1 epilogue-bound).
// 16 accumulator columns -> (scale,) shift, activation, fp16, two 16-byte stores
template <int ACT, bool HAS_SCALE>
__device__ __forceinline__ void epi_f16_store16(const TcEpilogue& e, const uint32_t (&acc)[16], int nb, __half* o) {
#pragma unroll
    for (int j = 0; j < 16; j += 8) {
        float v[8];
        const float4 h0 = __ldg(reinterpret_cast<const float4*>(e.shift + nb + j));
        const float4 h1 = __ldg(reinterpret_cast<const float4*>(e.shift + nb + j + 4));
        if (HAS_SCALE) {
            const float4 s0 = __ldg(reinterpret_cast<const float4*>(e.scale + nb + j));
            const float4 s1 = __ldg(reinterpret_cast<const float4*>(e.scale + nb + j + 4));
            v[0] = fmaf(__uint_as_float(acc[j + 0]), s0.x, h0.x); v[1] = fmaf(__uint_as_float(acc[j + 1]), s0.y, h0.y);
            v[2] = fmaf(__uint_as_float(acc[j + 2]), s0.z, h0.z); v[3] = fmaf(__uint_as_float(acc[j + 3]), s0.w, h0.w);
            v[4] = fmaf(__uint_as_float(acc[j + 4]), s1.x, h1.x); v[5] = fmaf(__uint_as_float(acc[j + 5]), s1.y, h1.y);
            v[6] = fmaf(__uint_as_float(acc[j + 6]), s1.z, h1.z); v[7] = fmaf(__uint_as_float(acc[j + 7]), s1.w, h1.w);
        } else {
            v[0] = __uint_as_float(acc[j + 0]) + h0.x; v[1] = __uint_as_float(acc[j + 1]) + h0.y;
            v[2] = __uint_as_float(acc[j + 2]) + h0.z; v[3] = __uint_as_float(acc[j + 3]) + h0.w;
            v[4] = __uint_as_float(acc[j + 4]) + h1.x; v[5] = __uint_as_float(acc[j + 5]) + h1.y;
            v[6] = __uint_as_float(acc[j + 6]) + h1.z; v[7] = __uint_as_float(acc[j + 7]) + h1.w;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = act_t<ACT>(v[k]);
        *reinterpret_cast<uint4*>(o + j) = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
    }
}

// F16 epilogue of one accumulator, specialised at compile time on the activation and on the presence of a per-column
// scale (folded BatchNorm): the generic path below re-tests `act` / `scale` per element, which doubled the instruction
// count of the GELU epilogue (ncu: 32 instead of 16 issue slots per element, making fc1 epilogue-bound). The TMEM
// reads are software-pipelined in 16-column halves: the next half is in flight while the current one is computed
// (the tcgen05.ld -> first-use stall was 16 % of the epilogue's samples).
template <int ACT, bool HAS_SCALE, class Release>
__device__ __forceinline__ void epilogue_f16_fast(const TcEpilogue& e, uint32_t t_addr, size_t out_off, bool store, int n0, int n_chunks,
                                                  int half, int last_c, bool do_release, Release release) {
    if (last_c < 0) {
        if (do_release) release();
        return;
    }
    __half* obase = reinterpret_cast<__half*>(e.out) + out_off + n0;
    uint32_t a[16], b[16];
    ptx::tmem_ld16(t_addr + half * 32, a);
    ptx::tmem_ld_wait();
    for (int c = half; c < n_chunks; c += 2) {
        ptx::tmem_ld16(t_addr + c * 32 + 16, b);
        if (store) epi_f16_store16<ACT, HAS_SCALE>(e, a, n0 + c * 32, obase + c * 32);
        ptx::tmem_ld_wait();
        const bool more = c + 2 < n_chunks;
        if (more) ptx::tmem_ld16(t_addr + (c + 2) * 32, a);
        else if (do_release) release();  // every TMEM read of this thread has completed
        if (store) epi_f16_store16<ACT, HAS_SCALE>(e, b, n0 + c * 32 + 16, obase + c * 32 + 16);
        if (more) ptx::tmem_ld_wait();
    }
}

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == TC_ACT_RELU) return fmaxf(v, 0.0f);
    if (act == TC_ACT_GELU) return gelu_erf_fast(v);
    return v;
}

// Drains one 128 x (32 * n_chunks) fp32 accumulator (TMEM address t_addr already carries the lane quadrant) through
// the fused epilogue. `m` is the global output row of this thread; `release()` hands the TMEM buffer back to the MMA
// issuer and is called right after this thread's last tcgen05.ld when `do_release` is set.
template <class Release>
__device__ __forceinline__ void epilogue_tile(const TcEpilogue& e, uint32_t t_addr, int m, bool row_ok, int n0, int n_chunks,
                                              int half, int last_c, float* head_w_s, bool do_release, Release release) {
    if (e.kind == TC_EPI_HEAD) {
        // Fused BN + ReLU + 1x1 head: the two warps of a lane quadrant each reduce 32 of the 64 channels of their
        // pixel, exchange the partial class sums through shared memory and the lower warp stores fp32 NCHW.
        // head_s: hw[64][8] | hb[8] | (scale, shift)[64] | part[8][128]   (filled by head_smem_fill)
        const float4* hw4 = reinterpret_cast<const float4*>(head_w_s);
        const float* hb = head_w_s + HEAD_S_BIAS;
        const float2* ss = reinterpret_cast<const float2*>(head_w_s + HEAD_S_SS);
        float* part = head_w_s + HEAD_S_PART;
        uint32_t acc[32];
        ptx::tmem_ld32(t_addr + half * 32, acc);
        ptx::tmem_ld_wait();
        if (do_release) release();
        float hs[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) hs[k] = 0.0f;
        const int nc = e.head_nc;
        if (nc <= 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int n = half * 32 + j;
                const float2 sc = ss[n];
                const float v = fmaxf(fmaf(__uint_as_float(acc[j]), sc.x, sc.y), 0.0f);
                const float4 w0 = hw4[2 * n];
                hs[0] = fmaf(w0.x, v, hs[0]); hs[1] = fmaf(w0.y, v, hs[1]);
            }
        } else if (nc <= 4) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int n = half * 32 + j;
                const float2 sc = ss[n];
                const float v = fmaxf(fmaf(__uint_as_float(acc[j]), sc.x, sc.y), 0.0f);
                const float4 w0 = hw4[2 * n];
                hs[0] = fmaf(w0.x, v, hs[0]); hs[1] = fmaf(w0.y, v, hs[1]); hs[2] = fmaf(w0.z, v, hs[2]); hs[3] = fmaf(w0.w, v, hs[3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int n = half * 32 + j;
                const float2 sc = ss[n];
                const float v = fmaxf(fmaf(__uint_as_float(acc[j]), sc.x, sc.y), 0.0f);
                const float4 w0 = hw4[2 * n], w1 = hw4[2 * n + 1];
                hs[0] = fmaf(w0.x, v, hs[0]); hs[1] = fmaf(w0.y, v, hs[1]); hs[2] = fmaf(w0.z, v, hs[2]); hs[3] = fmaf(w0.w, v, hs[3]);
                hs[4] = fmaf(w1.x, v, hs[4]); hs[5] = fmaf(w1.y, v, hs[5]); hs[6] = fmaf(w1.z, v, hs[6]); hs[7] = fmaf(w1.w, v, hs[7]);
            }
        }
        const int prow = (int)(threadIdx.x & 127);  // quadrant * 32 + lane
        const int bar_id = 1 + (prow >> 5);
        if (half == 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < nc) part[k * 128 + prow] = hs[k];
        }
        ptx::named_bar_sync(bar_id, 64);
        if (half == 0 && row_ok) {
            const int img = m / e.head_hw, pix = m - img * e.head_hw;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                hs[k] = hs[k] + part[k * 128 + prow] + hb[k];
                if (k < nc) e.head_out[((size_t)img * nc + k) * e.head_hw + pix] = hs[k];
            }
            if (e.head_argmax != nullptr) {
                // the arg-max plane the post-processing consumes (cellvit.py:369-375: torch.argmax of the soft-maxed map = arg-max
                // of the logits, first maximum wins), written here so that cvb_postproc_argmax reads 1 B/px instead of 4 nc B/px
                int best = 0;
                float bv = hs[0];
#pragma unroll
                for (int k = 1; k < 8; ++k)
                    if (k < e.head_argmax_nc && hs[k] > bv) { bv = hs[k]; best = k; }
                e.head_argmax[(size_t)img * e.head_hw + pix] = (uint8_t)best;
            }
        }
        ptx::named_bar_sync(bar_id, 64);  // part[] may be overwritten by the next tile only after it was read
        return;
    }

    int orow = -1;
    size_t out_off = 0;
    int ct_b = 0, ct_y = 0, ct_x = 0;
    if (row_ok) {
        if (e.kind == TC_EPI_CONVT) {
            const int hw = e.ct_hin * e.ct_win;
            ct_b = m / hw;
            const int rem = m - ct_b * hw;
            ct_y = rem / e.ct_win;
            ct_x = rem - ct_y * e.ct_win;
            orow = m;
        } else {
            orow = map_out_row(e, m);
            out_off = (size_t)(orow < 0 ? 0 : orow) * (size_t)e.ldc;
        }
    }
    const float* res_row = nullptr;
    if (e.kind == TC_EPI_RES_F32 && e.res != nullptr && orow >= 0) {
        const long long rr = e.res_mod > 0 ? (long long)(m % e.res_mod) + e.res_off : (long long)orow;
        res_row = e.res + rr * e.ldres;
    }
    if (e.kind == TC_EPI_F16 && e.shift != nullptr) {
        const bool st = orow >= 0;
        if (e.scale != nullptr) {
            if (e.act == TC_ACT_RELU) epilogue_f16_fast<TC_ACT_RELU, true>(e, t_addr, out_off, st, n0, n_chunks, half, last_c, do_release, release);
            else if (e.act == TC_ACT_GELU) epilogue_f16_fast<TC_ACT_GELU, true>(e, t_addr, out_off, st, n0, n_chunks, half, last_c, do_release, release);
            else epilogue_f16_fast<TC_ACT_NONE, true>(e, t_addr, out_off, st, n0, n_chunks, half, last_c, do_release, release);
        } else {
            if (e.act == TC_ACT_RELU) epilogue_f16_fast<TC_ACT_RELU, false>(e, t_addr, out_off, st, n0, n_chunks, half, last_c, do_release, release);
            else if (e.act == TC_ACT_GELU) epilogue_f16_fast<TC_ACT_GELU, false>(e, t_addr, out_off, st, n0, n_chunks, half, last_c, do_release, release);
            else epilogue_f16_fast<TC_ACT_NONE, false>(e, t_addr, out_off, st, n0, n_chunks, half, last_c, do_release, release);
        }
        return;
    }
    if (last_c < 0) if (do_release) release();

    for (int c = half; c < n_chunks; c += 2) {
        uint32_t acc[32];
        ptx::tmem_ld32(t_addr + c * 32, acc);
        ptx::tmem_ld_wait();
        if (c == last_c) if (do_release) release();
        if (orow < 0) continue;
        const int nb = n0 + c * 32;
        if (e.kind == TC_EPI_F16) {
            __half* o = reinterpret_cast<__half*>(e.out) + out_off + nb;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                float4 s0 = make_float4(1.f, 1.f, 1.f, 1.f), s1 = s0;
                float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f), h1 = h0;
                if (e.scale) {
                    s0 = __ldg(reinterpret_cast<const float4*>(e.scale + nb + j));
                    s1 = __ldg(reinterpret_cast<const float4*>(e.scale + nb + j + 4));
                }
                if (e.shift) {
                    h0 = __ldg(reinterpret_cast<const float4*>(e.shift + nb + j));
                    h1 = __ldg(reinterpret_cast<const float4*>(e.shift + nb + j + 4));
                }
                const float v0 = apply_act(fmaf(__uint_as_float(acc[j + 0]), s0.x, h0.x), e.act);
                const float v1 = apply_act(fmaf(__uint_as_float(acc[j + 1]), s0.y, h0.y), e.act);
                const float v2 = apply_act(fmaf(__uint_as_float(acc[j + 2]), s0.z, h0.z), e.act);
                const float v3 = apply_act(fmaf(__uint_as_float(acc[j + 3]), s0.w, h0.w), e.act);
                const float v4 = apply_act(fmaf(__uint_as_float(acc[j + 4]), s1.x, h1.x), e.act);
                const float v5 = apply_act(fmaf(__uint_as_float(acc[j + 5]), s1.y, h1.y), e.act);
                const float v6 = apply_act(fmaf(__uint_as_float(acc[j + 6]), s1.z, h1.z), e.act);
                const float v7 = apply_act(fmaf(__uint_as_float(acc[j + 7]), s1.w, h1.w), e.act);
                *reinterpret_cast<uint4*>(o + j) = make_uint4(pack_h2(v0, v1), pack_h2(v2, v3), pack_h2(v4, v5), pack_h2(v6, v7));
            }
        } else if (e.kind == TC_EPI_RES_F32) {
            float* o = reinterpret_cast<float*>(e.out) + out_off + nb;
            // all residual loads first: `res` may alias `out`, so the compiler cannot hoist them over the
            // stores itself and the chunk would pay eight dependent memory round trips instead of one
            float4 rv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                rv[j] = res_row ? *reinterpret_cast<const float4*>(res_row + nb + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 sv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (e.shift) sv = __ldg(reinterpret_cast<const float4*>(e.shift + nb + 4 * j));
                float4 ov;
                ov.x = __uint_as_float(acc[4 * j + 0]) + sv.x + rv[j].x;
                ov.y = __uint_as_float(acc[4 * j + 1]) + sv.y + rv[j].y;
                ov.z = __uint_as_float(acc[4 * j + 2]) + sv.z + rv[j].z;
                ov.w = __uint_as_float(acc[4 * j + 3]) + sv.w + rv[j].w;
                *reinterpret_cast<float4*>(o + 4 * j) = ov;
            }
        } else {  // TC_EPI_CONVT: 32 consecutive n share (dy,dx) because Cout % 32 == 0
            const int q = nb / e.ct_cout, co = nb - q * e.ct_cout;
            const int dy = q >> 1, dx = q & 1;
            uint4 pk[4];
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f), h1 = h0;
                if (e.shift) {
                    h0 = __ldg(reinterpret_cast<const float4*>(e.shift + co + j));
                    h1 = __ldg(reinterpret_cast<const float4*>(e.shift + co + j + 4));
                }
                pk[j >> 3] = make_uint4(pack_h2(__uint_as_float(acc[j + 0]) + h0.x, __uint_as_float(acc[j + 1]) + h0.y),
                                        pack_h2(__uint_as_float(acc[j + 2]) + h0.z, __uint_as_float(acc[j + 3]) + h0.w),
                                        pack_h2(__uint_as_float(acc[j + 4]) + h1.x, __uint_as_float(acc[j + 5]) + h1.y),
                                        pack_h2(__uint_as_float(acc[j + 6]) + h1.z, __uint_as_float(acc[j + 7]) + h1.w));
            }
            // A lane owns one input pixel: its four 16-byte pieces would go to four different 128-byte lines per store
            // instruction (32 lines per warp instruction). A 4x4 transpose inside each group of four lanes (two shuffle
            // rounds) makes lane j hold piece j of the group's four pixels, so every instruction writes 64 contiguous
            // bytes per lane group -- 8 lines instead of 32.
            const int l4 = (int)(threadIdx.x & 3);
            {
                const bool odd = l4 & 1;
                uint4 s0 = odd ? pk[0] : pk[1], s1 = odd ? pk[2] : pk[3];
                s0.x = __shfl_xor_sync(0xffffffffu, s0.x, 1); s0.y = __shfl_xor_sync(0xffffffffu, s0.y, 1);
                s0.z = __shfl_xor_sync(0xffffffffu, s0.z, 1); s0.w = __shfl_xor_sync(0xffffffffu, s0.w, 1);
                s1.x = __shfl_xor_sync(0xffffffffu, s1.x, 1); s1.y = __shfl_xor_sync(0xffffffffu, s1.y, 1);
                s1.z = __shfl_xor_sync(0xffffffffu, s1.z, 1); s1.w = __shfl_xor_sync(0xffffffffu, s1.w, 1);
                if (odd) { pk[0] = s0; pk[2] = s1; } else { pk[1] = s0; pk[3] = s1; }
                const bool hi = l4 & 2;
                uint4 t0 = hi ? pk[0] : pk[2], t1 = hi ? pk[1] : pk[3];
                t0.x = __shfl_xor_sync(0xffffffffu, t0.x, 2); t0.y = __shfl_xor_sync(0xffffffffu, t0.y, 2);
                t0.z = __shfl_xor_sync(0xffffffffu, t0.z, 2); t0.w = __shfl_xor_sync(0xffffffffu, t0.w, 2);
                t1.x = __shfl_xor_sync(0xffffffffu, t1.x, 2); t1.y = __shfl_xor_sync(0xffffffffu, t1.y, 2);
                t1.z = __shfl_xor_sync(0xffffffffu, t1.z, 2); t1.w = __shfl_xor_sync(0xffffffffu, t1.w, 2);
                if (hi) { pk[0] = t0; pk[1] = t1; } else { pk[2] = t0; pk[3] = t1; }
            }
            // pk[k] = piece l4 of the pixel of lane (lane & ~3) + k; the four pixels are consecutive in x (win % 4 == 0)
            const int x_base = ct_x - l4;
            const size_t opix = ((size_t)ct_b * (2 * e.ct_hin) + (2 * ct_y + dy)) * (size_t)(2 * e.ct_win) + (2 * x_base + dx);
            __half* o = reinterpret_cast<__half*>(e.out) + opix * (size_t)e.ldc + co + 8 * l4;
#pragma unroll
            for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(o + (size_t)(2 * k) * (size_t)e.ldc) = pk[k];
        }
    }
}

// PAIR: the two CTAs of a cluster (one TPC) run one 256 x block_n tile with tcgen05.mma.cta_group::2 -- each CTA
// stages its own 128 rows of A and HALF of the B tile, so the L2 -> SM operand traffic per FLOP drops by a third
// (the tile engine is bound by that feed, ~53 B/clk/SM, not by the tensor pipe). Only the leader CTA issues MMAs;
// TMA completions of both CTAs land on the leader's full barrier; tcgen05.commit multicasts to both CTAs.
template <bool PAIR>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
          const __grid_constant__ CUtensorMap tmB, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // provably warp-uniform
    const int stages = p.stages;
    const uint32_t rank = PAIR ? ptx::cluster_ctarank() : 0u;
    const int tile0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    constexpr int TILE_M = PAIR ? 2 * BLOCK_M : BLOCK_M;
    const int b_rows = PAIR ? p.block_n / 2 : p.block_n;  // B rows staged by this CTA
    const uint32_t b_stage_bytes = (uint32_t)b_rows * BLOCK_K * 2;
    const uint32_t smem_a = smem_base;
    const uint32_t smem_b = smem_a + stages * A_STAGE_BYTES;
    const uint32_t bar_base = smem_b + stages * b_stage_bytes;
    // barriers: full[s], empty[s], tmem_full[2], tmem_empty[2]
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * MAX_STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * MAX_STAGES + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_STAGES + 4);
    const uint32_t aux_off = (bar_base - smem_base) + 8u * (2 * MAX_STAGES + 4) + 16u;
    float* head_w_s = reinterpret_cast<float*>(smem_gen + aux_off);  // [8*64] + [8]
    // dynamic tile scheduler (common.cuh): warp 2 of the leader CTA claims tiles, the 11 (+10 in the peer CTA) other role
    // warps consume the same sequence; without a counter every role walks the static list tile0, tile0 + tile_step, ...
    ptx::SchedRing sched;
    sched.carve(smem_base + aux_off + HEAD_S_BYTES);
    const bool dyn = p.epi.sched_counter != nullptr;
    auto next_tile = [&](int k) -> int {
        if (dyn) return sched.consume(k, rank, warp == 0 && rank == 0);
        const int t = tile0 + k * tile_step;
        return t < p.n_tiles ? t : -1;
    };

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA0);
        ptx::prefetch_tmap(&tmB);
        if (p.conv && p.chunks_per_tap > p.chunks0) ptx::prefetch_tmap(&tmA1);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            ptx::mbar_init(full_bar(s), PAIR ? 4 : 2);  // one arrival per producer thread (A and B, per CTA)
            ptx::mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(tfull_bar(s), 1);
            ptx::mbar_init(tempty_bar(s), (PAIR ? 2 : 1) * (NUM_THREADS - 128));
        }
        sched.init(rank == 0 ? 11 : 10, PAIR);   // role warps of this CTA: 0, 3, 4..11 and, in the leader, the MMA warp
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        if (PAIR) { ptx::tmem_alloc_pair(tmem_slot, p.tmem_cols); ptx::tmem_relinquish_pair(); }
        else { ptx::tmem_alloc(tmem_slot, p.tmem_cols); ptx::tmem_relinquish(); }
    }
    if (p.epi.kind == TC_EPI_HEAD && warp >= 4) {
        head_smem_fill(head_w_s, p.epi, threadIdx.x - 128, NUM_THREADS - 128);
    }
    ptx::tc_fence_before();
    if (PAIR) ptx::cluster_sync(); else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

    const int num_kb = p.num_kb;

    if (warp == 0 || warp == 3) {
        // ===================================================== TMA producers: warp 0 streams A, warp 3 streams B.
        // Two issuing threads because one thread's wait + expect_tx + 2 x UTMALDG chain (~700 clk per k-block) is
        // slower than a k-block of MMA for narrow tiles (128 clk at N = 64).
        const bool is_a = warp == 0;
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t tx_bytes = (PAIR ? 2u : 1u) * (is_a ? (uint32_t)A_STAGE_BYTES : b_stage_bytes);
        for (int kt = 0;; ++kt) {
            const int tile = next_tile(kt);
            if (tile < 0) break;
            const int mt = tile / p.n_tiles_n, nt = tile - mt * p.n_tiles_n;
            const int m0 = mt * TILE_M + (int)rank * BLOCK_M, n0 = nt * p.block_n + (int)rank * b_rows;
            int img = 0, y0 = 0, x0 = 0;
            if (p.conv) {
                const int hw = p.H * p.W;
                img = m0 / hw;
                const int rem = m0 - img * hw;
                y0 = rem / p.W;
                x0 = rem - y0 * p.W;
            }
            int tap = 0, cc = 0;  // conv: k-block = (tap, 64-channel chunk), chunk fastest
            for (int kb = 0; kb < num_kb; ++kb) {
                ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
                const bool issue = ptx::elect_one();  // the whole warp walks the loop, one lane issues
                if (issue && (!PAIR || rank == 0)) ptx::mbar_expect_tx(full_bar(stage), tx_bytes);
                if (is_a) {
                    if (!p.conv && p.l2_prefetch > 0) {
                        // A rows are touched for the first time by all n-tiles of an m-block at once: without this the
                        // ring waits on a cold DRAM line fill (~3 us under load) for every k-block
                        int pk = kb + p.l2_prefetch, pm = m0;
                        if (pk >= num_kb) { pk -= num_kb; pm = -1; const int t2 = tile + tile_step;
                            if (t2 < p.n_tiles) pm = (t2 / p.n_tiles_n) * TILE_M + (int)rank * BLOCK_M; }
                        if (issue && pm >= 0 && pk < num_kb && (nt == 0 || pm != m0)) ptx::tma_prefetch_2d(&tmA0, pk * BLOCK_K, pm);
                    }
                    const uint32_t dst_a = smem_a + stage * A_STAGE_BYTES;
                    const CUtensorMap* ta = &tmA0;
                    int c0 = kb * BLOCK_K, dx = 0, dy = 0;
                    if (p.conv) {
                        dy = tap / 3 - 1;
                        dx = tap - (tap / 3) * 3 - 1;
                        if (cc < p.chunks0) c0 = cc * BLOCK_K;
                        else { ta = &tmA1; c0 = (cc - p.chunks0) * BLOCK_K; }
                        if (++cc == p.chunks_per_tap) { cc = 0; ++tap; }
                    }
                    if (issue) {
                        if (PAIR) {
                            if (!p.conv) ptx::tma_load_2d_pair(dst_a, ta, full_bar(stage), c0, m0);
                            else ptx::tma_load_4d_pair(dst_a, ta, full_bar(stage), c0, x0 + dx, y0 + dy, img);
                        } else {
                            if (!p.conv) ptx::tma_load_2d(dst_a, ta, full_bar(stage), c0, m0);
                            else ptx::tma_load_4d(dst_a, ta, full_bar(stage), c0, x0 + dx, y0 + dy, img);
                        }
                    }
                } else if (issue) {
                    if (PAIR) ptx::tma_load_2d_pair(smem_b + stage * b_stage_bytes, &tmB, full_bar(stage), kb * BLOCK_K, n0);
                    else ptx::tma_load_2d(smem_b + stage * b_stage_bytes, &tmB, full_bar(stage), kb * BLOCK_K, n0);
                }
                if (issue && PAIR && rank != 0) ptx::mbar_arrive_cluster(full_bar(stage), 0);
                if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ===================================================== MMA issuer (leader CTA only in PAIR mode)
        // instruction descriptor: D=f32, A=B=f16, both K-major, N>>3 @17, M>>4 @24
        const uint32_t idesc = (1u << 4) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
        // smem matrix descriptor: SWIZZLE_128B (2 @61), version 1 @46, SBO = 1024 B (8 rows x 128 B), LBO unused
        const uint64_t desc_hi = (2ull << 61) | (1ull << 46) | ((uint64_t)(1024 >> 4) << 32);
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0;; ++it) {
            if (next_tile(it) < 0) break;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            ptx::mbar_wait(tempty_bar(as), aphase ^ 1u);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.block_n);
            for (int kb = 0; kb < num_kb; ++kb) {
                ptx::mbar_wait(full_bar(stage), phase);
                ptx::tc_fence_after();
                const uint64_t a_desc = desc_hi | (uint64_t)(((smem_a + stage * A_STAGE_BYTES) >> 4) & 0x3FFF);
                const uint64_t b_desc = desc_hi | (uint64_t)(((smem_b + stage * b_stage_bytes) >> 4) & 0x3FFF);
                if (ptx::elect_one()) {  // always lane 0 of the full warp, so tcgen05.commit tracks this thread's MMAs
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 16; ++k) {  // 16 fp16 = 32 B -> +2 in the (addr >> 4) field
                        if (PAIR) ptx::umma_f16_pair(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        else ptx::umma_f16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    if (PAIR) {
                        ptx::umma_commit_pair(empty_bar(stage), 3);
                        if (kb == num_kb - 1) ptx::umma_commit_pair(tfull_bar(as), 3);
                    } else {
                        ptx::umma_commit(empty_bar(stage));
                        if (kb == num_kb - 1) ptx::umma_commit(tfull_bar(as));
                    }
                }
                __syncwarp();
                if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ===================================================== epilogue (TMEM -> registers -> global)
        // 8 warps: TMEM lane quadrant = warp & 3 (hardware restriction), and the two warps that share a quadrant
        // take alternating 32-column chunks, so a 128 x 256 tile drains in 4 chunk-steps per warp.
        const TcEpilogue& e = p.epi;
        const int quad = warp & 3, half = (warp - 4) >> 2;
        const int r = quad * 32 + lane;
        const int n_chunks = p.block_n / 32;  // block_n % 32 == 0 enforced on the host
        const int last_c = n_chunks - 1 - (((n_chunks - 1) & 1) != half ? 1 : 0);  // last chunk of this warp (may be < 0)
        auto release_tmem = [&](int as_) {
            ptx::tc_fence_before();
            if (PAIR) ptx::mbar_arrive_cluster(tempty_bar(as_), 0);
            else ptx::mbar_arrive(tempty_bar(as_));
        };
        for (int it = 0;; ++it) {
            const int tile = next_tile(it);
            if (tile < 0) break;
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int mt = tile / p.n_tiles_n, nt = tile - mt * p.n_tiles_n;
            const int m = mt * TILE_M + (int)rank * BLOCK_M + r, n0 = nt * p.block_n;
            const bool row_ok = m < p.M;
            ptx::mbar_wait(tfull_bar(as), aphase);
            ptx::tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * p.block_n);

            epilogue_tile(e, t_addr, m, row_ok, n0, n_chunks, half, last_c, head_w_s, true, [&]() { release_tmem(as); });
        }
    } else if (warp == 2 && dyn) {
        if (rank == 0) sched.produce<PAIR>(p.epi.sched_counter, p.n_tiles);
        else sched.relay();
    }

    ptx::tc_fence_before();
    if (PAIR) ptx::cluster_sync(); else __syncthreads();
    ptx::tc_fence_after();
    if (warp == 2) {
        if (PAIR) ptx::tmem_dealloc_pair(tmem_base, p.tmem_cols);
        else ptx::tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------- patch-resident 3x3 conv
// The k-block conv above re-reads every input pixel nine times through the L2 -> SM fabric, which is what bounds it
// (the wide, shallow layers at 512^2 / 1024^2 run at ~0.2 of the tensor peak). This variant stages, per 64-channel
// chunk, the (R+2) x 130-pixel input patch of an R-row x 128-pixel output tile ONCE (one 4-D TMA box, zero-filled
// halo) and forms the A operand of tap (dy,dx) and output row r by pointing the UMMA shared-memory descriptor at
// patch row (r+dy)*130 + dx: 128 consecutive 128-byte rows, the canonical K-major SWIZZLE_128B layout. The start
// address is then not 1024-byte aligned; measured on B200 (tools/probe_conv_patch.py): the 128B swizzle is a pure
// function of the absolute shared-memory address for both TMA and UMMA, so the descriptor needs NO base offset
// (setting (addr >> 7) & 7 there gives wrong results). Weights go through their own ring, one [N x 64] tile per
// (chunk, tap), shared by the R accumulators; when the whole weight set fits (Cin = 64, N = 64: 72 KB) it is
// loaded once per CTA and stays resident. Input bytes per output pixel drop 9x -> (R+2)/R * 130/128 = 2.03x.
constexpr int CP_R = 2;
constexpr int CP_PW = 130;                                 // patch width: 128 + 2 halo pixels
constexpr int CP_PATCH_BYTES = (CP_R + 2) * CP_PW * 128;   // 66,560 B = 65 KiB
constexpr int CP_MAX_BSTAGES = 12;

struct CpParams {
    int NB, H, W, N;
    int chunks0, chunks;  // 64-channel chunks of source 0 / of both sources
    int n_tiles, tiles_x, tiles_y;
    int b_stages, resident;  // resident: b_stages == 9 * chunks, every weight tile loaded once per CTA
    uint32_t tmem_cols;
    TcEpilogue epi;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_patch_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmB, const CpParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // provably warp-uniform
    const uint32_t b_stage_bytes = (uint32_t)p.N * BLOCK_K * 2;
    const uint32_t smem_patch = smem_base;
    const uint32_t smem_b = smem_patch + 2 * CP_PATCH_BYTES;
    const uint32_t bar_base = smem_b + p.b_stages * b_stage_bytes;
    auto pfull = [&](int s) { return bar_base + 8u * s; };
    auto pempty = [&](int s) { return bar_base + 8u * (2 + s); };
    auto bfull = [&](int s) { return bar_base + 8u * (4 + s); };
    auto bempty = [&](int s) { return bar_base + 8u * (4 + CP_MAX_BSTAGES + s); };
    auto tfull = [&](int s) { return bar_base + 8u * (4 + 2 * CP_MAX_BSTAGES + s); };
    auto tempty = [&](int s) { return bar_base + 8u * (6 + 2 * CP_MAX_BSTAGES + s); };
    const uint32_t tmem_slot = bar_base + 8u * (8 + 2 * CP_MAX_BSTAGES);
    const uint32_t aux_off = (tmem_slot - smem_base) + 16u;
    float* head_w_s = reinterpret_cast<float*>(smem_gen + aux_off);
    ptx::SchedRing sched;   // dynamic tile scheduler, as in tc_kernel (producer: warp 2)
    sched.carve(smem_base + aux_off + HEAD_S_BYTES);
    const bool dyn = p.epi.sched_counter != nullptr;
    auto next_tile = [&](int k) -> int {
        if (dyn) return sched.consume(k, 0, warp == 0);
        const int t = (int)blockIdx.x + k * (int)gridDim.x;
        return t < p.n_tiles ? t : -1;
    };

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmA0);
        ptx::prefetch_tmap(&tmB);
        if (p.chunks > p.chunks0) ptx::prefetch_tmap(&tmA1);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(pfull(s), 1);
            ptx::mbar_init(pempty(s), 1);
            ptx::mbar_init(tfull(s), 1);
            ptx::mbar_init(tempty(s), NUM_THREADS - 128);
        }
        for (int s = 0; s < p.b_stages; ++s) {
            ptx::mbar_init(bfull(s), 1);
            ptx::mbar_init(bempty(s), 1);
        }
        sched.init(11, false);
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_slot, p.tmem_cols);
        ptx::tmem_relinquish();
    }
    if (p.epi.kind == TC_EPI_HEAD && warp >= 4) {
        head_smem_fill(head_w_s, p.epi, threadIdx.x - 128, NUM_THREADS - 128);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

    auto tile_origin = [&](int tile, int& img, int& y0, int& x0) {
        const int xb = tile % p.tiles_x;
        const int t2 = tile / p.tiles_x;
        const int yb = t2 % p.tiles_y;
        img = t2 / p.tiles_y;
        y0 = yb * CP_R;
        x0 = xb * BLOCK_M;
    };

    if (warp == 0) {
        // ===================================================== patch producer (whole warp in the loop, one lane issues)
        int ps = 0;
        uint32_t pphase = 0;
        for (int kt = 0;; ++kt) {
            const int tile = next_tile(kt);
            if (tile < 0) break;
            int img, y0, x0;
            tile_origin(tile, img, y0, x0);
            for (int c = 0; c < p.chunks; ++c) {
                ptx::mbar_wait(pempty(ps), pphase ^ 1u);
                const bool first = c < p.chunks0;
                if (ptx::elect_one()) {
                    ptx::mbar_expect_tx(pfull(ps), CP_PATCH_BYTES);
                    // one TMA per patch row: a single large box is serviced at ~14 GB/s, concurrent boxes overlap
#pragma unroll
                    for (int pr = 0; pr < CP_R + 2; ++pr)
                        ptx::tma_load_4d(smem_patch + ps * CP_PATCH_BYTES + pr * (CP_PW * 128), first ? &tmA0 : &tmA1, pfull(ps),
                                         (first ? c : c - p.chunks0) * BLOCK_K, x0 - 1, y0 - 1 + pr, img);
                }
                if (++ps == 2) { ps = 0; pphase ^= 1u; }
            }
        }
    } else if (warp == 3) {
        // ===================================================== weight producer: one [N x 64] tile per (chunk, tap)
        int bs = 0;
        uint32_t bphase = 0;
        for (int kt = 0;; ++kt) {
            if (next_tile(kt) < 0) break;
            if (p.resident && kt > 0) continue;  // resident weights: one pass fills every slot for good (the tile sequence is still consumed)
            for (int c = 0; c < p.chunks; ++c) {
                for (int tap = 0; tap < 9; ++tap) {
                    ptx::mbar_wait(bempty(bs), bphase ^ 1u);
                    if (ptx::elect_one()) {
                        ptx::mbar_expect_tx(bfull(bs), b_stage_bytes);
                        ptx::tma_load_2d(smem_b + bs * b_stage_bytes, &tmB, bfull(bs), (tap * p.chunks + c) * BLOCK_K, 0);
                    }
                    if (++bs == p.b_stages) { bs = 0; bphase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer (whole warp in the loop, one lane issues)
        const uint32_t idesc = (1u << 4) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
        const uint64_t desc_hi = (2ull << 61) | (1ull << 46) | ((uint64_t)(1024 >> 4) << 32);
        int ps = 0, bs = 0;
        uint32_t pphase = 0, bphase = 0;
        for (int it = 0;; ++it) {
            if (next_tile(it) < 0) break;
            const int as = it & 1;
            ptx::mbar_wait(tempty(as), ((it >> 1) & 1) ^ 1u);
            ptx::tc_fence_after();
            for (int c = 0; c < p.chunks; ++c) {
                ptx::mbar_wait(pfull(ps), pphase);
                ptx::tc_fence_after();
                const uint32_t patch = smem_patch + ps * CP_PATCH_BYTES;
                for (int tap = 0; tap < 9; ++tap) {
                    ptx::mbar_wait(bfull(bs), p.resident ? 0u : bphase);  // resident: phase 0 completes once, stays complete
                    ptx::tc_fence_after();
                    const int dy = tap / 3, dx = tap - dy * 3;
                    const uint64_t b_desc = desc_hi | (uint64_t)(((smem_b + bs * b_stage_bytes) >> 4) & 0x3FFF);
                    if (ptx::elect_one()) {
#pragma unroll
                        for (int r = 0; r < CP_R; ++r) {
                            const uint32_t a_addr = patch + (uint32_t)((r + dy) * CP_PW + dx) * 128u;
                            const uint64_t a_desc = desc_hi | (uint64_t)((a_addr >> 4) & 0x3FFF);
                            const uint32_t d_tmem = tmem_base + (uint32_t)((as * CP_R + r) * p.N);
#pragma unroll
                            for (int k = 0; k < BLOCK_K / 16; ++k)
                                ptx::umma_f16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (c | tap | k) != 0 ? 1u : 0u);
                        }
                        if (!p.resident) ptx::umma_commit(bempty(bs));
                        if (tap == 8) ptx::umma_commit(pempty(ps));
                        if (tap == 8 && c == p.chunks - 1) ptx::umma_commit(tfull(as));
                    }
                    __syncwarp();
                    if (++bs == p.b_stages) { bs = 0; bphase ^= 1u; }
                }
                if (++ps == 2) { ps = 0; pphase ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // ===================================================== epilogue: R accumulators of 128 pixels each
        const TcEpilogue& e = p.epi;
        const int quad = warp & 3, half = (warp - 4) >> 2;
        const int n_chunks = p.N / 32;
        const int last_c = n_chunks - 1 - (((n_chunks - 1) & 1) != half ? 1 : 0);
        for (int it = 0;; ++it) {
            const int tile = next_tile(it);
            if (tile < 0) break;
            const int as = it & 1;
            int img, y0, x0;
            tile_origin(tile, img, y0, x0);
            ptx::mbar_wait(tfull(as), (it >> 1) & 1);
            ptx::tc_fence_after();
#pragma unroll
            for (int r = 0; r < CP_R; ++r) {
                const int m = (img * p.H + y0 + r) * p.W + x0 + quad * 32 + lane;
                const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((as * CP_R + r) * p.N);
                epilogue_tile(e, t_addr, m, true, 0, n_chunks, half, last_c, head_w_s, r == CP_R - 1, [&]() {
                    ptx::tc_fence_before();
                    ptx::mbar_arrive(tempty(as));
                });
            }
        }
    } else if (warp == 2 && dyn) {
        sched.produce<false>(p.epi.sched_counter, p.n_tiles);
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------- host side
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
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

int make_tmap(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box) {
    auto fn = get_encode_fn();
    CVB_CHECK(fn != nullptr, CVB_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    uint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CVB_CHECK(r == CUDA_SUCCESS, CVB_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, base %p)", (int)r,
              rank, base);
    return CVB_OK;
}

int check_epilogue(const TcEpilogue& e, int N, int block_n) {
    CVB_CHECK(block_n >= 32 && block_n <= 256 && block_n % 32 == 0 && N % block_n == 0, CVB_ESHAPE,
              "tc: block_n %d must be a multiple of 32 in [32,256] dividing N=%d", block_n, N);
    if (e.kind == TC_EPI_F16 || e.kind == TC_EPI_RES_F32)
        CVB_CHECK(e.out != nullptr && e.ldc % 8 == 0, CVB_EARG, "tc: epilogue needs out and ldc %% 8 == 0");
    if (e.kind == TC_EPI_CONVT)
        CVB_CHECK(e.out != nullptr && e.ct_cout % 32 == 0 && N == 4 * e.ct_cout && e.ldc % 8 == 0 && e.ct_win % 4 == 0, CVB_ESHAPE,
                  "tc: CONVT needs Cout %% 32 == 0, N == 4*Cout and an input width that is a multiple of 4");
    if (e.kind == TC_EPI_HEAD)
        CVB_CHECK(N == 64 && block_n == 64 && e.head_nc >= 1 && e.head_nc <= 8 && e.head_out && e.head_w && e.head_b &&
                      e.scale && e.shift,
                  CVB_ESHAPE, "tc: HEAD epilogue needs N == 64, 1..8 classes, scale and shift");
    if (e.row_map == TC_ROW_WINDOW || e.row_map == TC_ROW_TO_WINDOW)
        CVB_CHECK(e.win_size > 0 && e.win_grid > 0 && e.tok_h > 0 && e.tok_w > 0, CVB_EARG, "tc: bad window map");
    if (e.row_map == TC_ROW_SEQ) CVB_CHECK(e.row_seq > 0, CVB_EARG, "tc: bad seq map");
    return CVB_OK;
}

// Optional per-launch timing of the tile engine (bench.py roofline leg): event pairs around every launch on the
// launching stream; read back with cvb_tc_profile_end.
struct TcProfile {
    bool on = false;
    std::vector<cudaEvent_t> ev;
    size_t used = 0;
    double flops = 0.0;
} g_prof;

bool g_pair_enabled = true;
int g_max_ctas = 1 << 30;  // debug knob (cvb_tc_set_max_ctas): restrict the persistent grid, for feed-bandwidth probes
int g_l2_prefetch = 0;  // measured on B200: no effect on the encoder GEMMs (the A stream is not cold-miss bound); kept as a hook

// pair mode needs at least two 256-row tiles per pair-CTA to pay off and an even B split in 16-row units
bool use_pair(int M, int N, int block_n) {
    if (!g_pair_enabled || block_n % 32 != 0) return false;
    const long long pair_tiles = (long long)((M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * (N / block_n);
    return pair_tiles >= cvb_num_sms() / 2;
}

int launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, TcParams& p, bool pair, cudaStream_t stream) {
    const int b_stage = (pair ? p.block_n / 2 : p.block_n) * BLOCK_K * 2;
    int stages = SMEM_BUDGET / (A_STAGE_BYTES + b_stage);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    p.stages = stages;
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * p.block_n)) cols <<= 1;
    p.tmem_cols = cols;
    // 1024 B alignment slack + stages + barriers/tmem slot + head weights; always > half an SM so that
    // one CTA (and one 512-column TMEM allocation) lives on an SM at a time.
    size_t smem = 1024 + (size_t)stages * (A_STAGE_BYTES + b_stage) + 8 * (2 * MAX_STAGES + 4) + 16 + HEAD_S_BYTES + ptx::SCHED_BYTES;
    if (smem < 120 * 1024) smem = 120 * 1024;
    static unsigned long long configured = 0;  // one bit per device: function attributes are per device
    const int cfg_dev = cvb_current_device();
    if (!((configured >> cfg_dev) & 1ull)) {
        CVB_CUDA(cudaFuncSetAttribute(tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        CVB_CUDA(cudaFuncSetAttribute(tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        configured |= 1ull << cfg_dev;
    }
    p.n_tiles_n = p.N / p.block_n;
    p.n_tiles = cdiv(p.M, pair ? 2 * BLOCK_M : BLOCK_M) * p.n_tiles_n;
    const int sms = cvb_num_sms() < g_max_ctas ? cvb_num_sms() : g_max_ctas;
    int grid = pair ? 2 * (p.n_tiles < sms / 2 ? p.n_tiles : sms / 2) : (p.n_tiles < sms ? p.n_tiles : sms);
    const bool prof = g_prof.on && g_prof.used + 2 <= g_prof.ev.size();
    if (prof) CVB_CUDA(cudaEventRecord(g_prof.ev[g_prof.used], stream));
    if (pair) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(NUM_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CVB_CUDA(cudaLaunchKernelEx(&cfg, tc_kernel<true>, a0, a1, b, p));
    } else {
        tc_kernel<false><<<grid, NUM_THREADS, smem, stream>>>(a0, a1, b, p);
    }
    if (prof) {
        CVB_CUDA(cudaEventRecord(g_prof.ev[g_prof.used + 1], stream));
        g_prof.used += 2;
        g_prof.flops += 2.0 * (double)p.M * (double)p.N * (double)p.K;
    }
    CVB_CUDA(cudaGetLastError());
    cvb_note_launches(1);
    return CVB_OK;
}

}  // namespace

int tc_pick_block_n(int n) {
    for (int bn = 256; bn >= 32; bn -= 32)
        if (n % bn == 0) return bn;
    return 0;
}

int tc_gemm(const __half* A, int M, int K, long long lda, const __half* W, int N, long long ldw, int block_n,
            const TcEpilogue& epi, cudaStream_t stream) {
    CVB_CHECK(A && W && M > 0 && N > 0 && K > 0, CVB_EARG, "tc_gemm: null operand or empty shape");
    CVB_CHECK(K % BLOCK_K == 0 && lda % 8 == 0 && ldw % 8 == 0, CVB_ESHAPE,
              "tc_gemm: K=%d must be a multiple of 64 and lda/ldw multiples of 8", K);
    CVB_CHECK(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0, CVB_EARG, "tc_gemm: operands must be 16-byte aligned");
    CVB_TRY(check_epilogue(epi, N, block_n));
    if (epi.kind == TC_EPI_CONVT)  // the scatter epilogue transposes inside lane groups: no partially valid warps
        CVB_CHECK(M % BLOCK_M == 0 && M == (M / (epi.ct_hin * epi.ct_win)) * epi.ct_hin * epi.ct_win, CVB_ESHAPE,
                  "tc_gemm: CONVT needs M = NB*hin*win to be a multiple of %d (M=%d)", BLOCK_M, M);
    const bool pair = use_pair(M, N, block_n) && epi.kind != TC_EPI_HEAD;
    CUtensorMap ta, tb;
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
        uint64_t str[1] = {(uint64_t)lda * 2};
        uint32_t box[2] = {BLOCK_K, BLOCK_M};
        CVB_TRY(make_tmap(&ta, A, 2, dims, str, box));
    }
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
        uint64_t str[1] = {(uint64_t)ldw * 2};
        uint32_t box[2] = {BLOCK_K, (uint32_t)(pair ? block_n / 2 : block_n)};
        CVB_TRY(make_tmap(&tb, W, 2, dims, str, box));
    }
    TcParams p{};
    p.M = M; p.N = N; p.K = K; p.block_n = block_n; p.num_kb = K / BLOCK_K; p.conv = 0; p.epi = epi;
    p.l2_prefetch = g_l2_prefetch;
    return launch(ta, ta, tb, p, pair, stream);
}

static int g_patch_mode = 1;  // 0: k-block conv only, 1: patch-resident conv where the shape allows it

static int conv_patch_launch(const __half* src0, int C0, const __half* src1, int C1, int NB, int H, int W, const __half* Wp, int N,
                             const TcEpilogue& epi, cudaStream_t stream) {
    CUtensorMap t0, t1, tb;
    auto mk = [&](CUtensorMap* tm, const __half* src, int C) {
        uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
        uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
        uint32_t box[4] = {BLOCK_K, (uint32_t)CP_PW, 1, 1};
        return make_tmap(tm, src, 4, dims, str, box);
    };
    CVB_TRY(mk(&t0, src0, C0));
    if (C1 > 0) CVB_TRY(mk(&t1, src1, C1)); else t1 = t0;
    const int K = 9 * (C0 + C1);
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
        uint64_t str[1] = {(uint64_t)K * 2};
        uint32_t box[2] = {BLOCK_K, (uint32_t)N};
        CVB_TRY(make_tmap(&tb, Wp, 2, dims, str, box));
    }
    CpParams p{};
    p.NB = NB; p.H = H; p.W = W; p.N = N; p.chunks0 = C0 / 64; p.chunks = (C0 + C1) / 64;
    p.tiles_x = W / BLOCK_M; p.tiles_y = H / CP_R; p.n_tiles = NB * p.tiles_x * p.tiles_y;
    p.epi = epi;
    const int b_stage = N * BLOCK_K * 2;
    const int fixed = 1024 + 2 * CP_PATCH_BYTES + 8 * (8 + 2 * CP_MAX_BSTAGES) + 16 + HEAD_S_BYTES + ptx::SCHED_BYTES;
    int bst = (226 * 1024 - fixed) / b_stage;
    if (bst > CP_MAX_BSTAGES) bst = CP_MAX_BSTAGES;
    CVB_CHECK(bst >= 3, CVB_ESHAPE, "conv_patch: not enough shared memory for the weight ring (N=%d)", N);
    p.resident = 9 * p.chunks <= bst ? 1 : 0;
    if (p.resident) bst = 9 * p.chunks;
    p.b_stages = bst;
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * CP_R * N)) cols <<= 1;
    p.tmem_cols = cols;
    const size_t smem = (size_t)fixed + (size_t)bst * b_stage;
    static unsigned long long configured = 0;  // one bit per device: function attributes are per device
    const int cfg_dev = cvb_current_device();
    if (!((configured >> cfg_dev) & 1ull)) {
        CVB_CUDA(cudaFuncSetAttribute(conv_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        configured |= 1ull << cfg_dev;
    }
    const int grid = p.n_tiles < cvb_num_sms() ? p.n_tiles : cvb_num_sms();
    const bool prof = g_prof.on && g_prof.used + 2 <= g_prof.ev.size();
    if (prof) CVB_CUDA(cudaEventRecord(g_prof.ev[g_prof.used], stream));
    conv_patch_kernel<<<grid, NUM_THREADS, smem, stream>>>(t0, t1, tb, p);
    if (prof) {
        CVB_CUDA(cudaEventRecord(g_prof.ev[g_prof.used + 1], stream));
        g_prof.used += 2;
        g_prof.flops += 2.0 * (double)NB * H * W * (double)N * (double)K;
    }
    CVB_CUDA(cudaGetLastError());
    cvb_note_launches(1);
    return CVB_OK;
}

int tc_conv3x3(const __half* src0, int C0, const __half* src1, int C1, int NB, int H, int W, const __half* Wp,
               int N, int block_n, const TcEpilogue& epi, cudaStream_t stream) {
    CVB_CHECK(src0 && Wp && NB > 0 && H > 0 && W > 0, CVB_EARG, "tc_conv3x3: null operand or empty shape");
    if (!src1) C1 = 0;
    CVB_CHECK(C0 > 0 && C0 % 64 == 0 && C1 % 64 == 0, CVB_ESHAPE, "tc_conv3x3: channels (%d,%d) must be multiples of 64", C0, C1);
    if (g_patch_mode != 0 && W % BLOCK_M == 0 && H % CP_R == 0 && (N == 64 || N == 128) && block_n == N &&
        (epi.kind == TC_EPI_F16 || epi.kind == TC_EPI_HEAD) && epi.row_map == TC_ROW_IDENTITY) {
        CVB_TRY(check_epilogue(epi, N, block_n));
        return conv_patch_launch(src0, C0, src1, C1, NB, H, W, Wp, N, epi, stream);
    }
    const int TW = W < BLOCK_M ? W : BLOCK_M;
    CVB_CHECK(BLOCK_M % TW == 0 && W % TW == 0 && H % (BLOCK_M / TW) == 0, CVB_ESHAPE,
              "tc_conv3x3: image %dx%d does not tile into 128-pixel blocks", H, W);
    CVB_TRY(check_epilogue(epi, N, block_n));
    const int TH = BLOCK_M / TW;
    CUtensorMap t0, t1, tb;
    auto mk = [&](CUtensorMap* tm, const __half* src, int C) {
        uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
        uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
        uint32_t box[4] = {BLOCK_K, (uint32_t)TW, (uint32_t)TH, 1};
        return make_tmap(tm, src, 4, dims, str, box);
    };
    CVB_TRY(mk(&t0, src0, C0));
    if (C1 > 0) CVB_TRY(mk(&t1, src1, C1)); else t1 = t0;
    const int K = 9 * (C0 + C1);
    const bool pair = use_pair(NB * H * W, N, block_n) && (NB * H * W) % (2 * BLOCK_M) == 0;
    {
        uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
        uint64_t str[1] = {(uint64_t)K * 2};
        uint32_t box[2] = {BLOCK_K, (uint32_t)(pair ? block_n / 2 : block_n)};
        CVB_TRY(make_tmap(&tb, Wp, 2, dims, str, box));
    }
    TcParams p{};
    p.M = NB * H * W; p.N = N; p.K = K; p.block_n = block_n; p.num_kb = K / BLOCK_K; p.conv = 1;
    p.chunks0 = C0 / 64; p.chunks_per_tap = (C0 + C1) / 64; p.H = H; p.W = W; p.TW = TW; p.epi = epi;
    return launch(t0, t1, tb, p, pair, stream);
}

extern "C" __attribute__((visibility("default"))) int cvb_tc_profile_begin(int max_launches) {
    CVB_CHECK(max_launches > 0, CVB_EARG, "cvb_tc_profile_begin: max_launches must be positive");
    while (g_prof.ev.size() < (size_t)max_launches * 2) {
        cudaEvent_t e;
        CVB_CUDA(cudaEventCreate(&e));
        g_prof.ev.push_back(e);
    }
    g_prof.used = 0;
    g_prof.flops = 0.0;
    g_prof.on = true;
    return CVB_OK;
}

// Synchronises the recorded events. total_ms = sum of per-launch durations, flops = executed 2*M*N*K summed.
extern "C" __attribute__((visibility("default"))) int cvb_tc_profile_end(double* total_ms, int* n_launches, double* flops) {
    g_prof.on = false;
    double ms = 0.0;
    for (size_t i = 0; i + 1 < g_prof.used; i += 2) {
        CVB_CUDA(cudaEventSynchronize(g_prof.ev[i + 1]));
        float t = 0.f;
        CVB_CUDA(cudaEventElapsedTime(&t, g_prof.ev[i], g_prof.ev[i + 1]));
        ms += t;
    }
    if (total_ms) *total_ms = ms;
    if (n_launches) *n_launches = (int)(g_prof.used / 2);
    if (flops) *flops = g_prof.flops;
    return CVB_OK;
}

// Test / ablation hook: 0 disables the CTA-pair (cta_group::2) path, 1 enables it (default).
extern "C" __attribute__((visibility("default"))) void cvb_tc_set_max_ctas(int n) { g_max_ctas = n > 0 ? n : (1 << 30); }
extern "C" __attribute__((visibility("default"))) void cvb_tc_set_pair_mode(int on) { g_pair_enabled = on != 0; }

// Test / ablation hook: 0 = k-block conv only, 1 = patch-resident conv where the shape allows it (default).
extern "C" __attribute__((visibility("default"))) void cvb_tc_set_conv_patch_mode(int mode) { g_patch_mode = mode; }

// Test / ablation hook: number of A k-blocks prefetched into L2 ahead of the ring in GEMM mode (0 disables).
extern "C" __attribute__((visibility("default"))) void cvb_tc_set_l2_prefetch(int k) { g_l2_prefetch = k < 0 ? 0 : k; }
