// attention.cu -- multi-head self-attention of the ViT encoders (windowed and global, SAM and ViT-S).
//
//  * relpos_kernel: the decomposed relative-position terms of SAM (image_encoder.py:354-392). They depend on the
//    UNSCALED query, so they are two small per-query tables  rel_h[q, kh], rel_w[q, kw]  computed once per block
//    with CUDA cores (2 * (gh + gw) * hd MACs per query -- 3 % of the attention FLOPs).
//  * flash_kernel<HD>: S = scale * q k^T + rel_h[q, k / gw] + rel_w[q, k % gw]; online softmax; O = P v.
//    Scores never leave the SM (the reference materialises a 4.3 GB fp32 score tensor at B=4, image_encoder.py:244-251).
//    Tensor-core path: mma.sync m16n8k16 fp16 -> fp32 with ldmatrix-fed fragments, cp.async double-buffered K/V.
//    TODO(round 2): move QK^T / PV to tcgen05 with S and O in TMEM; attention is 6 % of the forward FLOPs.
#include "ops.h"

namespace {

constexpr int RP_THREADS = 128;
constexpr int RP_MAXQ = 256;

__global__ void __launch_bounds__(RP_THREADS)
relpos_kernel(const __half* __restrict__ qkv, int heads, int hd, int gh, int gw, int rows_per_cta,
              const float* __restrict__ Rh, const float* __restrict__ Rw, float* __restrict__ rel_h, float* __restrict__ rel_w) {
    extern __shared__ float rp_smem[];
    const int ld = hd + 1;
    float* q_s = rp_smem;                               // [nq][ld]
    const int nq = rows_per_cta * gw;
    float* th_s = q_s + nq * ld;                        // [2gh-1][ld]
    float* tw_s = th_s + (2 * gh - 1) * ld;             // [2gw-1][ld]
    const int g = blockIdx.y, bp = g / heads, head = g - bp * heads;
    const int S = gh * gw, D3 = 3 * heads * hd;
    const int qh0 = blockIdx.x * rows_per_cta;
    const int t0 = qh0 * gw;
    for (int i = threadIdx.x; i < nq * hd; i += RP_THREADS) {
        const int ql = i / hd, c = i - ql * hd;
        q_s[ql * ld + c] = __half2float(qkv[((long long)bp * S + t0 + ql) * D3 + head * hd + c]);
    }
    for (int i = threadIdx.x; i < (2 * gh - 1) * hd; i += RP_THREADS) th_s[(i / hd) * ld + i % hd] = Rh[i];
    for (int i = threadIdx.x; i < (2 * gw - 1) * hd; i += RP_THREADS) tw_s[(i / hd) * ld + i % hd] = Rw[i];
    __syncthreads();
    const int per_q = gh + gw;
    for (int o = threadIdx.x; o < nq * per_q; o += RP_THREADS) {
        const int ql = o / per_q, k = o - ql * per_q;
        const int qh = qh0 + ql / gw, qw = ql % gw;
        const float* qv = q_s + ql * ld;
        const float* tv = k < gh ? th_s + (qh - k + gh - 1) * ld : tw_s + (qw - (k - gh) + gw - 1) * ld;
        float acc = 0.f;
        for (int c = 0; c < hd; ++c) acc = fmaf(qv[c], tv[c], acc);
        const long long row = (long long)g * S + t0 + ql;
        if (k < gh) rel_h[row * gh + k] = acc;
        else rel_w[row * gw + (k - gh)] = acc;
    }
}

// ------------------------------------------------------------------------------------------ flash attention
constexpr int FA_BQ = 64, FA_BK = 64, FA_THREADS = 128;

template <int HD>
struct FaSmem {
    static constexpr int LD = HD + 8;  // halves per smem row: (HD+8)*2 B is an odd multiple of 16 B -> conflict-free ldmatrix
    __half q[FA_BQ * LD];
    __half k[2][FA_BK * LD];
    __half v[2][FA_BK * LD];
    float rel_h[FA_BQ * 64];
    float rel_w[FA_BQ * 64];
};

template <int HD>
__device__ __forceinline__ void fa_load_tile(__half* dst, const __half* src_base, long long row_stride, int row0, int n_rows,
                                             int tid) {
    constexpr int LD = FaSmem<HD>::LD;
    constexpr int CH = HD / 8;  // 16-byte chunks per row
    for (int i = tid; i < FA_BK * CH; i += FA_THREADS) {
        const int r = i / CH, c = i - r * CH;
        const int row = row0 + r;
        const bool ok = row < n_rows;
        const __half* src = src_base + (long long)(ok ? row : 0) * row_stride + c * 8;
        ptx::cp_async16(ptx::smem_u32(dst + r * LD + c * 8), src, ok);
    }
}

template <int HD>
__global__ void __launch_bounds__(FA_THREADS)
flash_kernel(const __half* __restrict__ qkv, int S, int heads, float scale, const float* __restrict__ rel_h,
             const float* __restrict__ rel_w, int gh, int gw, __half* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t fa_smem_raw[];
    FaSmem<HD>& sm = *reinterpret_cast<FaSmem<HD>*>(fa_smem_raw);
    constexpr int LD = FaSmem<HD>::LD;
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
    const bool has_bias = rel_h != nullptr;
    const int n_kt = (S + FA_BK - 1) / FA_BK;

    fa_load_tile<HD>(sm.q, q_base, row_stride, q0, S, tid);
    fa_load_tile<HD>(sm.k[0], k_base, row_stride, 0, S, tid);
    fa_load_tile<HD>(sm.v[0], v_base, row_stride, 0, S, tid);
    ptx::cp_async_commit();
    if (has_bias) {
        for (int i = tid; i < FA_BQ * gh; i += FA_THREADS) {
            const int ql = i / gh, kk = i - ql * gh;
            const int qrow = min(q0 + ql, S - 1);
            sm.rel_h[ql * 64 + kk] = rel_h[((long long)g * S + qrow) * gh + kk];
        }
        for (int i = tid; i < FA_BQ * gw; i += FA_THREADS) {
            const int ql = i / gw, kk = i - ql * gw;
            const int qrow = min(q0 + ql, S - 1);
            sm.rel_w[ql * 64 + kk] = rel_w[((long long)g * S + qrow) * gw + kk];
        }
    }

    float o_acc[NT_O][4];
#pragma unroll
    for (int i = 0; i < NT_O; ++i) { o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    uint32_t q_frag[KSTEPS][4];
    const int r_lo = warp * 16 + (lane >> 2);  // local query row of c0/c1; c2/c3 are r_lo + 8
    const float sl2 = scale * 1.4426950408889634f;
    constexpr float L2E = 1.4426950408889634f;

    for (int kt = 0; kt < n_kt; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < n_kt) {
            fa_load_tile<HD>(sm.k[buf ^ 1], k_base, row_stride, (kt + 1) * FA_BK, S, tid);
            fa_load_tile<HD>(sm.v[buf ^ 1], v_base, row_stride, (kt + 1) * FA_BK, S, tid);
            ptx::cp_async_commit();
            ptx::cp_async_wait<1>();
        } else {
            ptx::cp_async_wait<0>();
        }
        __syncthreads();
        if (kt == 0) {
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int col = ks * 16 + (lane >> 4) * 8;
                ptx::ldmatrix_x4(ptx::smem_u32(sm.q + row * LD + col), q_frag[ks][0], q_frag[ks][1], q_frag[ks][2], q_frag[ks][3]);
            }
        }
        // ---- S = Q K^T (16 x 64 per warp)
        float s_acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { s_acc[i][0] = s_acc[i][1] = s_acc[i][2] = s_acc[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {  // pairs of 8-key n-tiles
                uint32_t b0, b1, b2, b3;
                const int row = np * 16 + (lane & 7) + ((lane >> 4) << 3);
                const int col = ks * 16 + ((lane >> 3) & 1) * 8;
                ptx::ldmatrix_x4(ptx::smem_u32(sm.k[buf] + row * LD + col), b0, b1, b2, b3);
                ptx::mma_16816(s_acc[2 * np], q_frag[ks], b0, b1);
                ptx::mma_16816(s_acc[2 * np + 1], q_frag[ks], b2, b3);
            }
        }
        // ---- scale, bias, mask (log2 domain)
        const int kbase = kt * FA_BK;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int kcol = kbase + nt * 8 + 2 * (lane & 3) + e;
                float b_lo = 0.f, b_hi = 0.f;
                if (has_bias) {
                    const int kh = kcol / gw, kw = kcol - kh * gw;
                    if (kcol < S) {
                        b_lo = sm.rel_h[r_lo * 64 + kh] + sm.rel_w[r_lo * 64 + kw];
                        b_hi = sm.rel_h[(r_lo + 8) * 64 + kh] + sm.rel_w[(r_lo + 8) * 64 + kw];
                    }
                }
                const bool valid = kcol < S;
                s_acc[nt][e] = valid ? fmaf(s_acc[nt][e], sl2, b_lo * L2E) : -INFINITY;
                s_acc[nt][2 + e] = valid ? fmaf(s_acc[nt][2 + e], sl2, b_hi * L2E) : -INFINITY;
            }
        }
        // ---- online softmax
        float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            mx[0] = fmaxf(mx[0], fmaxf(s_acc[nt][0], s_acc[nt][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s_acc[nt][2], s_acc[nt][3]));
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
        }
        float alpha[2], rs[2] = {0.f, 0.f};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            alpha[h] = exp2f(m_run[h] - mx[h]);
            m_run[h] = mx[h];
        }
        uint32_t p_frag[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float p0 = exp2f(s_acc[nt][0] - mx[0]), p1 = exp2f(s_acc[nt][1] - mx[0]);
            const float p2 = exp2f(s_acc[nt][2] - mx[1]), p3 = exp2f(s_acc[nt][3] - mx[1]);
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
        __syncthreads();
    }
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

template <int HD>
int launch_flash(const __half* qkv, int Gb, int S, int heads, float scale, const float* rel_h, const float* rel_w,
                 int gh, int gw, __half* out, cudaStream_t stream) {
    static bool configured = false;
    const int smem = (int)sizeof(FaSmem<HD>);
    if (!configured) {
        CVB_CUDA(cudaFuncSetAttribute(flash_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    dim3 grid(cdiv(S, FA_BQ), Gb * heads);
    flash_kernel<HD><<<grid, FA_THREADS, smem, stream>>>(qkv, S, heads, scale, rel_h, rel_w, gh, gw, out);
    cvb_note_launches(1);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}

}  // namespace

int op_relpos(const __half* qkv, int Gb, int heads, int hd, int gh, int gw, const float* Rh, const float* Rw,
              float* rel_h, float* rel_w, cudaStream_t stream) {
    CVB_CHECK(qkv && Rh && Rw && rel_h && rel_w, CVB_EARG, "relpos: null operand");
    CVB_CHECK(gh <= 64 && gw <= 64, CVB_ESHAPE, "relpos: token grid %dx%d exceeds 64x64", gh, gw);
    const int rows_per_cta = gw >= 32 ? 1 : (gh * gw <= RP_MAXQ ? gh : 1);
    const int nq = rows_per_cta * gw;
    const size_t smem = (size_t)(nq + 2 * gh - 1 + 2 * gw - 1) * (hd + 1) * sizeof(float);
    static size_t configured = 0;
    if (smem > configured) {
        CVB_CUDA(cudaFuncSetAttribute(relpos_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = 200 * 1024;
    }
    CVB_CHECK(smem <= 200 * 1024, CVB_ESHAPE, "relpos: shared memory %zu too large", smem);
    dim3 grid(gh / rows_per_cta, Gb * heads);
    relpos_kernel<<<grid, RP_THREADS, smem, stream>>>(qkv, heads, hd, gh, gw, rows_per_cta, Rh, Rw, rel_h, rel_w);
    cvb_note_launches(1);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}

int op_attention(const __half* qkv, int Gb, int S, int heads, int hd, float scale, const float* rel_h,
                 const float* rel_w, int gh, int gw, __half* out, cudaStream_t stream) {
    CVB_CHECK(qkv && out && Gb > 0 && S > 0, CVB_EARG, "attention: null operand or empty shape");
    CVB_CHECK((rel_h == nullptr) == (rel_w == nullptr), CVB_EARG, "attention: rel_h and rel_w must both be set or both null");
    if (rel_h) CVB_CHECK(gh * gw == S && gh <= 64 && gw <= 64, CVB_ESHAPE, "attention: bias grid %dx%d does not match S=%d", gh, gw, S);
    if (!rel_h) { gh = 1; gw = S; }
    if (hd == 80) return launch_flash<80>(qkv, Gb, S, heads, scale, rel_h, rel_w, gh, gw, out, stream);
    if (hd == 64) return launch_flash<64>(qkv, Gb, S, heads, scale, rel_h, rel_w, gh, gw, out, stream);
    cvb_set_error("attention: head dim %d not supported (64 or 80)", hd);
    return CVB_ESHAPE;
}
