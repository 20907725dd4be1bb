// ops_misc.cu -- HBM-bound helper kernels of the forward: patch im2col, LayerNorm (with window partition),
// casts / layout transforms, the 3->32 stem stencil and the two tiny tissue-classifier tails.
#include "ops.h"

namespace {

// ------------------------------------------------------------------------------------------ im2col
__global__ void patch_im2col_kernel(const float* __restrict__ x, int B, int H, int W, int P, __half* __restrict__ out) {
    // one thread = one (row, c, ky) run of P contiguous pixels (P == 16: 64 B in, 32 B out)
    const int gw = W / P, gh = H / P;
    const long long total = (long long)B * gh * gw * 3 * P;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int px = (int)(i % gw);
        long long r = i / gw;
        const int ky = (int)(r % P); r /= P;
        const int c = (int)(r % 3); r /= 3;
        const int py = (int)(r % gh);
        const int b = (int)(r / gh);
        const float* src = x + (((long long)b * 3 + c) * H + (py * P + ky)) * W + px * P;
        __half* dst = out + (((long long)b * gh + py) * gw + px) * (3 * P * P) + c * P * P + ky * P;
        for (int k = 0; k < P; k += 8) {
            const float4 a = *reinterpret_cast<const float4*>(src + k);
            const float4 d = *reinterpret_cast<const float4*>(src + k + 4);
            *reinterpret_cast<uint4*>(dst + k) = make_uint4(pack_h2(a.x, a.y), pack_h2(a.z, a.w), pack_h2(d.x, d.y), pack_h2(d.z, d.w));
        }
    }
}

// ------------------------------------------------------------------------------------------ LayerNorm
constexpr int LN_MAXV = 10;  // float4 per lane: D <= 1280

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256)
layernorm_f16_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                     int rows_dst, int D, __half* __restrict__ out, int map, int tok_h, int tok_w, int ws, int g) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows_dst) return;
    long long src = warp;
    if (map == 1) {
        const int per_img = g * g * ws * ws;
        const int b = warp / per_img;
        const int rem = warp - b * per_img;
        const int win = rem / (ws * ws), t = rem - win * ws * ws;
        const int y = (win / g) * ws + t / ws, xq = (win % g) * ws + t % ws;
        src = (y < tok_h && xq < tok_w) ? ((long long)b * tok_h + y) * tok_w + xq : -1;
    }
    __half* o = out + (long long)warp * D;
    const int nv = D >> 2;  // float4 count
    if (src < 0) {
        for (int i = lane; i < (D >> 3); i += 32) reinterpret_cast<uint4*>(o)[i] = make_uint4(0, 0, 0, 0);
        return;
    }
    const float4* xr = reinterpret_cast<const float4*>(x + src * D);
    float4 v[LN_MAXV];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAXV; ++k) {
        const int i = lane + 32 * k;
        if (i < nv) { v[k] = xr[i]; s += (v[k].x + v[k].y) + (v[k].z + v[k].w); }
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAXV; ++k) {
        const int i = lane + 32 * k;
        if (i < nv) {
            const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
#pragma unroll
    for (int k = 0; k < LN_MAXV; ++k) {
        const int i = lane + 32 * k;
        if (i < nv) {
            const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + i);
            const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + i);
            const float a = (v[k].x - mean) * rstd * gm.x + bt.x, b = (v[k].y - mean) * rstd * gm.y + bt.y;
            const float c = (v[k].z - mean) * rstd * gm.z + bt.z, d = (v[k].w - mean) * rstd * gm.w + bt.w;
            reinterpret_cast<uint2*>(o)[i] = make_uint2(pack_h2(a, b), pack_h2(c, d));
        }
    }
}

__global__ void __launch_bounds__(256)
residual_ln_kernel(const float* __restrict__ x, LnFuse f, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                   int rows_dst, int D, __half* __restrict__ out, int map, int tok_h, int tok_w, int ws, int g) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows_dst) return;
    long long src = warp;
    if (map == 1) {
        const int per_img = g * g * ws * ws;
        const int b = warp / per_img;
        const int rem = warp - b * per_img;
        const int win = rem / (ws * ws), t = rem - win * ws * ws;
        const int y = (win / g) * ws + t / ws, xq = (win % g) * ws + t % ws;
        src = (y < tok_h && xq < tok_w) ? ((long long)b * tok_h + y) * tok_w + xq : -1;
    }
    const int nv = D >> 2;
    if (src < 0) {
        if (f.norm) {
            __half* o = out + (long long)warp * D;
            for (int i = lane; i < (D >> 3); i += 32) reinterpret_cast<uint4*>(o)[i] = make_uint4(0, 0, 0, 0);
        }
        return;
    }
    const float4* xr = reinterpret_cast<const float4*>(x + src * D);
    const uint2* ar = nullptr;
    if (f.add) {
        long long arow = src;
        if (f.add_map == 1) {
            const int per_tok = tok_h * tok_w;
            const int b = (int)(src / per_tok);
            const int rem = (int)(src - (long long)b * per_tok);
            const int y = rem / tok_w, xq = rem - y * tok_w;
            arow = ((long long)(b * g * g + (y / ws) * g + xq / ws) * ws + (y % ws)) * ws + xq % ws;
        }
        ar = reinterpret_cast<const uint2*>(f.add + arow * D);
    }
    float4 v[LN_MAXV];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAXV; ++k) {
        const int i = lane + 32 * k;
        if (i < nv) {
            v[k] = xr[i];
            if (ar) {
                const uint2 h = ar[i];
                const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&h.x));
                const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
                v[k].x += a.x; v[k].y += a.y; v[k].z += b.x; v[k].w += b.y;
            }
            s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
        }
    }
    if (f.x_out) {
        float4* xo = reinterpret_cast<float4*>(f.x_out + src * D);
#pragma unroll
        for (int k = 0; k < LN_MAXV; ++k) {
            const int i = lane + 32 * k;
            if (i < nv) xo[i] = v[k];
        }
    }
    if (f.cast_out) {
        const int b = (int)(src / f.tok_per_item);
        const int t = (int)(src - (long long)b * f.tok_per_item);
        if (t >= f.cast_skip) {
            uint2* co = reinterpret_cast<uint2*>(f.cast_out + ((long long)b * (f.tok_per_item - f.cast_skip) + t - f.cast_skip) * D);
#pragma unroll
            for (int k = 0; k < LN_MAXV; ++k) {
                const int i = lane + 32 * k;
                if (i < nv) co[i] = make_uint2(pack_h2(v[k].x, v[k].y), pack_h2(v[k].z, v[k].w));
            }
        }
    }
    if (!f.norm) return;
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAXV; ++k) {
        const int i = lane + 32 * k;
        if (i < nv) {
            const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
    __half* o = out + (long long)warp * D;
#pragma unroll
    for (int k = 0; k < LN_MAXV; ++k) {
        const int i = lane + 32 * k;
        if (i < nv) {
            const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + i);
            const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + i);
            const float a = (v[k].x - mean) * rstd * gm.x + bt.x, b = (v[k].y - mean) * rstd * gm.y + bt.y;
            const float c = (v[k].z - mean) * rstd * gm.z + bt.z, d = (v[k].w - mean) * rstd * gm.w + bt.w;
            reinterpret_cast<uint2*>(o)[i] = make_uint2(pack_h2(a, b), pack_h2(c, d));
        }
    }
}

// ------------------------------------------------------------------------------------------ casts / layouts
__global__ void cast_rows_f16_kernel(const float* __restrict__ x, int B, int T_src, int skip, int D, __half* __restrict__ out) {
    const int T = T_src - skip;
    const long long nv = (long long)B * T * (D >> 2);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
        const int dv = (int)(i % (D >> 2));
        const long long r = i / (D >> 2);
        const int t = (int)(r % T);
        const int b = (int)(r / T);
        const float4 a = reinterpret_cast<const float4*>(x + ((long long)b * T_src + skip + t) * D)[dv];
        reinterpret_cast<uint2*>(out + r * D)[dv] = make_uint2(pack_h2(a.x, a.y), pack_h2(a.z, a.w));
    }
}

__global__ void tokens_nchw_kernel(const float* __restrict__ x, int T_src, int skip, int D, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int T = T_src - skip;
    const int b = blockIdx.z, t0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int t = t0 + j, d = d0 + threadIdx.x;
        tile[j][threadIdx.x] = (t < T && d < D) ? x[((long long)b * T_src + skip + t) * D + d] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int d = d0 + j, t = t0 + threadIdx.x;
        if (t < T && d < D) out[((long long)b * D + d) * T + t] = tile[threadIdx.x][j];
    }
}

// ------------------------------------------------------------------------------------------ stem conv 3->32
__global__ void __launch_bounds__(128)
stem_conv_kernel(const float* __restrict__ x, int B, int H, int W, const float* __restrict__ w, const float* __restrict__ scale,
                 const float* __restrict__ shift, __half* __restrict__ out, int cpad) {
    __shared__ float ws[27 * 32];  // [tap(c,ky,kx)][co]
    __shared__ float sc[32], sh[32];
    for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) {
        const int co = i & 31, tap = i >> 5;
        ws[i] = w[co * 27 + tap];
    }
    if (threadIdx.x < 32) { sc[threadIdx.x] = scale[threadIdx.x]; sh[threadIdx.x] = shift[threadIdx.x]; }
    __syncthreads();
    const long long total = (long long)B * H * W;
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
        const int xq = (int)(p % W);
        const int y = (int)((p / W) % H);
        const int b = (int)(p / ((long long)W * H));
        float in[27];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int yy = y + ky - 1, xx = xq + kx - 1;
                    in[c * 9 + ky * 3 + kx] = (yy >= 0 && yy < H && xx >= 0 && xx < W)
                                                  ? __ldg(x + (((long long)b * 3 + c) * H + yy) * W + xx) : 0.f;
                }
        __half* o = out + p * cpad;
#pragma unroll
        for (int c8 = 0; c8 < 32; c8 += 8) {
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
            for (int t = 0; t < 27; ++t)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fmaf(in[t], ws[t * 32 + c8 + j], acc[j]);
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float a = fmaxf(fmaf(acc[2 * j], sc[c8 + 2 * j], sh[c8 + 2 * j]), 0.f);
                const float d = fmaxf(fmaf(acc[2 * j + 1], sc[c8 + 2 * j + 1], sh[c8 + 2 * j + 1]), 0.f);
                pk[j] = pack_h2(a, d);
            }
            *reinterpret_cast<uint4*>(o + c8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        for (int c8 = 32; c8 < cpad; c8 += 8) *reinterpret_cast<uint4*>(o + c8) = make_uint4(0, 0, 0, 0);
    }
}

// ------------------------------------------------------------------------------------------ tissue heads
// Stage 1: grid (chunks, B), 8 warps; each block LayerNorms LNM_ROWS token rows (one warp per row at a time) and writes the
// per-channel sum of its rows to part[b][chunk][C]. Stage 2: one block per image adds the chunk partials in a fixed
// order (deterministic), divides by T and applies the linear layer. C <= 256, C % 32 == 0.
constexpr int LNM_ROWS = 64;
__global__ void __launch_bounds__(256)
ln_mean_partial_kernel(const float* __restrict__ y, int T, int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                       float eps, float* __restrict__ part) {
    __shared__ float sp[8][256 + 1];
    const int b = blockIdx.y, chunk = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = C >> 5;  // channels per lane (per <= 8)
    float acc[8], gm[8], bt[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        acc[k] = 0.f;
        gm[k] = k < per ? gamma[lane + 32 * k] : 0.f;
        bt[k] = k < per ? beta[lane + 32 * k] : 0.f;
    }
    const int t_end = min(T, (chunk + 1) * LNM_ROWS);
    for (int t = chunk * LNM_ROWS + warp; t < t_end; t += 8) {
        const float* r = y + ((long long)b * T + t) * C;
        float v[8], s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < per) { v[k] = r[lane + 32 * k]; s += v[k]; }
        const float mean = warp_sum(s) / (float)C;
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < per) { const float d = v[k] - mean; q += d * d; }
        const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < per) acc[k] += (v[k] - mean) * rstd * gm[k] + bt[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (k < per) sp[warp][lane + 32 * k] = acc[k];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int wq = 0; wq < 8; ++wq) s += sp[wq][c];
        part[((long long)b * gridDim.x + chunk) * C + c] = s;
    }
}
__global__ void __launch_bounds__(256)
mean_linear_kernel(const float* __restrict__ part, int chunks, int T, int C, const float* __restrict__ w, const float* __restrict__ bias,
                   int n_out, float* __restrict__ out) {
    __shared__ float meanv[256];
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < chunks; ++k) s += part[((long long)b * chunks + k) * C + c];
        meanv[c] = s / (float)T;
    }
    __syncthreads();
    for (int o = warp; o < n_out; o += 8) {
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s += meanv[c] * w[o * C + c];
        s = warp_sum(s);
        if (lane == 0) out[b * n_out + o] = s + bias[o];
    }
}

__global__ void __launch_bounds__(256)
cls_head_kernel(const float* __restrict__ x, int T, int D, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                const float* __restrict__ w, const float* __restrict__ bias, int n_out, float* __restrict__ out) {
    __shared__ float v[2048];
    __shared__ float red[2][8];
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* r = x + (long long)b * T * D;
    float s = 0.f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) { v[i] = r[i]; s += v[i]; }
    s = warp_sum(s);
    if (lane == 0) red[0][warp] = s;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += red[0][i];
    const float mean = tot / (float)D;
    float q = 0.f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) { const float d = v[i] - mean; q += d * d; }
    q = warp_sum(q);
    if (lane == 0) red[1][warp] = q;
    __syncthreads();
    float tq = 0.f;
    for (int i = 0; i < 8; ++i) tq += red[1][i];
    const float rstd = 1.0f / sqrtf(tq / (float)D + eps);
    for (int i = threadIdx.x; i < D; i += blockDim.x) v[i] = (v[i] - mean) * rstd * gamma[i] + beta[i];
    __syncthreads();
    for (int o = warp; o < n_out; o += 8) {
        float a = 0.f;
        for (int c = lane; c < D; c += 32) a += v[c] * w[o * D + c];
        a = warp_sum(a);
        if (lane == 0) out[b * n_out + o] = a + bias[o];
    }
}

int grid_for(long long work_items, int block) {
    long long g = (work_items + block - 1) / block;
    const long long cap = (long long)cvb_num_sms() * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}


// Rows of a window-partitioned fp16 matrix [B, g*g windows, ws*ws tokens][N] that are PADDING (token outside the tok_h x tok_w
// grid) <- the per-column constant `bias`: what a Linear layer gives for the all-zero rows that the reference pads with
// AFTER norm1 (image_encoder.py:180-184, 263-288) -- so the GEMM itself only has to run over the real tokens.
__global__ void __launch_bounds__(256)
window_pad_fill_kernel(__half* __restrict__ out, const float* __restrict__ bias, int B, int N, int tok_h, int tok_w, int ws, int g) {
    // one warp per (padding token, 1024-column chunk): the padding tokens of an image are enumerated directly -- the strip right
    // of the real grid, then the rows below it -- instead of testing every window-partitioned row (one row in six is padding)
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int P = g * ws;                                    // padded grid edge
    const int right = tok_h * (P - tok_w), per_img = right + (P - tok_h) * P;
    const int chunks = (N + 1023) >> 10;
    if (warp >= B * per_img * chunks) return;
    const int chunk = warp % chunks, r = warp / chunks;
    const int b = r / per_img, i = r - b * per_img;
    int y, x;
    if (i < right) { y = i / (P - tok_w); x = tok_w + i % (P - tok_w); }
    else { const int j = i - right; y = tok_h + j / P; x = j % P; }
    const int wy = y / ws, wx = x / ws;
    const long long row = (long long)b * P * P + (wy * g + wx) * ws * ws + (y - wy * ws) * ws + (x - wx * ws);
    __half* o = out + row * N;
    const int end = min(N, (chunk + 1) << 10);
    for (int c = (chunk << 10) + lane * 8; c < end; c += 256) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(bias + c));
        const float4 d = __ldg(reinterpret_cast<const float4*>(bias + c + 4));
        *reinterpret_cast<uint4*>(o + c) = make_uint4(pack_h2(a.x, a.y), pack_h2(a.z, a.w), pack_h2(d.x, d.y), pack_h2(d.z, d.w));
    }
}

// ------------------------------------------------------------------------------------------ canvas helpers
// Tiles whose token grid is not one of the tile engine's native grids run their decoder on a zero-extended canvas
// (model.cu): planes are embedded top-left into the canvas, every layer's margin is re-zeroed (so that the next 3x3
// convolution sees the zero padding of the real image border), and the head outputs are cropped back.
template <class T> __device__ __forceinline__ T zero_of() { return T(0); }
template <> __device__ __forceinline__ uint4 zero_of<uint4>() { return make_uint4(0, 0, 0, 0); }

// dst[p][y][x] = (y < sH && x < sW) ? src[p][y][x] : 0   for y < dH, x < dW  (crop when dst is smaller, zero-extend when larger)
template <class T>
__global__ void copy_planes_kernel(const T* __restrict__ src, int sH, int sW, T* __restrict__ dst, int dH, int dW, long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % dW);
        const long long r = i / dW;
        const int y = (int)(r % dH);
        const long long p = r / dH;
        T v = zero_of<T>();
        if (y < sH && x < sW) v = src[(p * sH + y) * sW + x];
        dst[i] = v;
    }
}

// zero the pixels (y >= vH or x >= vW) of NHWC planes [planes][H][W][vec] (vec = 16-byte units per pixel)
__global__ void zero_margin_kernel(uint4* __restrict__ buf, long long planes, int H, int W, int vec, int vH, int vW) {
    const long long right = (long long)vH * (W - vW), per_plane = right + (long long)(H - vH) * W;
    const long long total = planes * per_plane * vec;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % vec);
        long long r = i / vec;
        const long long p = r / per_plane;
        r -= p * per_plane;
        int y, x;
        if (r < right) { y = (int)(r / (W - vW)); x = vW + (int)(r % (W - vW)); }
        else { r -= right; y = vH + (int)(r / W); x = (int)(r % W); }
        buf[((p * H + y) * W + x) * vec + v] = make_uint4(0, 0, 0, 0);
    }
}

}  // namespace

int op_patch_im2col(const float* x, int B, int H, int W, int P, __half* out, cudaStream_t stream) {
    CVB_CHECK(x && out && B > 0 && P % 8 == 0 && H % P == 0 && W % P == 0, CVB_ESHAPE, "patch_im2col: bad shape %dx%d P=%d", H, W, P);
    const long long total = (long long)B * (H / P) * (W / P) * 3 * P;
    patch_im2col_kernel<<<grid_for(total, 256), 256, 0, stream>>>(x, B, H, W, P, out);
    cvb_note_launches(1);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}

int op_layernorm_f16(const float* x, const float* gamma, const float* beta, float eps, int rows_dst, int D, __half* out,
                     int map, int B, int tok_h, int tok_w, int ws, int g, cudaStream_t stream) {
    (void)B;
    CVB_CHECK(x && gamma && beta && out && rows_dst > 0, CVB_EARG, "layernorm: null operand");
    CVB_CHECK(D % 8 == 0 && D <= LN_MAXV * 128, CVB_ESHAPE, "layernorm: D=%d must be a multiple of 8 and <= %d", D, LN_MAXV * 128);
    const int blocks = cdiv(rows_dst, 8);
    layernorm_f16_kernel<<<blocks, 256, 0, stream>>>(x, gamma, beta, eps, rows_dst, D, out, map, tok_h, tok_w, ws, g);
    cvb_note_launches(1);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}

int op_residual_ln(const float* x, const LnFuse& f, const float* gamma, const float* beta, float eps, int rows_dst, int D,
                   __half* out, int map, int B, int tok_h, int tok_w, int ws, int g, cudaStream_t stream) {
    (void)B;
    CVB_CHECK(x && rows_dst > 0, CVB_EARG, "residual_ln: null operand");
    CVB_CHECK(!f.norm || (gamma && beta && out), CVB_EARG, "residual_ln: norm requested without gamma/beta/out");
    CVB_CHECK(D % 8 == 0 && D <= LN_MAXV * 128, CVB_ESHAPE, "residual_ln: D=%d must be a multiple of 8 and <= %d", D, LN_MAXV * 128);
    CVB_CHECK(!f.cast_out || f.tok_per_item > f.cast_skip, CVB_EARG, "residual_ln: bad cast geometry");
    residual_ln_kernel<<<cdiv(rows_dst, 8), 256, 0, stream>>>(x, f, gamma, beta, eps, rows_dst, D, out, map, tok_h, tok_w, ws, g);
    cvb_note_launches(1);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}

int op_cast_rows_f16(const float* x, int B, int T_src, int skip, int D, __half* out, cudaStream_t stream) {
    CVB_CHECK(x && out && D % 4 == 0 && T_src > skip, CVB_EARG, "cast_rows: bad arguments");
    cast_rows_f16_kernel<<<grid_for((long long)B * (T_src - skip) * (D / 4), 256), 256, 0, stream>>>(x, B, T_src, skip, D, out);
    cvb_note_launches(1);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}

int op_tokens_nchw(const float* x, int B, int T_src, int skip, int D, float* out, cudaStream_t stream) {
    CVB_CHECK(x && out && T_src > skip, CVB_EARG, "tokens_nchw: bad arguments");
    dim3 grid(cdiv(T_src - skip, 32), cdiv(D, 32), B), block(32, 8);
    tokens_nchw_kernel<<<grid, block, 0, stream>>>(x, T_src, skip, D, out);
    cvb_note_launches(1);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}


int op_window_pad_fill(__half* out, const float* bias, int B, int N, int tok_h, int tok_w, int ws, int g, cudaStream_t stream) {
    CVB_CHECK(out && bias && B > 0 && N % 8 == 0 && ws > 0 && g > 0 && (((uintptr_t)out | (uintptr_t)bias) & 15) == 0, CVB_EARG,
              "window_pad_fill: bad arguments");
    if (g * ws == tok_h && g * ws == tok_w) return CVB_OK;  // no padding
    const int P = g * ws;
    CVB_CHECK(tok_h <= P && tok_w <= P, CVB_EARG, "window_pad_fill: token grid larger than the window grid");
    const long long warps = (long long)B * ((long long)tok_h * (P - tok_w) + (long long)(P - tok_h) * P) * ((N + 1023) / 1024);
    window_pad_fill_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, stream>>>(out, bias, B, N, tok_h, tok_w, ws, g);
    cvb_note_launches(1);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}

int op_copy_planes(const void* src, int sH, int sW, void* dst, int dH, int dW, long long planes, int elem_bytes, cudaStream_t stream) {
    CVB_CHECK(src && dst && sH > 0 && sW > 0 && dH > 0 && dW > 0 && planes > 0, CVB_EARG, "copy_planes: bad arguments");
    const long long total = planes * dH * dW;
    const int grid = grid_for(total, 256);
    if (elem_bytes == 16) {
        CVB_CHECK((((uintptr_t)src | (uintptr_t)dst) & 15) == 0, CVB_EARG, "copy_planes: 16-byte elements need 16-byte alignment");
        copy_planes_kernel<uint4><<<grid, 256, 0, stream>>>((const uint4*)src, sH, sW, (uint4*)dst, dH, dW, total);
    } else if (elem_bytes == 4) {
        copy_planes_kernel<float><<<grid, 256, 0, stream>>>((const float*)src, sH, sW, (float*)dst, dH, dW, total);
    } else if (elem_bytes == 1) {
        copy_planes_kernel<uint8_t><<<grid, 256, 0, stream>>>((const uint8_t*)src, sH, sW, (uint8_t*)dst, dH, dW, total);
    } else {
        cvb_set_error("copy_planes: element size %d not supported (1, 4, 16)", elem_bytes);
        return CVB_EARG;
    }
    cvb_note_launches(1);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}

int op_zero_margin(void* buf, long long planes, int H, int W, int bytes_per_px, int vH, int vW, cudaStream_t stream) {
    CVB_CHECK(buf && planes > 0 && bytes_per_px > 0 && bytes_per_px % 16 == 0 && vH > 0 && vW > 0 && vH <= H && vW <= W &&
                  (((uintptr_t)buf) & 15) == 0, CVB_EARG, "zero_margin: bad arguments");
    const long long per_plane = (long long)vH * (W - vW) + (long long)(H - vH) * W;
    if (per_plane == 0) return CVB_OK;
    const int vec = bytes_per_px / 16;
    zero_margin_kernel<<<grid_for(planes * per_plane * vec, 256), 256, 0, stream>>>((uint4*)buf, planes, H, W, vec, vH, vW);
    cvb_note_launches(1);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}

int op_stem_conv(const float* x, int B, int H, int W, const float* w, const float* scale, const float* shift,
                 __half* out, int cpad, cudaStream_t stream) {
    CVB_CHECK(x && w && scale && shift && out && cpad >= 32 && cpad % 8 == 0, CVB_EARG, "stem_conv: bad arguments");
    stem_conv_kernel<<<grid_for((long long)B * H * W, 128), 128, 0, stream>>>(x, B, H, W, w, scale, shift, out, cpad);
    cvb_note_launches(1);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}

size_t op_ln_mean_linear_scratch_floats(int B, int T, int C) { return (size_t)B * ((T + LNM_ROWS - 1) / LNM_ROWS) * C; }

int op_ln_mean_linear(const float* y, int B, int T, int C, const float* gamma, const float* beta, float eps,
                      const float* w, const float* b, int n_out, float* out, float* scratch, cudaStream_t stream) {
    CVB_CHECK(y && gamma && beta && w && b && out && scratch, CVB_EARG, "ln_mean_linear: null operand");
    CVB_CHECK(C % 32 == 0 && C <= 256, CVB_ESHAPE, "ln_mean_linear: C=%d must be a multiple of 32 and <= 256", C);
    const int chunks = (T + LNM_ROWS - 1) / LNM_ROWS;
    ln_mean_partial_kernel<<<dim3(chunks, B), 256, 0, stream>>>(y, T, C, gamma, beta, eps, scratch);
    mean_linear_kernel<<<B, 256, 0, stream>>>(scratch, chunks, T, C, w, b, n_out, out);
    cvb_note_launches(2);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}

int op_cls_head(const float* x, int B, int T, int D, const float* gamma, const float* beta, float eps,
                const float* w, const float* b, int n_out, float* out, cudaStream_t stream) {
    CVB_CHECK(x && gamma && beta && w && b && out && D <= 2048, CVB_EARG, "cls_head: bad arguments");
    cls_head_kernel<<<B, 256, 0, stream>>>(x, T, D, gamma, beta, eps, w, b, n_out, out);
    cvb_note_launches(1);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}
