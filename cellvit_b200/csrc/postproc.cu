// postproc.cu -- HoVer-Net post-processing of CellViT head maps on the device, bit-exact with the CPU oracle
// (oracle/postproc_oracle.c), which is itself pinned to the reference's cv2 / scipy code path
// (cell_segmentation/utils/post_proc_cellvit.py:155-249 stages P1-P7, :95-151 instance table).
//
// Stage -> kernel map (all HBM-bound integer / fp64 work on [B, H, W] planes, no tensor cores):
//   P1/P2  4-connected components of the NP mask + size filter (<10 px)        ccl_* , count, blb
//   P3     per-tile min-max normalisation of the HV maps (fp32 fma)            minmax_f32 (+ fused into sobel)
//   P4     Sobel k=21/11, fp64, OpenCV's exact operation order (no FMA)        sobel_kernel
//   P5     second normalisation, energy map, 3x3 Gaussian                      energy_kernel, blur_kernel
//   P6     marker mask, hole filling (background CCL), 5x5 ellipse opening,
//          CCL with scipy's raster-order numbering, size filter                fill/erode/dilate/ccl/scan/marker
//   P7     marker-controlled watershed: strict total order (value, age, index); every blob of the NP mask is an
//          independent flood -> one warp per blob, binary heap in shared memory    watershed_kernel
//   P8/P9  per-instance bbox / centroid / class histogram by atomics, compaction in id order   table_*
#include <float.h>

#include <vector>

#include "../../include/cellvit_b200.h"
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ uint32_t f32_ordered(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float f32_unordered(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}
__device__ __forceinline__ unsigned long long f64_ordered(double d) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double f64_unordered(unsigned long long u) {
    return __longlong_as_double((long long)((u >> 63) ? (u & 0x7FFFFFFFFFFFFFFFull) : ~u));
}
__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}
// cv2.normalize(NORM_MINMAX, 0..1, CV_32F) parameters (oracle minmax_params)
__device__ __forceinline__ void minmax_params(double mn, double mx, float* scale, float* shift) {
    const double d = __dsub_rn(mx, mn);
    const double sc = d > DBL_EPSILON ? __ddiv_rn(1.0, d) : 0.0;
    *scale = __double2float_rn(sc);
    *shift = __fsub_rn(0.0f, __double2float_rn(__dmul_rn(mn, (double)*scale)));
}

struct Dims { int B, H, W, N; };

// ------------------------------------------------------------------------------------------ input preparation
__global__ void prep_float_kernel(const float* __restrict__ np_map, const float* __restrict__ nt_map, int n_types, Dims d,
                                  uint8_t* __restrict__ npbin, uint8_t* __restrict__ tmap) {
    const long long total = (long long)d.B * d.N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / d.N), p = (int)(i - (long long)b * d.N);
        const float* q = np_map + (long long)b * 2 * d.N + p;
        npbin[i] = q[d.N] > q[0];  // torch.argmax: first maximum wins
        if (nt_map) {
            const float* t = nt_map + (long long)b * n_types * d.N + p;
            int best = 0;
            float bv = t[0];
            for (int c = 1; c < n_types; ++c) {
                const float v = t[(long long)c * d.N];
                if (v > bv) { bv = v; best = c; }
            }
            tmap[i] = (uint8_t)best;
        }
    }
}

__global__ void prep_maps_kernel(const int32_t* __restrict__ type_map, long long total, uint8_t* __restrict__ tmap) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = type_map[i];
        tmap[i] = (uint8_t)(t < 0 ? 255 : (t > 254 ? 255 : t));
    }
}

// per (tile, plane) min / max of a float plane -> ordered-u32 atomics; mm[(b*2+plane)*2 + {0,1}]
__global__ void minmax_f32_kernel(const float* __restrict__ hv, Dims d, uint32_t* __restrict__ mm) {
    const int plane = blockIdx.y;  // b*2 + c
    const float* src = hv + (long long)plane * d.N;
    float mn = INFINITY, mx = -INFINITY;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.N; i += gridDim.x * blockDim.x) {
        const float v = src[i];
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&mm[plane * 2 + 0], f32_ordered(mn));
        atomicMax(&mm[plane * 2 + 1], f32_ordered(mx));
    }
}

// ------------------------------------------------------------------------------------------ connected components
__device__ __forceinline__ int uf_find(volatile int* L, int a) {
    int p = L[a];
    while (p != a) { a = p; p = L[a]; }
    return a;
}
__device__ __forceinline__ void uf_unite(int* L, int a, int b) {
    bool done;
    do {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a < b) { const int old = atomicMin(&L[b], a); done = (old == b); b = old; }
        else if (b < a) { const int old = atomicMin(&L[a], b); done = (old == a); a = old; }
        else done = true;
    } while (!done);
}

// One warp per 32-pixel row segment. Labels start as the index of the first pixel of the horizontal run inside
// the segment (tile-local pixel index); -1 outside the mask.
__global__ void ccl_init_kernel(const uint8_t* __restrict__ mask, int want, Dims d, int* __restrict__ L) {
    const int segs = (d.W + 31) >> 5;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long total = (long long)d.B * d.H * segs;
    if (warp >= total) return;
    const int seg = (int)(warp % segs);
    const long long row = warp / segs;  // b*H + y
    const int x = seg * 32 + lane;
    const bool in = x < d.W;
    const long long gi = row * d.W + x;
    const bool fg = in && (mask[gi] != 0) == (want != 0);
    const uint32_t bits = __ballot_sync(0xffffffffu, fg);
    if (!in) return;
    int lab = -1;
    if (fg) {
        const uint32_t inv = ~bits & ((1u << lane) - 1u);
        const int start = inv ? 32 - __clz(inv) : 0;
        const int y = (int)(row % d.H);
        lab = y * d.W + seg * 32 + start;
    }
    L[gi] = lab;
}

__global__ void ccl_merge_kernel(const uint8_t* __restrict__ mask, int want, Dims d, int* __restrict__ L) {
    const long long total = (long long)d.B * d.N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / d.N), p = (int)(i - (long long)b * d.N);
        int* Lt = L + (long long)b * d.N;
        if (Lt[p] < 0) continue;
        const uint8_t* mt = mask + (long long)b * d.N;
        const int y = p / d.W, x = p - y * d.W;
        auto is = [&](int q) { return (mt[q] != 0) == (want != 0); };
        // horizontal: only the first pixel of a 32-segment can continue a run from the previous segment
        if ((x & 31) == 0 && x > 0 && is(p - 1)) uf_unite(Lt, p, p - 1);
        // vertical: skip when the pair to the left already joins the same two runs
        if (y > 0 && is(p - d.W)) {
            const bool dup = x > 0 && (x & 31) != 0 && is(p - 1) && is(p - d.W - 1);
            if (!dup) uf_unite(Lt, p, p - d.W);
        }
    }
}

__global__ void ccl_compress_kernel(Dims d, int* __restrict__ L) {
    const long long total = (long long)d.B * d.N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / d.N);
        int* Lt = L + (long long)b * d.N;
        const int p = (int)(i - (long long)b * d.N);
        if (Lt[p] >= 0) Lt[p] = uf_find(Lt, p);
    }
}

__global__ void count_kernel(const int* __restrict__ L, Dims d, int* __restrict__ cnt) {
    const long long total = (long long)d.B * d.N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = L[i];
        if (r >= 0) atomicAdd(&cnt[(i / d.N) * d.N + r], 1);
    }
}

__global__ void blb_kernel(const int* __restrict__ L, const int* __restrict__ cnt, Dims d, int min_size, uint8_t* __restrict__ blb) {
    const long long total = (long long)d.B * d.N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = L[i];
        blb[i] = (r >= 0 && cnt[(i / d.N) * d.N + r] >= min_size) ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------ exclusive scan (per tile)
constexpr int SCAN_BLOCK = 1024;  // elements per block (256 threads x 4)

__device__ __forceinline__ int block_excl_scan_256(int v, int* total_out) {
    __shared__ int wsum[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int base = 0, tot = 0;
    for (int k = 0; k < 8; ++k) {
        if (k < warp) base += wsum[k];
        tot += wsum[k];
    }
    __syncthreads();
    if (total_out) *total_out = tot;
    return base + inc - v;
}

// in: values (flag != 0 -> 1 when `as_flag`), bsum[b*nb + blk]
__global__ void __launch_bounds__(256) scan_reduce_kernel(const int* __restrict__ in, int as_flag, int n, int nb, int* __restrict__ bsum) {
    const int b = blockIdx.y, blk = blockIdx.x;
    const int* src = in + (long long)b * n;
    int s = 0;
    for (int k = 0; k < 4; ++k) {
        const int i = blk * SCAN_BLOCK + k * 256 + threadIdx.x;
        if (i < n) s += as_flag ? (src[i] != 0) : src[i];
    }
    int tot;
    block_excl_scan_256(s, &tot);
    if (threadIdx.x == 0) bsum[b * nb + blk] = tot;
}
__global__ void __launch_bounds__(256) scan_bsums_kernel(int nb, int* __restrict__ bsum) {
    int* s = bsum + (long long)blockIdx.x * nb;
    int carry = 0;
    for (int base = 0; base < nb; base += 256) {
        const int i = base + threadIdx.x;
        const int v = i < nb ? s[i] : 0;
        int tot;
        const int ex = block_excl_scan_256(v, &tot);
        if (i < nb) s[i] = carry + ex;
        carry += tot;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) scan_apply_kernel(const int* __restrict__ in, int as_flag, int n, int nb, const int* __restrict__ bsum,
                                                         int* __restrict__ out) {
    const int b = blockIdx.y, blk = blockIdx.x;
    const int* src = in + (long long)b * n;
    int* dst = out + (long long)b * n;
    int v[4], s = 0;
    const int i0 = blk * SCAN_BLOCK + threadIdx.x * 4;
    for (int k = 0; k < 4; ++k) {
        const int i = i0 + k;
        v[k] = i < n ? (as_flag ? (src[i] != 0) : src[i]) : 0;
        s += v[k];
    }
    int ex = block_excl_scan_256(s, nullptr) + bsum[b * nb + blk];
    for (int k = 0; k < 4; ++k) {
        const int i = i0 + k;
        if (i < n) dst[i] = ex;
        ex += v[k];
    }
}

// ------------------------------------------------------------------------------------------ Sobel (P3 + P4)
struct SobelTaps { double kd[32]; double ks[32]; int ksize; };
constexpr int SB_TW = 64, SB_TH = 32, SB_MAXR = 15;

// blockIdx.z = b*2 + plane. plane 0: h map, row kernel = derivative, column kernel = smoothing (symmetric);
// plane 1: v map, row kernel = smoothing, column kernel = derivative (anti-symmetric).
__global__ void __launch_bounds__(256)
sobel_kernel(const float* __restrict__ hv, Dims d, const uint32_t* __restrict__ mm, const SobelTaps taps,
             double* __restrict__ sob, unsigned long long* __restrict__ mm64) {
    extern __shared__ __align__(16) uint8_t sb_smem[];
    const int r = taps.ksize >> 1, ks = taps.ksize;
    const int in_w = SB_TW + 2 * r, in_h = SB_TH + 2 * r;
    float* s_in = reinterpret_cast<float*>(sb_smem);                                    // [in_h][in_w]
    double* s_tmp = reinterpret_cast<double*>(sb_smem + (((size_t)in_h * in_w * 4 + 15) & ~(size_t)15));  // [in_h][SB_TW]
    __shared__ double s_k[2][32];
    __shared__ unsigned long long s_mm[2];
    const int plane = blockIdx.z & 1, pl = blockIdx.z;
    const float* src = hv + (long long)pl * d.N;
    double* dst = sob + (long long)pl * d.N;
    const int x0 = blockIdx.x * SB_TW, y0 = blockIdx.y * SB_TH;
    float a, bsh;
    minmax_params((double)f32_unordered(mm[pl * 2]), (double)f32_unordered(mm[pl * 2 + 1]), &a, &bsh);
    if (threadIdx.x < 32) {
        s_k[0][threadIdx.x] = plane == 0 ? taps.kd[threadIdx.x] : taps.ks[threadIdx.x];  // row (x) kernel
        s_k[1][threadIdx.x] = plane == 0 ? taps.ks[threadIdx.x] : taps.kd[threadIdx.x];  // column (y) kernel
    }
    if (threadIdx.x == 0) { s_mm[0] = ~0ull; s_mm[1] = 0ull; }
    for (int i = threadIdx.x; i < in_h * in_w; i += 256) {
        const int yy = i / in_w, xx = i - yy * in_w;
        const int gy = reflect101(y0 - r + yy, d.H), gx = reflect101(x0 - r + xx, d.W);
        s_in[i] = __fmaf_rn(src[(long long)gy * d.W + gx], a, bsh);  // P3: one fused multiply-add in float
    }
    __syncthreads();
    // row pass: acc = k[0]*s[0]; acc += k[j]*s[j] (ascending j, product and sum rounded separately)
    for (int i = threadIdx.x; i < in_h * SB_TW; i += 256) {
        const int yy = i / SB_TW, xx = i - yy * SB_TW;
        const float* row = s_in + yy * in_w + xx;
        double acc = __dmul_rn(s_k[0][0], (double)row[0]);
        for (int j = 1; j < ks; ++j) acc = __dadd_rn(acc, __dmul_rn(s_k[0][j], (double)row[j]));
        s_tmp[i] = acc;
    }
    __syncthreads();
    double mn = INFINITY, mx = -INFINITY;
    for (int i = threadIdx.x; i < SB_TH * SB_TW; i += 256) {
        const int yy = i / SB_TW, xx = i - yy * SB_TW;
        const int gy = y0 + yy, gx = x0 + xx;
        if (gy >= d.H || gx >= d.W) continue;
        const double* col = s_tmp + (yy + r) * SB_TW + xx;
        double acc;
        if (plane == 0) {
            acc = __dmul_rn(s_k[1][r], col[0]);
            for (int j = 1; j <= r; ++j)
                acc = __dadd_rn(acc, __dmul_rn(s_k[1][r + j], __dadd_rn(col[j * SB_TW], col[-j * SB_TW])));
        } else {
            acc = 0.0;
            for (int j = 1; j <= r; ++j)
                acc = __dadd_rn(acc, __dmul_rn(s_k[1][r + j], __dsub_rn(col[j * SB_TW], col[-j * SB_TW])));
        }
        dst[(long long)gy * d.W + gx] = acc;
        mn = fmin(mn, acc);
        mx = fmax(mx, acc);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&s_mm[0], f64_ordered(mn));
        atomicMax(&s_mm[1], f64_ordered(mx));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicMin(&mm64[pl * 2 + 0], s_mm[0]);
        atomicMax(&mm64[pl * 2 + 1], s_mm[1]);
    }
}

// ------------------------------------------------------------------------------------------ energy map (P5) + marker mask (P6 head)
__global__ void energy_kernel(const double* __restrict__ sob, const unsigned long long* __restrict__ mm64, const uint8_t* __restrict__ blb,
                              Dims d, double* __restrict__ dist0, uint8_t* __restrict__ mk) {
    const long long total = (long long)d.B * d.N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / d.N), p = (int)(i - (long long)b * d.N);
        float m = 0.f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int pl = b * 2 + c;
            float af, bf;
            minmax_params(f64_unordered(mm64[pl * 2]), f64_unordered(mm64[pl * 2 + 1]), &af, &bf);
            // f64 -> f32 convertTo: (float)(src*(double)scale + (double)shift), product and sum rounded separately
            const float nrm = __double2float_rn(__dadd_rn(__dmul_rn(sob[(long long)pl * d.N + p], (double)af), (double)bf));
            const float one_minus = __fsub_rn(1.0f, nrm);
            m = c == 0 ? one_minus : (one_minus > m ? one_minus : m);  // np.maximum(sobelh, sobelv)
        }
        const int bl = blb[i];
        double o = __dsub_rn((double)m, (double)(1 - bl));
        if (o < 0) o = 0.0;
        dist0[i] = __dmul_rn(__dsub_rn(1.0, o), (double)bl);
        mk[i] = (bl - (o >= 0.4 ? 1 : 0)) > 0 ? 1 : 0;
    }
}

// dist = -GaussianBlur(dist0, (3,3), 0): kernel [.25,.5,.25], REFLECT_101, row pass ascending taps, column pass centre + pair
__global__ void blur_kernel(const double* __restrict__ dist0, Dims d, double* __restrict__ dist) {
    const long long total = (long long)d.B * d.N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / d.N), p = (int)(i - (long long)b * d.N);
        const int y = p / d.W, x = p - y * d.W;
        const double* src = dist0 + (long long)b * d.N;
        const int xm = reflect101(x - 1, d.W), xp = reflect101(x + 1, d.W);
        double t[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int yy = reflect101(y + k - 1, d.H);
            const double* row = src + (long long)yy * d.W;
            double acc = __dmul_rn(0.25, row[xm]);
            acc = __dadd_rn(acc, __dmul_rn(0.5, row[x]));
            acc = __dadd_rn(acc, __dmul_rn(0.25, row[xp]));
            t[k] = acc;
        }
        double acc = __dmul_rn(0.5, t[1]);
        acc = __dadd_rn(acc, __dmul_rn(0.25, __dadd_rn(t[2], t[0])));
        dist[i] = -acc;
    }
}

// ------------------------------------------------------------------------------------------ hole filling + opening (P6 a,b)
__global__ void border_flag_kernel(const int* __restrict__ Lbg, Dims d, int* __restrict__ flag) {
    const int per = 2 * d.W + 2 * d.H;
    const long long total = (long long)d.B * per;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / per);
        int k = (int)(i - (long long)b * per), p;
        if (k < d.W) p = k;
        else if (k < 2 * d.W) p = (d.H - 1) * d.W + (k - d.W);
        else if (k < 2 * d.W + d.H) p = (k - 2 * d.W) * d.W;
        else p = (k - 2 * d.W - d.H) * d.W + d.W - 1;
        const int r = Lbg[(long long)b * d.N + p];
        if (r >= 0) flag[(long long)b * d.N + r] = 1;
    }
}
__global__ void fill_kernel(const uint8_t* __restrict__ mk, const int* __restrict__ Lbg, const int* __restrict__ flag, Dims d,
                            uint8_t* __restrict__ out) {
    const long long total = (long long)d.B * d.N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = Lbg[i];
        out[i] = (mk[i] || (r >= 0 && flag[(i / d.N) * d.N + r] == 0)) ? 1 : 0;
    }
}
// 5x5 MORPH_ELLIPSE: rows {0,0,1,0,0},{1,1,1,1,1}x3,{0,0,1,0,0}; taps outside the image are ignored
__global__ void morph5_kernel(const uint8_t* __restrict__ src, Dims d, int dilate, uint8_t* __restrict__ dst) {
    const long long total = (long long)d.B * d.N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / d.N), p = (int)(i - (long long)b * d.N);
        const int y = p / d.W, x = p - y * d.W;
        const uint8_t* s = src + (long long)b * d.N;
        int v = dilate ? 0 : 1;
#pragma unroll
        for (int dy = -2; dy <= 2; ++dy) {
            const int yy = y + dy;
            if (yy < 0 || yy >= d.H) continue;
            const int half = (dy == -2 || dy == 2) ? 0 : 2;
            for (int dx = -half; dx <= half; ++dx) {
                const int xx = x + dx;
                if (xx < 0 || xx >= d.W) continue;
                const int q = s[yy * d.W + xx];
                if (dilate) v |= q; else v &= q;
            }
        }
        dst[i] = (uint8_t)v;
    }
}

// ------------------------------------------------------------------------------------------ marker ids (P6 c,d)
__global__ void root_flag_kernel(const int* __restrict__ L, Dims d, int* __restrict__ flag) {
    const long long total = (long long)d.B * d.N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        flag[i] = (L[i] == (int)(i % d.N)) ? 1 : 0;
}
// marker id = 1 + raster rank of the component's first pixel (scipy.ndimage.label), 0 if smaller than object_size.
// Also initialises the flood output: labels = marker inside blb, 0 elsewhere (skimage: markers * mask).
__global__ void marker_kernel(const int* __restrict__ L, const int* __restrict__ cnt, const int* __restrict__ rank, const uint8_t* __restrict__ blb,
                              Dims d, int object_size, int* __restrict__ marker, int* __restrict__ labels) {
    const long long total = (long long)d.B * d.N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = L[i];
        int id = 0;
        if (r >= 0) {
            const long long base = (i / d.N) * d.N;
            if (cnt[base + r] >= object_size) id = rank[base + r] + 1;
        }
        marker[i] = id;
        labels[i] = blb[i] ? id : 0;
    }
}

// ------------------------------------------------------------------------------------------ blob work lists (P7 setup)
// blobpix[b*N + off[root] + k] = tile-local pixel index of the k-th pixel of the blob (order irrelevant)
__global__ void blob_scatter_kernel(const int* __restrict__ L1, const uint8_t* __restrict__ blb, const int* __restrict__ off, Dims d,
                                    int* __restrict__ fill, int* __restrict__ blobpix) {
    const long long total = (long long)d.B * d.N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        if (!blb[i]) continue;
        const long long base = (i / d.N) * d.N;
        const int r = L1[i];
        const int k = atomicAdd(&fill[base + r], 1);
        blobpix[base + off[base + r] + k] = (int)(i - base);
    }
}
// Work queues by blob size class (largest first, so that the flood kernels run longest-job-first and end without a
// long tail): class 0 = more than small_cap pixels (LARGE launch), classes 1..5 = (small_cap/2, small_cap], ... halving,
// class 5 = everything smaller. qmeta[c] = count of class c, qmeta[8 + c] = its pop cursor. Entries are b*N + root.
constexpr int NQ_CLASSES = 6;
__device__ __forceinline__ int blob_size_class(int n, int small_cap) {
    if (n > small_cap) return 0;
    int c = 1, lim = small_cap >> 1;
    while (c < NQ_CLASSES - 1 && n <= lim) { ++c; lim >>= 1; }
    return c;
}
__global__ void blob_queue_kernel(const int* __restrict__ L1, const int* __restrict__ cnt, Dims d, int min_size, int small_cap,
                                  int* __restrict__ qmeta, int* __restrict__ queue, int qstride) {
    const long long total = (long long)d.B * d.N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i % d.N);
        if (L1[i] != p) continue;
        const int n = cnt[i];
        if (n < min_size) continue;
        const int c = blob_size_class(n, small_cap);
        queue[(long long)c * qstride + atomicAdd(&qmeta[c], 1)] = (int)i;
    }
}

// ------------------------------------------------------------------------------------------ watershed (P7)
struct HeapItem { double v; int age; int idx; };
__device__ __forceinline__ bool h_less(double va, int aa, int ia, double vb, int ab, int ib) {
    if (va != vb) return va < vb;
    if (aa != ab) return aa < ab;
    return ia < ib;
}
struct Heap {
    double* key;
    int2* pay;  // (age, idx)
    int n;
    __device__ __forceinline__ void sift_down(int i, double v, int age, int idx) {
        for (;;) {
            int c = 2 * i + 1;
            if (c >= n) break;
            double cv = key[c];
            int2 cp = pay[c];
            if (c + 1 < n) {
                const double rv = key[c + 1];
                const int2 rp = pay[c + 1];
                if (h_less(rv, rp.x, rp.y, cv, cp.x, cp.y)) { ++c; cv = rv; cp = rp; }
            }
            if (!h_less(cv, cp.x, cp.y, v, age, idx)) break;
            key[i] = cv; pay[i] = cp;
            i = c;
        }
        key[i] = v; pay[i] = make_int2(age, idx);
    }
    __device__ __forceinline__ void push(double v, int age, int idx) {
        int i = n++;
        while (i > 0) {
            const int par = (i - 1) >> 1;
            const double pv = key[par];
            const int2 pp = pay[par];
            if (!h_less(v, age, idx, pv, pp.x, pp.y)) break;
            key[i] = pv; pay[i] = pp;
            i = par;
        }
        key[i] = v; pay[i] = make_int2(age, idx);
    }
    __device__ __forceinline__ int pop() {  // returns idx of the minimum
        const int top = pay[0].y;
        --n;
        if (n > 0) sift_down(0, key[n], pay[n].x, pay[n].y);
        return top;
    }
};

// 4-ary heap of packed 16-byte entries in shared memory (one LDS.128 per entry, four independent child loads per
// level, depth log4 n). Any correct priority queue gives the same flood because (value, age, index) is a strict
// total order; the array has 4 slack entries so that the child loads never need a bounds check.
// Entries are two 64-bit integers: k = order-preserving image of the fp64 value (with -0.0 folded onto +0.0, as the
// floating-point comparison does), ai = age << 32 | index -- so "less" is two integer compares instead of a chain of
// fp64 ones.
struct __align__(16) HItem { unsigned long long k, ai; };
__device__ __forceinline__ unsigned long long f64_order_key(double v) {
    const long long b = __double_as_longlong(v + 0.0);
    return b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}
__device__ __forceinline__ HItem make_item(double v, int age, int idx) {
    HItem it;
    it.k = f64_order_key(v);
    it.ai = ((unsigned long long)(unsigned)age << 32) | (unsigned)idx;
    return it;
}
__device__ __forceinline__ bool it_less(const HItem& a, const HItem& b) { return a.k < b.k || (a.k == b.k && a.ai < b.ai); }
struct Heap4 {
    HItem* a;
    int n;
    __device__ __forceinline__ void sift_down(int i, HItem x) {
        for (;;) {
            const int c = 4 * i + 1;
            if (c >= n) break;
            HItem best = a[c];
            const HItem t1 = a[c + 1], t2 = a[c + 2], t3 = a[c + 3];
            int bi = c;
            if (c + 1 < n && it_less(t1, best)) { best = t1; bi = c + 1; }
            if (c + 2 < n && it_less(t2, best)) { best = t2; bi = c + 2; }
            if (c + 3 < n && it_less(t3, best)) { best = t3; bi = c + 3; }
            if (!it_less(best, x)) break;
            a[i] = best;
            i = bi;
        }
        a[i] = x;
    }
    __device__ __forceinline__ void push(HItem x) {
        int i = n++;
        while (i > 0) {
            const int par = (i - 1) >> 2;
            const HItem pv = a[par];
            if (!it_less(x, pv)) break;
            a[i] = pv;
            i = par;
        }
        a[i] = x;
    }
    __device__ __forceinline__ HItem pop() {
        const HItem top = a[0];
        --n;
        if (n > 0) sift_down(0, a[n]);
        return top;
    }
};

// ---- fast path of the flood: rank transform + packed 32-bit keys.
// The flood only ever COMPARES dist values, so a blob's fp64 values are replaced by their rank among the blob's pixels
// (rank = number of strictly smaller values: equal values keep equal ranks, so ties still fall through to age and index).
// A heap entry then is ONE 32-bit integer  rank << 22 | age << 12 | cell  (blob <= 1023 pixels: rank, age < 1024; bounding box
// + apron <= 4096 cells) and the strict total order (value, age, index) is a single unsigned compare. The heap is 4-ary with
// node k stored at h[k + 3], so the four children of a node are one aligned LDS.128; vacated slots hold 0xFFFFFFFF, which no
// key reaches, so the sift needs no bounds checks. Per blob: heap 4 KB + ranks (u16) + labels (i32) of the staged bounding
// box = 21 KB -- small enough for a flood CTA to be resident BESIDE a tile-engine CTA (~196 KB of an SM's 228 KB), which is
// what lets the post-processing stream overlap the forward at all: a CTA that does not fit waits for a kernel boundary,
// occupies the SM alone, and the statically scheduled persistent GEMM that follows runs that SM's tiles late.
constexpr int FL_MAXN = 1023;                 // pixels per blob on the fast path
constexpr int FL_HC = FL_MAXN + 1 + 8;        // heap array entries (node k at [k + 3], 4 slack entries)
constexpr int FL_RC = 4096;                   // staged cells (bounding box + 1-pixel apron): the key's 12-bit cell field
constexpr int FL_SMEM = FL_HC * 4 + FL_RC * 4 + FL_RC * 2;   // 28,704 B
constexpr int Q_DEFER = NQ_CLASSES;           // extra queue: blobs of classes 1..5 whose bounding box exceeds FL_RC cells
constexpr uint32_t FL_SENT = 0xFFFFFFFFu;

struct RankHeap {
    uint32_t* h;  // node k at h[k + 3]
    int n;
    __device__ __forceinline__ void sift_down(int i, uint32_t x) {
        for (;;) {
            const int c = 4 * i + 1;
            if (c >= n) break;
            const uint4 ch = *reinterpret_cast<const uint4*>(h + c + 3);
            const uint32_t m = min(min(ch.x, ch.y), min(ch.z, ch.w));
            if (m >= x) break;
            const int bi = ch.x == m ? 0 : (ch.y == m ? 1 : (ch.z == m ? 2 : 3));
            h[i + 3] = m;
            i = c + bi;
        }
        h[i + 3] = x;
    }
    __device__ __forceinline__ void push(uint32_t x) {
        int i = n++;
        while (i > 0) {
            const int par = (i - 1) >> 2;
            const uint32_t pv = h[par + 3];
            if (x >= pv) break;
            h[i + 3] = pv;
            i = par;
        }
        h[i + 3] = x;
    }
    __device__ __forceinline__ uint32_t pop() {
        const uint32_t top = h[3];
        --n;
        const uint32_t x = h[n + 3];
        h[n + 3] = FL_SENT;
        if (n > 0) sift_down(0, x);
        return top;
    }
};

// One warp per blob, blobs taken largest class first from the size-class queues. Seeds = marker pixels that still have an
// unlabelled mask neighbour (interior seeds pop as no-ops and never change `age`, so leaving them out does not change the
// result). Blobs above the fast path's limits run the same flood with a binary heap of (fp64 value, age, index) in global
// scratch, one L2 round trip per pop.
__global__ void __launch_bounds__(32)
watershed_kernel(const int* __restrict__ queue_base, int qstride, int* __restrict__ qmeta, int cls_begin, int cls_end, const int* __restrict__ cnt,
                 const int* __restrict__ off, const int* __restrict__ blobpix, const uint8_t* __restrict__ blb, const int* __restrict__ marker,
                 const double* __restrict__ dist, Dims d, int cap_entries, int rcap, double* __restrict__ gkey, int2* __restrict__ gpay, int* labels_) {
    // cap_entries > 0 selects the LARGE layout [heap: (cap_entries + 4) x 16 B][region dist: rcap x 8 B][region labels: rcap x 4 B]
    // (blobs above the rank path's limits, fp64 values and 16-byte heap entries); else: [heap u32 x FL_HC][labels i32 x FL_RC (aliased by the blob's fp64 values while they are ranked)][ranks u16 x FL_RC]
    extern __shared__ __align__(16) uint8_t ws_smem[];
    volatile int* labels = labels_;
    uint32_t* sheap = reinterpret_cast<uint32_t*>(ws_smem);
    int* slab = reinterpret_cast<int*>(ws_smem + (size_t)FL_HC * 4);
    double* svals = reinterpret_cast<double*>(slab);
    uint16_t* srank = reinterpret_cast<uint16_t*>(ws_smem + (size_t)FL_HC * 4 + (size_t)FL_RC * 4);
    static_assert((FL_HC * 4) % 16 == 0 && FL_MAXN * 8 <= FL_RC * 4, "flood shared-memory layout");
    const int lane = threadIdx.x;
    for (int cls = cls_begin; cls < cls_end; ++cls) {
    const int qn = qmeta[cls];
    const int* queue = queue_base + (long long)cls * qstride;
    int* qhead = qmeta + 8 + cls;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(qhead, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= qn) break;
        const int gi = queue[item];
        const long long base = ((long long)gi / d.N) * d.N;
        const int root = (int)(gi - base);
        const int n = cnt[base + root];
        const int* list = blobpix + base + off[base + root];
        const uint8_t* m = blb + base;
        const int* mk = marker + base;
        const double* ds = dist + base;
        volatile int* out = labels + base;
        // ---- fast path: the blob's bounding box (+1 pixel apron) is staged into shared memory and lane 0 runs the serial
        // priority flood entirely out of shared memory. Local raster indices order like the global ones, so the
        // (value, age, index) total order is unchanged.
        if (cap_entries == 0 && n <= FL_MAXN) {
            int y0 = d.H, y1 = -1, x0 = d.W, x1 = -1;
            for (int i = lane; i < n; i += 32) {
                const int p = list[i];
                const int y = p / d.W, x = p - y * d.W;
                y0 = min(y0, y); y1 = max(y1, y); x0 = min(x0, x); x1 = max(x1, x);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o));
                y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
                x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o));
                x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o));
            }
            const int rw = x1 - x0 + 3, rh = y1 - y0 + 3;
            const long long cells_ll = (long long)rw * rh;
            if (cells_ll <= (long long)FL_RC) {
                const int cells = (int)cells_ll;
                int* gout = labels_ + base;
                auto cell_of = [&](int p) { const int y = p / d.W, x = p - y * d.W; return (y - y0 + 1) * rw + (x - x0 + 1); };
                // (1) rank transform: vals[i] = dist of the i-th blob pixel; rank = number of strictly smaller values
                for (int i = lane; i < n; i += 32) svals[i] = ds[list[i]];
                for (int i = lane; i < FL_HC; i += 32) sheap[i] = FL_SENT;
                __syncwarp();
                for (int i0 = lane; i0 < n; i0 += 128) {   // four pixels per lane per sweep over the values
                    const int i1 = i0 + 32, i2 = i0 + 64, i3 = i0 + 96;
                    const double v0 = svals[i0], v1 = svals[i1 < n ? i1 : i0], v2 = svals[i2 < n ? i2 : i0], v3 = svals[i3 < n ? i3 : i0];
                    int r0 = 0, r1 = 0, r2 = 0, r3 = 0;
                    for (int j = 0; j < n; ++j) {
                        const double vj = svals[j];
                        r0 += vj < v0; r1 += vj < v1; r2 += vj < v2; r3 += vj < v3;
                    }
                    srank[cell_of(list[i0])] = (uint16_t)r0;
                    if (i1 < n) srank[cell_of(list[i1])] = (uint16_t)r1;
                    if (i2 < n) srank[cell_of(list[i2])] = (uint16_t)r2;
                    if (i3 < n) srank[cell_of(list[i3])] = (uint16_t)r3;
                }
                __syncwarp();
                // (2) labels of the region: only this blob's pixels enter (via its pixel list); everything else, including
                // pixels of other blobs inside the bounding box and the whole apron, is "outside the mask" (-1)
                for (int c = lane; c < cells; c += 32) slab[c] = -1;
                __syncwarp();
                for (int i = lane; i < n; i += 32) {
                    const int p = list[i];
                    slab[cell_of(p)] = gout[p];
                }
                __syncwarp();
                // (3) seeds
                RankHeap hq;
                hq.h = sheap;
                int n_seed = 0;
                for (int c0 = 0; c0 < cells; c0 += 32) {
                    const int c = c0 + lane;
                    bool is_seed = false;
                    if (c < cells && slab[c] > 0)  // labelled cells are never on the apron, so the four neighbours exist
                        is_seed = slab[c - rw] == 0 || slab[c - 1] == 0 || slab[c + 1] == 0 || slab[c + rw] == 0;
                    const uint32_t bits = __ballot_sync(0xffffffffu, is_seed);
                    if (is_seed) sheap[3 + n_seed + __popc(bits & ((1u << lane) - 1u))] = ((uint32_t)srank[c] << 22) | (uint32_t)c;
                    n_seed += __popc(bits);
                }
                __syncwarp();
                // (4) serial priority flood
                if (lane == 0) {
                    hq.n = n_seed;
                    for (int i = (n_seed - 2) / 4; i >= 0 && n_seed > 1; --i) hq.sift_down(i, sheap[i + 3]);  // Floyd heapify
                    uint32_t age = 0;
                    while (hq.n > 0) {
                        const int c = (int)(hq.pop() & 0xFFFu);
                        const int lab = slab[c];
                        const int qs[4] = {c - rw, c - 1, c + 1, c + rw};  // up, left, right, down
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int q = qs[k];
                            if (slab[q] == 0) {
                                ++age;
                                slab[q] = lab;  // labelled at push time
                                hq.push(((uint32_t)srank[q] << 22) | (age << 12) | (uint32_t)q);
                            }
                        }
                    }
                }
                __syncwarp();
                for (int i = lane; i < n; i += 32) {
                    const int p = list[i];
                    const int lab = slab[cell_of(p)];
                    if (lab > 0 && gout[p] == 0) gout[p] = lab;  // pixels this flood labelled
                }
                __syncwarp();
                continue;
            }
        }
        if (cap_entries == 0) {
            // few pixels but a bounding box beyond the rank path's region (a thin diagonal chain of nuclei): hand the blob to the
            // LARGE-layout launch that follows instead of flooding it through L2 (2 us per pop: one such blob took 1.5 ms)
            if (lane == 0) const_cast<int*>(queue_base)[(long long)Q_DEFER * qstride + atomicAdd(&qmeta[Q_DEFER], 1)] = gi;
            continue;
        }
        {
            HItem* lheap = reinterpret_cast<HItem*>(ws_smem);
            double* ldist = reinterpret_cast<double*>(ws_smem + (size_t)(cap_entries + 4) * 16);
            int* llab = reinterpret_cast<int*>(ws_smem + (size_t)(cap_entries + 4) * 16 + (size_t)rcap * 8);
        {
            int y0 = d.H, y1 = -1, x0 = d.W, x1 = -1;
            for (int i = lane; i < n; i += 32) {
                const int p = list[i];
                const int y = p / d.W, x = p - y * d.W;
                y0 = min(y0, y); y1 = max(y1, y); x0 = min(x0, x); x1 = max(x1, x);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o));
                y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
                x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o));
                x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o));
            }
            const int rw = x1 - x0 + 3, rh = y1 - y0 + 3;
            const long long cells_ll = (long long)rw * rh;
            if (n <= cap_entries && cells_ll <= (long long)rcap) {
                const int cells = (int)cells_ll;
                int* gout = labels_ + base;
                // only this blob's pixels enter the region (via its pixel list): everything else, including pixels of
                // other blobs inside the bounding box and the whole apron, is "outside the mask"
                for (int c = lane; c < cells; c += 32) llab[c] = -1;
                __syncwarp();
                for (int i = lane; i < n; i += 32) {
                    const int p = list[i];
                    const int y = p / d.W, x = p - y * d.W;
                    const int c = (y - y0 + 1) * rw + (x - x0 + 1);
                    llab[c] = gout[p];
                    ldist[c] = ds[p];
                }
                __syncwarp();
                Heap4 hq;
                hq.a = lheap;
                int n_seed = 0;
                for (int c0 = 0; c0 < cells; c0 += 32) {
                    const int c = c0 + lane;
                    bool is_seed = false;
                    if (c < cells && llab[c] > 0)  // labelled cells are never on the apron, so the four neighbours exist
                        is_seed = llab[c - rw] == 0 || llab[c - 1] == 0 || llab[c + 1] == 0 || llab[c + rw] == 0;
                    const uint32_t bits = __ballot_sync(0xffffffffu, is_seed);
                    if (is_seed) lheap[n_seed + __popc(bits & ((1u << lane) - 1u))] = make_item(ldist[c], 0, c);
                    n_seed += __popc(bits);
                }
                __syncwarp();
                if (lane == 0) {
                    hq.n = n_seed;
                    for (int i = (n_seed - 2) / 4; i >= 0 && n_seed > 1; --i) hq.sift_down(i, lheap[i]);  // Floyd heapify
                    int age = 0;
                    while (hq.n > 0) {
                        const HItem t = hq.pop();
                        const int c = (int)(unsigned)t.ai;
                        const int lab = llab[c];
                        const int qs[4] = {c - rw, c - 1, c + 1, c + rw};  // up, left, right, down
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int q = qs[k];
                            if (llab[q] == 0) {
                                ++age;
                                llab[q] = lab;  // labelled at push time
                                hq.push(make_item(ldist[q], age, q));
                            }
                        }
                    }
                }
                __syncwarp();
                for (int i = lane; i < n; i += 32) {
                    const int p = list[i];
                    const int y = p / d.W, x = p - y * d.W;
                    const int lab = llab[(y - y0 + 1) * rw + (x - x0 + 1)];
                    if (lab > 0 && gout[p] == 0) gout[p] = lab;  // pixels this flood labelled
                }
                __syncwarp();
                continue;
            }
        }
        }
        Heap hp;
        hp.key = gkey + base + off[base + root];
        hp.pay = gpay + base + off[base + root];
        hp.n = 0;
        // ---- collect boundary seeds (any order: the heap order is a strict total order)
        int n_seed = 0;
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            bool is_seed = false;
            int p = 0;
            if (i < n) {
                p = list[i];
                if (mk[p] > 0) {
                    const int y = p / d.W, x = p - y * d.W;
                    is_seed = (y > 0 && m[p - d.W] && mk[p - d.W] == 0) || (x > 0 && m[p - 1] && mk[p - 1] == 0) ||
                              (x < d.W - 1 && m[p + 1] && mk[p + 1] == 0) || (y < d.H - 1 && m[p + d.W] && mk[p + d.W] == 0);
                }
            }
            const uint32_t bits = __ballot_sync(0xffffffffu, is_seed);
            if (is_seed) {
                const int slot = n_seed + __popc(bits & ((1u << lane) - 1u));
                hp.key[slot] = ds[p];
                hp.pay[slot] = make_int2(0, p);
            }
            n_seed += __popc(bits);
        }
        __syncwarp();
        hp.n = n_seed;
        if (lane == 0) {
            // Floyd heapify
            for (int i = n_seed / 2 - 1; i >= 0; --i) hp.sift_down(i, hp.key[i], hp.pay[i].x, hp.pay[i].y);
        }
        __syncwarp();
        // ---- flood: pop min; neighbours up, left, right, down; label at push time.
        // One global round trip per pop: lanes 0..3 fetch mask / label / dist of their neighbour and lane 4 the
        // label of the popped pixel, all in flight together.
        int age = 0;
        int hn = __shfl_sync(0xffffffffu, hp.n, 0);
        while (hn > 0) {
            int p = 0;
            if (lane == 0) p = hp.pop();
            p = __shfl_sync(0xffffffffu, p, 0);
            const int y = p / d.W, x = p - y * d.W;
            int q = -1;
            if (lane == 0 && y > 0) q = p - d.W;
            else if (lane == 1 && x > 0) q = p - 1;
            else if (lane == 2 && x < d.W - 1) q = p + 1;
            else if (lane == 3 && y < d.H - 1) q = p + d.W;
            int mq = 0, oq = 1;
            double qv = 0.0;
            if (q >= 0) { mq = m[q]; oq = out[q]; qv = ds[q]; }
            int lab = lane == 4 ? out[p] : 0;
            const bool take = q >= 0 && mq != 0 && oq == 0;
            const uint32_t bits = __ballot_sync(0xffffffffu, take) & 0xFu;
            lab = __shfl_sync(0xffffffffu, lab, 4);
            for (int k = 0; k < 4; ++k) {
                if (!(bits >> k & 1u)) continue;
                const int qq = __shfl_sync(0xffffffffu, q, k);
                const double vv = __shfl_sync(0xffffffffu, qv, k);
                if (lane == 0) {
                    ++age;
                    out[qq] = lab;
                    hp.push(vv, age, qq);
                }
            }
            __syncwarp();
            hn = __shfl_sync(0xffffffffu, hp.n, 0);
        }
    }
    }  // size classes
}

// ------------------------------------------------------------------------------------------ instance table (P8/P9)
struct Acc {  // 64 bytes per id
    int area;
    int rmin, rmax, cmin, cmax;
    int pad;
    unsigned long long sx, sy;
    int hist[8];
};
__global__ void table_init_kernel(Acc* acc, long long n, int H, int W) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        Acc a;
        a.area = 0; a.rmin = H; a.rmax = -1; a.cmin = W; a.cmax = -1; a.pad = 0; a.sx = 0; a.sy = 0;
        for (int k = 0; k < 8; ++k) a.hist[k] = 0;
        acc[i] = a;
    }
}
__global__ void table_accum_kernel(const int* __restrict__ labels, const uint8_t* __restrict__ tmap, Dims d, int cap, Acc* __restrict__ acc,
                                   int* __restrict__ status, int* __restrict__ maxid) {
    // Warp-aggregated: the 32 pixels of a warp are consecutive in one image row (H * W and the stride are multiples of 32) and
    // carry one or two distinct labels, so each group of equal labels is reduced inside the warp (match_any + redux) and its
    // leader issues ONE set of atomics -- the per-pixel version spent its time in ~8 contended atomics per foreground pixel.
    const long long total = (long long)d.B * d.N;
    const int lane = threadIdx.x & 31;
    const long long n_iter = (total + (long long)gridDim.x * blockDim.x - 1) / ((long long)gridDim.x * blockDim.x);
    for (long long it = 0; it < n_iter; ++it) {
        const long long i = it * (long long)gridDim.x * blockDim.x + blockIdx.x * (long long)blockDim.x + threadIdx.x;
        const bool in = i < total;
        const int l = in ? labels[i] : 0;
        const int b = in ? (int)(i / d.N) : 0, p = in ? (int)(i - (long long)b * d.N) : 0;
        if (in && l == 0) {  // only "is there any background" matters (the np.unique(...)[1:] quirk)
            if (acc[(long long)b * cap].area == 0) acc[(long long)b * cap].area = 1;
        }
        if (in && l >= cap) atomicOr(&status[b], 1);
        const bool fg = in && l > 0 && l < cap;
        const uint32_t act = __ballot_sync(0xffffffffu, fg);
        if (!fg) continue;
        const int y = p / d.W, x = p - y * d.W;
        const int t = tmap ? (int)tmap[i] : 8;
        // (b, label) identifies the instance; lanes of one warp can straddle two tiles only if N % 32 != 0 -- keyed anyway
        const uint32_t grp = __match_any_sync(act, (unsigned long long)b << 32 | (unsigned)l);
        const int cnt = __popc(grp);
        const int sx = __reduce_add_sync(grp, x), sy = __reduce_add_sync(grp, y);
        const int xmin = __reduce_min_sync(grp, x), xmax = __reduce_max_sync(grp, x);
        const int ymin = __reduce_min_sync(grp, y), ymax = __reduce_max_sync(grp, y);
        Acc* a = acc + (long long)b * cap + l;
        if (lane == __ffs(grp) - 1) {
            if (l > maxid[b]) atomicMax(&maxid[b], l);  // bounds the id range table_finalize_kernel scans
            atomicAdd(&a->area, cnt);
            atomicAdd(&a->sx, (unsigned long long)sx);
            atomicAdd(&a->sy, (unsigned long long)sy);
            if (ymin < a->rmin) atomicMin(&a->rmin, ymin);
            if (ymax > a->rmax) atomicMax(&a->rmax, ymax);
            if (xmin < a->cmin) atomicMin(&a->cmin, xmin);
            if (xmax > a->cmax) atomicMax(&a->cmax, xmax);
        }
        if (tmap) {
            const uint32_t g2 = __match_any_sync(grp, t);   // lanes of this instance with the same class
            if (t < 8 && lane == __ffs(g2) - 1) atomicAdd(&a->hist[t], __popc(g2));
        }
    }
}
// one block per tile: compact present ids in ascending order into rows
__global__ void __launch_bounds__(256)
table_finalize_kernel(const Acc* __restrict__ acc, int cap_, const int* __restrict__ maxid, int n_types, int max_rows,
                      cvb_inst_row* __restrict__ table, int* __restrict__ counts) {
    const int b = blockIdx.x;
    const int cap = min(cap_, maxid[b] + 1);  // ids above the largest label of the tile are absent
    const Acc* A = acc + (long long)b * cap_;
    cvb_inst_row* rows = table + (long long)b * max_rows;
    __shared__ int s_base, s_first;
    if (threadIdx.x == 0) { s_base = 0; s_first = (A[0].area > 0) ? 0 : 1; }  // no background -> drop the smallest id
    __syncthreads();
    for (int c0 = 1; c0 < cap; c0 += 256) {
        const int id = c0 + threadIdx.x;
        const bool present = id < cap && A[id].area > 0;
        int tot;
        const int ex = block_excl_scan_256(present ? 1 : 0, &tot);
        const int base = s_base, drop = s_first;
        __syncthreads();
        if (present) {
            const int pos = base + ex - drop;  // position among present ids, minus the dropped first one
            if (pos >= 0 && pos < max_rows) {
                const Acc a = A[id];
                cvb_inst_row r;
                r.id = id;
                r.rmin = a.rmin; r.cmin = a.cmin; r.rmax = a.rmax + 1; r.cmax = a.cmax + 1;
                r.area = a.area;
                const double m00 = (double)a.area;
                r.cx = __dadd_rn(__ddiv_rn((double)((long long)a.sx - (long long)a.cmin * a.area), m00), (double)a.cmin);
                r.cy = __dadd_rn(__ddiv_rn((double)((long long)a.sy - (long long)a.rmin * a.area), m00), (double)a.rmin);
                int best = -1, second = -1;
                for (int t = 0; t < n_types && t < 8; ++t) {
                    r.hist[t] = a.hist[t];
                    if (a.hist[t] == 0) continue;
                    if (best < 0 || a.hist[t] > a.hist[best]) { second = best; best = t; }
                    else if (second < 0 || a.hist[t] > a.hist[second]) second = t;
                }
                for (int t = (n_types < 8 ? (n_types < 0 ? 0 : n_types) : 8); t < 8; ++t) r.hist[t] = 0;
                int ty = best;
                if (ty == 0 && second >= 0) ty = second;
                r.type = ty;
                r.type_prob = ty >= 0 ? __ddiv_rn((double)a.hist[ty], __dadd_rn((double)a.area, 1.0e-6)) : 0.0;
                r.type_prob_f = (float)r.type_prob;
                rows[pos] = r;
            }
        }
        if (threadIdx.x == 0) s_base = base + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int n = s_base - s_first;
        counts[b] = n < 0 ? 0 : n;
    }
}

// ------------------------------------------------------------------------------------------ contours
// cv2.findContours(crop, RETR_TREE, CHAIN_APPROX_SIMPLE)[0][0] of every instance (post_proc_cellvit.py:106-125) =
// the Suzuki-Abe outer border of the instance, started at its raster-first pixel, 8-neighbour search order
// right, up-right, up, up-left, left, down-left, down, down-right, a point emitted whenever the step direction
// changes. Exact for instances with one 8-connected component (cv2 lists the LAST component first otherwise):
// an 8-connectivity CCL of the label map counts the components of every id, and ids with more than one are
// flagged (npts = -1) for the host to resolve with cv2 itself.
__global__ void lab8_init_kernel(const int* __restrict__ labels, Dims d, int* __restrict__ L) {
    const long long total = (long long)d.B * d.N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        L[i] = labels[i] > 0 ? (int)(i % d.N) : -1;
}
__global__ void lab8_merge_kernel(const int* __restrict__ labels, Dims d, int* __restrict__ L) {
    const long long total = (long long)d.B * d.N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / d.N), p = (int)(i - (long long)b * d.N);
        const int* lt = labels + (long long)b * d.N;
        const int id = lt[p];
        if (id <= 0) continue;
        int* Lt = L + (long long)b * d.N;
        const int y = p / d.W, x = p - y * d.W;
        if (x > 0 && lt[p - 1] == id) uf_unite(Lt, p, p - 1);
        if (y > 0) {
            if (lt[p - d.W] == id) uf_unite(Lt, p, p - d.W);
            else {  // diagonals only matter when the pixel above does not already join them
                if (x > 0 && lt[p - d.W - 1] == id && lt[p - 1] != id) uf_unite(Lt, p, p - d.W - 1);
                if (x < d.W - 1 && lt[p - d.W + 1] == id) uf_unite(Lt, p, p - d.W + 1);
            }
        }
    }
}
__global__ void lab8_count_kernel(const int* __restrict__ labels, const int* __restrict__ L, Dims d, int cap, int* __restrict__ ncomp) {
    const long long total = (long long)d.B * d.N;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i % d.N);
        if (L[i] != p) continue;
        const int id = labels[i];
        if (id > 0 && id < cap) atomicAdd(&ncomp[(i / d.N) * cap + id], 1);
    }
}
// one thread per table row
__global__ void contour_kernel(const int* __restrict__ labels, const cvb_inst_row* __restrict__ table, const int* __restrict__ counts,
                               const int* __restrict__ ncomp, Dims d, int cap, int max_rows, int max_pts, short2* __restrict__ pts,
                               int* __restrict__ npts) {
    const int b = blockIdx.y;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = min(counts[b], max_rows);
    if (row >= n) return;
    const cvb_inst_row r = table[(long long)b * max_rows + row];
    int* out_n = npts + (long long)b * max_rows + row;
    if (r.id >= cap || (ncomp != nullptr && ncomp[(long long)b * cap + r.id] != 1)) { *out_n = -1; return; }
    const int* lt = labels + (long long)b * d.N;
    short2* out = pts + ((long long)b * max_rows + row) * max_pts;
    const int id = r.id, W = d.W, H = d.H;
    auto in = [&](int y, int x) { return y >= 0 && y < H && x >= 0 && x < W && lt[y * W + x] == id; };
    const int dx[8] = {1, 1, 0, -1, -1, -1, 0, 1}, dy[8] = {0, -1, -1, -1, 0, 1, 1, 1};
    int y0 = r.rmin, x0 = r.cmin;
    while (x0 < r.cmax && !in(y0, x0)) ++x0;  // raster-first pixel: left-most pixel of the top bbox row
    int s = 4, s_end = 4;
    bool found = false;
    do {
        s = (s - 1) & 7;
        if (in(y0 + dy[s], x0 + dx[s])) { found = true; break; }
    } while (s != s_end);
    if (!found) {  // isolated pixel
        out[0] = make_short2((short)x0, (short)y0);
        *out_n = 1;
        return;
    }
    const int y1 = y0 + dy[s], x1 = x0 + dx[s];
    int y3 = y0, x3 = x0, prev_s = s ^ 4, k = 0;
    for (int guard = 0; guard < 4 * (H + W) * 8; ++guard) {
        int y4, x4;
        for (;;) {
            ++s;
            y4 = y3 + dy[s & 7];
            x4 = x3 + dx[s & 7];
            if (in(y4, x4)) break;
        }
        s &= 7;
        if (s != prev_s) {
            if (k < max_pts) out[k] = make_short2((short)x3, (short)y3);
            ++k;
        }
        prev_s = s;
        if (y4 == y0 && x4 == x0 && y3 == y1 && x3 == x1) break;
        y3 = y4; x3 = x4;
        s = (s + 4) & 7;
    }
    *out_n = k <= max_pts ? k : -1;
}

// ------------------------------------------------------------------------------------------ host orchestration
void sobel_taps_host(int ksize, SobelTaps* t) {
    // cv::getSobelKernels for ksize > 7 (integer recurrences), orders 1 and 0
    for (int order = 0; order < 2; ++order) {
        long long ker[64] = {0};
        ker[0] = 1;
        for (int i = 0; i < ksize - order - 1; ++i) {
            long long oldv = ker[0];
            for (int j = 1; j <= ksize; ++j) { const long long nv = ker[j] + ker[j - 1]; ker[j - 1] = oldv; oldv = nv; }
        }
        for (int i = 0; i < order; ++i) {
            long long oldv = -ker[0];
            for (int j = 1; j <= ksize; ++j) { const long long nv = ker[j - 1] - ker[j]; ker[j - 1] = oldv; oldv = nv; }
        }
        for (int j = 0; j < 32; ++j) (order ? t->kd : t->ks)[j] = j < ksize ? (double)ker[j] : 0.0;
    }
    t->ksize = ksize;
}

struct Carve {
    uint8_t* base; size_t off;
    template <class T> T* take(size_t n) { off = align_up(off, 256); T* p = base ? reinterpret_cast<T*>(base + off) : nullptr; off += n * sizeof(T); return p; }
};
struct Ws {
    uint8_t *npbin, *tmap, *blb, *mk, *mk2;
    int *L1, *cnt1, *off1, *fill1, *blobpix, *Lx, *cntx, *flagx, *rankx, *marker, *bsum, *queue, *qmeta, *status;
    int qstride;
    uint32_t* mm;
    unsigned long long* mm64;
    double *sob, *dist0, *dist, *gkey;
    int2* gpay;
    Acc* acc;
    int cap;
    size_t bytes;
};
// LARGE layout of the flood kernel: 4096 heap entries (16 B) + 11,264 staged cells (12 B) = 200 KB, one CTA per SM
constexpr int FL_LARGE_CAP = 4096, FL_LARGE_RCAP = 11264;
constexpr int FL_LARGE_SMEM = (FL_LARGE_CAP + 4) * 16 + FL_LARGE_RCAP * 12;
int g_debug_skip_flood = 0;
int g_flood_large_ctas = 0;   // 0: one CTA per SM (debug knob: fewer CTAs hold fewer SMs, each for longer)

// side stream + events for the forked flood launch of the big blobs (one set per host thread and device)
struct Fork { cudaStream_t side; cudaEvent_t fork, join; int dev; };
Fork* get_fork() {
    thread_local std::vector<Fork> forks;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    for (Fork& f : forks) if (f.dev == dev) return &f;
    Fork f{};
    f.dev = dev;
    if (cudaStreamCreateWithFlags(&f.side, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&f.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&f.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    forks.push_back(f);
    return &forks.back();
}

Ws carve(void* base, int B, int H, int W) {
    Ws w{};
    Carve c{reinterpret_cast<uint8_t*>(base), 0};
    const size_t BN = (size_t)B * H * W;
    const int nb = ((H * W) + SCAN_BLOCK - 1) / SCAN_BLOCK;
    w.cap = H * W / 8 + 16;
    w.npbin = c.take<uint8_t>(BN); w.tmap = c.take<uint8_t>(BN); w.blb = c.take<uint8_t>(BN);
    w.mk = c.take<uint8_t>(BN); w.mk2 = c.take<uint8_t>(BN);
    w.L1 = c.take<int>(BN); w.cnt1 = c.take<int>(BN); w.off1 = c.take<int>(BN); w.fill1 = c.take<int>(BN); w.blobpix = c.take<int>(BN);
    w.Lx = c.take<int>(BN); w.cntx = c.take<int>(BN); w.flagx = c.take<int>(BN); w.rankx = c.take<int>(BN); w.marker = c.take<int>(BN);
    w.bsum = c.take<int>((size_t)B * nb + 256);
    w.qstride = (int)(BN / 8 + 64);
    w.queue = c.take<int>((size_t)(NQ_CLASSES + 1) * w.qstride);
    w.qmeta = c.take<int>(16); w.status = c.take<int>(2 * B + 32);  // status[B + 16] | maxid[B]
    w.mm = c.take<uint32_t>((size_t)B * 4 + 16); w.mm64 = c.take<unsigned long long>((size_t)B * 4 + 16);
    w.sob = c.take<double>(BN * 2); w.dist0 = c.take<double>(BN); w.dist = c.take<double>(BN);
    w.gkey = c.take<double>(BN); w.gpay = c.take<int2>(BN);
    w.acc = c.take<Acc>((size_t)B * w.cap);
    w.bytes = c.off + 4096;
    return w;
}

int g_post_max_ctas = 0;   // 0: up to 32 CTAs per SM (debug knob, cellvit_b200_debug.h)
int grid1d(long long n, int block = 256) {
    long long g = (n + block - 1) / block;
    const long long cap = g_post_max_ctas > 0 ? g_post_max_ctas : (long long)cvb_num_sms() * 32;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

int scan_i32(const int* in, int as_flag, Dims d, int* bsum, int* out, cudaStream_t st) {
    const int nb = (d.N + SCAN_BLOCK - 1) / SCAN_BLOCK;
    scan_reduce_kernel<<<dim3(nb, d.B), 256, 0, st>>>(in, as_flag, d.N, nb, bsum);
    scan_bsums_kernel<<<d.B, 256, 0, st>>>(nb, bsum);
    scan_apply_kernel<<<dim3(nb, d.B), 256, 0, st>>>(in, as_flag, d.N, nb, bsum, out);
    cvb_note_launches(3);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}

int ccl(const uint8_t* mask, int want, Dims d, int* L, cudaStream_t st) {
    const long long warps = (long long)d.B * d.H * ((d.W + 31) / 32);
    ccl_init_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(mask, want, d, L);
    ccl_merge_kernel<<<grid1d((long long)d.B * d.N), 256, 0, st>>>(mask, want, d, L);
    ccl_compress_kernel<<<grid1d((long long)d.B * d.N), 256, 0, st>>>(d, L);
    cvb_note_launches(3);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}

int run_pipeline(const Ws& w, const float* hv, Dims d, int n_types, int object_size, int ksize, bool have_types, int32_t* labels,
                 cvb_inst_row* table, int32_t* counts, int max_rows, uint8_t* dbg_blb, double* dbg_dist, int32_t* dbg_marker,
                 cudaStream_t st) {
    const size_t BN = (size_t)d.B * d.N;
    const int g = grid1d((long long)BN);
    // ---- P1/P2
    CVB_TRY(ccl(w.npbin, 1, d, w.L1, st));
    CVB_CUDA(cudaMemsetAsync(w.cnt1, 0, BN * 4, st));
    count_kernel<<<g, 256, 0, st>>>(w.L1, d, w.cnt1);
    blb_kernel<<<g, 256, 0, st>>>(w.L1, w.cnt1, d, 10, w.blb);
    // ---- P3/P4
    CVB_CUDA(cudaMemsetAsync(w.mm, 0, (size_t)d.B * 4 * 4, st));
    {
        // min slots start at 0xFFFFFFFF, max slots at 0: set min slots with a strided memset2D
        CVB_CUDA(cudaMemset2DAsync(w.mm, 8, 0xFF, 4, (size_t)d.B * 2, st));
        CVB_CUDA(cudaMemsetAsync(w.mm64, 0, (size_t)d.B * 4 * 8, st));
        CVB_CUDA(cudaMemset2DAsync(w.mm64, 16, 0xFF, 8, (size_t)d.B * 2, st));
    }
    minmax_f32_kernel<<<dim3(148, d.B * 2), 256, 0, st>>>(hv, d, w.mm);
    SobelTaps taps;
    sobel_taps_host(ksize, &taps);
    {
        const int r = ksize / 2;
        const size_t in_bytes = (((size_t)(SB_TH + 2 * r) * (SB_TW + 2 * r) * 4) + 15) & ~(size_t)15;
        const size_t smem = in_bytes + (size_t)(SB_TH + 2 * r) * SB_TW * 8;
        static unsigned long long configured = 0;  // one bit per device: function attributes are per device
    const int cfg_dev = cvb_current_device();
        if (!((configured >> cfg_dev) & 1ull)) {
            CVB_CUDA(cudaFuncSetAttribute(sobel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            configured |= 1ull << cfg_dev;
        }
        dim3 grid((d.W + SB_TW - 1) / SB_TW, (d.H + SB_TH - 1) / SB_TH, d.B * 2);
        sobel_kernel<<<grid, 256, smem, st>>>(hv, d, w.mm, taps, w.sob, w.mm64);
    }
    // ---- P5
    energy_kernel<<<g, 256, 0, st>>>(w.sob, w.mm64, w.blb, d, w.dist0, w.mk);
    blur_kernel<<<g, 256, 0, st>>>(w.dist0, d, w.dist);
    // ---- P6a: fill holes = background components that do not touch the border
    CVB_TRY(ccl(w.mk, 0, d, w.Lx, st));
    CVB_CUDA(cudaMemsetAsync(w.flagx, 0, BN * 4, st));
    border_flag_kernel<<<grid1d((long long)d.B * (2 * d.W + 2 * d.H)), 256, 0, st>>>(w.Lx, d, w.flagx);
    fill_kernel<<<g, 256, 0, st>>>(w.mk, w.Lx, w.flagx, d, w.mk2);
    // ---- P6b: opening
    morph5_kernel<<<g, 256, 0, st>>>(w.mk2, d, 0, w.mk);
    morph5_kernel<<<g, 256, 0, st>>>(w.mk, d, 1, w.mk2);
    // ---- P6c/d: label markers in raster order, drop small ones
    CVB_TRY(ccl(w.mk2, 1, d, w.Lx, st));
    CVB_CUDA(cudaMemsetAsync(w.cntx, 0, BN * 4, st));
    count_kernel<<<g, 256, 0, st>>>(w.Lx, d, w.cntx);
    root_flag_kernel<<<g, 256, 0, st>>>(w.Lx, d, w.flagx);
    CVB_TRY(scan_i32(w.flagx, 1, d, w.bsum, w.rankx, st));
    marker_kernel<<<g, 256, 0, st>>>(w.Lx, w.cntx, w.rankx, w.blb, d, object_size, w.marker, labels);
    // ---- P7: per-blob floods
    CVB_TRY(scan_i32(w.cnt1, 0, d, w.bsum, w.off1, st));
    CVB_CUDA(cudaMemsetAsync(w.fill1, 0, BN * 4, st));
    CVB_CUDA(cudaMemsetAsync(w.qmeta, 0, 16 * 4, st));
    blob_scatter_kernel<<<g, 256, 0, st>>>(w.L1, w.blb, w.off1, d, w.fill1, w.blobpix);
    blob_queue_kernel<<<g, 256, 0, st>>>(w.L1, w.cnt1, d, 10, FL_MAXN, w.qmeta, w.queue, w.qstride);
    {
        // Two launches of single-warp CTAs (disjoint blobs, so they run concurrently; fork / join with events keeps the sequence
        // capturable and ordered on the caller's stream):
        //   class 0 (more than 1023 pixels, ~5 % of the blobs, 15 % of the pixels): fp64 values and 16-byte heap entries of the
        //             staged bounding box in 200 KB of shared memory, one CTA per SM (~0.35 us per pop; a flood through L2 costs
        //             2 us per pop, and the largest blob of a tile IS the duration of the post-processing)
        //   classes 1..5: rank-transformed flood, 21 KB per CTA
        {
            static unsigned long long configured = 0;  // one bit per device: function attributes are per device
            const int cfg_dev = cvb_current_device();
            if (!((configured >> cfg_dev) & 1ull)) {
                CVB_CUDA(cudaFuncSetAttribute(watershed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FL_LARGE_SMEM));
                configured |= 1ull << cfg_dev;
            }
        }
        Fork* fk = get_fork();
        CVB_CHECK(fk != nullptr, CVB_ECUDA, "cvb_postproc: could not create the side stream");
        if (g_debug_skip_flood) goto after_flood;   // timing experiments only (cellvit_b200_debug.h): label maps are then incomplete
        CVB_CUDA(cudaEventRecord(fk->fork, st));
        CVB_CUDA(cudaStreamWaitEvent(fk->side, fk->fork, 0));
        watershed_kernel<<<g_flood_large_ctas > 0 ? g_flood_large_ctas : cvb_num_sms(), 32, FL_LARGE_SMEM, fk->side>>>(w.queue, w.qstride, w.qmeta, 0, 1, w.cnt1, w.off1, w.blobpix, w.blb, w.marker,
                                                                        w.dist, d, FL_LARGE_CAP, FL_LARGE_RCAP, w.gkey, w.gpay, labels);
        CVB_CUDA(cudaEventRecord(fk->join, fk->side));
        watershed_kernel<<<cvb_num_sms() * 6, 32, FL_SMEM, st>>>(w.queue, w.qstride, w.qmeta, 1, NQ_CLASSES, w.cnt1, w.off1, w.blobpix, w.blb,
                                                                 w.marker, w.dist, d, 0, 0, w.gkey, w.gpay, labels);
        CVB_CUDA(cudaStreamWaitEvent(st, fk->join, 0));
        // deferred blobs of the rank launch (usually none: the launch then costs a few microseconds)
        watershed_kernel<<<cvb_num_sms(), 32, FL_LARGE_SMEM, st>>>(w.queue, w.qstride, w.qmeta, Q_DEFER, Q_DEFER + 1, w.cnt1, w.off1, w.blobpix, w.blb,
                                                                   w.marker, w.dist, d, FL_LARGE_CAP, FL_LARGE_RCAP, w.gkey, w.gpay, labels);
    }
after_flood:
    // ---- P8/P9
    if (table && counts) {
        CVB_CUDA(cudaMemsetAsync(w.status, 0, (size_t)(2 * d.B + 16) * 4, st));  // status[B+16] and maxid[B]
        table_init_kernel<<<grid1d((long long)d.B * w.cap), 256, 0, st>>>(w.acc, (long long)d.B * w.cap, d.H, d.W);
        table_accum_kernel<<<g, 256, 0, st>>>(labels, have_types ? w.tmap : nullptr, d, w.cap, w.acc, w.status, w.status + d.B + 16);
        table_finalize_kernel<<<d.B, 256, 0, st>>>(w.acc, w.cap, w.status + d.B + 16, have_types ? n_types : 0, max_rows, table, counts);
    }
    cvb_note_launches(18 + ((table && counts) ? 3 : 0));
    if (dbg_blb) CVB_CUDA(cudaMemcpyAsync(dbg_blb, w.blb, BN, cudaMemcpyDeviceToDevice, st));
    if (dbg_dist) CVB_CUDA(cudaMemcpyAsync(dbg_dist, w.dist, BN * 8, cudaMemcpyDeviceToDevice, st));
    if (dbg_marker) CVB_CUDA(cudaMemcpyAsync(dbg_marker, w.marker, BN * 4, cudaMemcpyDeviceToDevice, st));
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}

// ------------------------------------------------------------------------------------------ cell tokens
// cell_detection.py:397-409: every cell's embedding is the mean of the z4 tokens under its bounding box,
// tokens[b, :, floor(rmin/P):ceil(rmax/P), floor(cmin/P):ceil(cmax/P)] averaged over the window. One block per cell row,
// threads over the embedding dimension (reads are strided by h*w in the reference NCHW token layout, but a window is
// only a handful of tokens; the write is coalesced).
__global__ void __launch_bounds__(256)
cell_tokens_kernel(const float* __restrict__ tokens, const cvb_inst_row* __restrict__ table, const int* __restrict__ counts, int D,
                   int th, int tw, int patch, int max_rows, float* __restrict__ out) {
    const int b = blockIdx.y;
    const int n = min(counts[b], max_rows);
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const cvb_inst_row r = table[(long long)b * max_rows + i];
        const int r0 = r.rmin / patch, r1 = min(th, (r.rmax + patch - 1) / patch);
        const int c0 = r.cmin / patch, c1 = min(tw, (r.cmax + patch - 1) / patch);
        const float inv = 1.0f / (float)((r1 - r0) * (c1 - c0));
        for (int dch = threadIdx.x; dch < D; dch += blockDim.x) {
            const float* t = tokens + ((long long)b * D + dch) * th * tw;
            float s = 0.0f;
            for (int y = r0; y < r1; ++y)
                for (int x = c0; x < c1; ++x) s += t[y * tw + x];
            out[((long long)b * max_rows + i) * D + dch] = s * inv;
        }
    }
}

int check_common(int B, int H, int W, int ksize, int max_rows, const void* ws, size_t ws_bytes, size_t need) {
    CVB_CHECK(B > 0 && H > 0 && W > 0, CVB_EARG, "cvb_postproc: empty shape");
    CVB_CHECK((long long)B * H * W < (1ll << 31), CVB_ESHAPE, "cvb_postproc: batch too large for 32-bit pixel indices");
    CVB_CHECK(ksize % 2 == 1 && ksize >= 3 && ksize <= 2 * SB_MAXR + 1, CVB_ESHAPE, "cvb_postproc: Sobel ksize %d not supported", ksize);
    CVB_CHECK(max_rows >= 0, CVB_EARG, "cvb_postproc: negative max_rows");
    CVB_CHECK(ws != nullptr && ((uintptr_t)ws & 255) == 0, CVB_EARG, "cvb_postproc: workspace must be non-null and 256-byte aligned");
    CVB_CHECK(ws_bytes >= need, CVB_EWORKSPACE, "cvb_postproc: workspace %zu < required %zu bytes", ws_bytes, need);
    return CVB_OK;
}

}  // namespace

#define CVB_API extern "C" __attribute__((visibility("default")))

CVB_API void cvb_debug_postproc_skip_flood(int on) { g_debug_skip_flood = on; }
CVB_API void cvb_debug_postproc_max_ctas(int n) { g_post_max_ctas = n; }
CVB_API void cvb_debug_flood_large_ctas(int n) { g_flood_large_ctas = n; }

CVB_API int cvb_postproc_workspace_bytes(int B, int H, int W, size_t* out) {
    CVB_CHECK(out && B > 0 && H > 0 && W > 0, CVB_EARG, "cvb_postproc_workspace_bytes: bad arguments");
    *out = carve(nullptr, B, H, W).bytes;
    return CVB_OK;
}

CVB_API int cvb_postproc(const float* np_map, const float* hv, const float* nt_map, int B, int H, int W, int n_types,
                         int magnification, int32_t* labels, cvb_inst_row* table, int32_t* counts, int max_rows,
                         void* workspace, size_t ws_bytes, void* stream) {
    CVB_CHECK(np_map && hv && labels, CVB_EARG, "cvb_postproc: null map");
    int object_size, ksize;
    if (magnification == 40) { object_size = 10; ksize = 21; }
    else if (magnification == 20) { object_size = 3; ksize = 11; }
    else { cvb_set_error("Unknown magnification"); return CVB_EARG; }
    CVB_CHECK(nt_map == nullptr || (n_types >= 1 && n_types <= 8), CVB_ESHAPE, "cvb_postproc: n_types must be in 1..8");
    size_t need = 0;
    CVB_TRY(cvb_postproc_workspace_bytes(B, H, W, &need));
    CVB_TRY(check_common(B, H, W, ksize, max_rows, workspace, ws_bytes, need));
    const Ws w = carve(workspace, B, H, W);
    const Dims d{B, H, W, H * W};
    cudaStream_t st = (cudaStream_t)stream;
    prep_float_kernel<<<grid1d((long long)B * d.N), 256, 0, st>>>(np_map, nt_map, n_types, d, w.npbin, w.tmap);
    cvb_note_launches(1);
    return run_pipeline(w, hv, d, n_types, object_size, ksize, nt_map != nullptr, labels, table, counts, max_rows, nullptr, nullptr,
                        nullptr, st);
}

CVB_API int cvb_postproc_argmax(const uint8_t* np_argmax, const float* hv, const uint8_t* nt_argmax, int B, int H, int W, int n_types,
                                int magnification, int32_t* labels, cvb_inst_row* table, int32_t* counts, int max_rows, void* workspace,
                                size_t ws_bytes, void* stream) {
    CVB_CHECK(np_argmax && hv && labels, CVB_EARG, "cvb_postproc_argmax: null map");
    int object_size, ksize;
    if (magnification == 40) { object_size = 10; ksize = 21; }
    else if (magnification == 20) { object_size = 3; ksize = 11; }
    else { cvb_set_error("Unknown magnification"); return CVB_EARG; }
    CVB_CHECK(nt_argmax == nullptr || (n_types >= 1 && n_types <= 8), CVB_ESHAPE, "cvb_postproc_argmax: n_types must be in 1..8");
    size_t need = 0;
    CVB_TRY(cvb_postproc_workspace_bytes(B, H, W, &need));
    CVB_TRY(check_common(B, H, W, ksize, max_rows, workspace, ws_bytes, need));
    Ws w = carve(workspace, B, H, W);
    // the caller's planes are used in place (read-only): no preparation pass, 1 B/px instead of 4 * (2 + n_types) B/px
    w.npbin = const_cast<uint8_t*>(np_argmax);
    if (nt_argmax) w.tmap = const_cast<uint8_t*>(nt_argmax);
    const Dims d{B, H, W, H * W};
    return run_pipeline(w, hv, d, n_types, object_size, ksize, nt_argmax != nullptr, labels, table, counts, max_rows, nullptr, nullptr,
                        nullptr, (cudaStream_t)stream);
}

CVB_API int cvb_postproc_maps(const uint8_t* np_bin, const float* hv, const int32_t* type_map, int B, int H, int W, int n_types,
                              int object_size, int ksize, int32_t* labels, cvb_inst_row* table, int32_t* counts, int max_rows,
                              uint8_t* dbg_blb, double* dbg_dist, int32_t* dbg_marker, void* workspace, size_t ws_bytes,
                              void* stream) {
    CVB_CHECK(np_bin && hv && labels, CVB_EARG, "cvb_postproc_maps: null map");
    CVB_CHECK(type_map == nullptr || (n_types >= 1 && n_types <= 8), CVB_ESHAPE, "cvb_postproc_maps: n_types must be in 1..8");
    size_t need = 0;
    CVB_TRY(cvb_postproc_workspace_bytes(B, H, W, &need));
    CVB_TRY(check_common(B, H, W, ksize, max_rows, workspace, ws_bytes, need));
    const Ws w = carve(workspace, B, H, W);
    const Dims d{B, H, W, H * W};
    cudaStream_t st = (cudaStream_t)stream;
    CVB_CUDA(cudaMemcpyAsync(w.npbin, np_bin, (size_t)B * d.N, cudaMemcpyDeviceToDevice, st));
    if (type_map) prep_maps_kernel<<<grid1d((long long)B * d.N), 256, 0, st>>>(type_map, (long long)B * d.N, w.tmap);
    return run_pipeline(w, hv, d, n_types, object_size, ksize, type_map != nullptr, labels, table, counts, max_rows, dbg_blb, dbg_dist,
                        dbg_marker, st);
}

CVB_API int cvb_contours_workspace_bytes(int B, int H, int W, size_t* out) {
    CVB_CHECK(out && B > 0 && H > 0 && W > 0, CVB_EARG, "cvb_contours_workspace_bytes: bad arguments");
    *out = align_up((size_t)B * H * W * 4, 256) + align_up((size_t)B * (H * W / 8 + 16) * 4, 256) + 4096;
    return CVB_OK;
}

// Contours of the instances listed in table[b, :counts[b]] traced on the device label maps.
// pts int16 (x,y) pairs [B, max_rows, max_pts]; npts int32 [B, max_rows]: number of points, or -1 when the instance
// must be resolved on the host (several 8-connected components, or more than max_pts points).
CVB_API int cvb_contours(const int32_t* labels, const cvb_inst_row* table, const int32_t* counts, int B, int H, int W, int max_rows,
                         int max_pts, int16_t* pts, int32_t* npts, void* workspace, size_t ws_bytes, void* stream) {
    CVB_CHECK(labels && table && counts && pts && npts, CVB_EARG, "cvb_contours: null argument");
    CVB_CHECK(H < 32768 && W < 32768 && max_rows > 0 && max_pts > 0, CVB_ESHAPE, "cvb_contours: bad shape");
    const Dims d{B, H, W, H * W};
    const int cap = H * W / 8 + 16;
    if (workspace == nullptr) {
        // the caller vouches that every id is ONE 8-connected component -- true for label maps written by cvb_postproc: a marker
        // is a 4-connected component and the flood only ever labels 4-neighbours of labelled pixels -- so the 8-connected
        // labelling that finds multi-component ids (a quarter of the contour stage's time) is skipped
        contour_kernel<<<dim3((max_rows + 63) / 64, B), 64, 0, (cudaStream_t)stream>>>(labels, table, counts, nullptr, d, cap, max_rows, max_pts,
                                                                                     reinterpret_cast<short2*>(pts), npts);
        cvb_note_launches(1);
        CVB_CUDA(cudaGetLastError());
        return CVB_OK;
    }
    size_t need = 0;
    CVB_TRY(cvb_contours_workspace_bytes(B, H, W, &need));
    CVB_CHECK(ws_bytes >= need && ((uintptr_t)workspace & 255) == 0, CVB_EWORKSPACE, "cvb_contours: workspace %zu < %zu or unaligned", ws_bytes, need);
    int* L = reinterpret_cast<int*>(workspace);
    int* ncomp = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(workspace) + align_up((size_t)B * d.N * 4, 256));
    cudaStream_t st = (cudaStream_t)stream;
    const int g = grid1d((long long)B * d.N);
    CVB_CUDA(cudaMemsetAsync(ncomp, 0, (size_t)B * cap * 4, st));
    lab8_init_kernel<<<g, 256, 0, st>>>(labels, d, L);
    lab8_merge_kernel<<<g, 256, 0, st>>>(labels, d, L);
    ccl_compress_kernel<<<g, 256, 0, st>>>(d, L);
    lab8_count_kernel<<<g, 256, 0, st>>>(labels, L, d, cap, ncomp);
    contour_kernel<<<dim3((max_rows + 63) / 64, B), 64, 0, st>>>(labels, table, counts, ncomp, d, cap, max_rows, max_pts,
                                                               reinterpret_cast<short2*>(pts), npts);
    cvb_note_launches(5);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}

CVB_API int cvb_cell_tokens(const float* tokens, const cvb_inst_row* table, const int32_t* counts, int B, int D, int th, int tw,
                            int patch, int max_rows, float* out, void* stream) {
    CVB_CHECK(tokens && table && counts && out, CVB_EARG, "cvb_cell_tokens: null argument");
    CVB_CHECK(B > 0 && D > 0 && th > 0 && tw > 0 && patch > 0 && max_rows > 0, CVB_EARG, "cvb_cell_tokens: empty shape");
    const int gx = max_rows < 4096 ? max_rows : 4096;
    cell_tokens_kernel<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(tokens, table, counts, D, th, tw, patch, max_rows, out);
    cvb_note_launches(1);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}
