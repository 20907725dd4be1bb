/*
 * oracle/postproc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded CPU restatement of the reference's HoVer-Net
 * post-processing for one tile (the checker for the CUDA path). Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline/--impl reference legs
 * may load this; the product path (cellvit_b200/) never does.
 *
 * Reference followed (file:line relative to the reference tree):
 *   cell_segmentation/utils/post_proc_cellvit.py:155-249  (__proc_np_hv, stages P1..P7)
 *   cell_segmentation/utils/post_proc_cellvit.py:95-151   (per-instance table, P8/P9)
 *   cell_segmentation/utils/tools.py:24-34, 61-101        (bbox, remove_small_objects)
 * Third-party arithmetic restated here because the reference calls into it:
 *   OpenCV 4.x  cv2.normalize / cv2.Sobel / cv2.GaussianBlur / cv2.morphologyEx
 *               (operation order pinned against cv2 4.13 in tests/test_oracle_postproc.py)
 *   SciPy       ndimage.label (4-connectivity, raster-order ids), binary_fill_holes
 *   scikit-image==0.19.3 (requirements.txt:21) skimage.segmentation.watershed --
 *               NOT installed and not vendored: PARITY UNPINNED for this one stage.
 *               Restated from the published algorithm (_watershed_cy.pyx): seeds pushed
 *               in raster order with age 0; pop order = (value, age) with ties broken
 *               here by a strict total order (value, age, pixel index); a neighbour is
 *               labelled at push time; neighbours visited up, left, right, down.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -shared -fPIC).
 * -ffp-contract=off matters: OpenCV's double filters run without FMA contraction.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CVO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ helpers */

static inline int reflect101(int p, int len) {
    /* cv2.BORDER_REFLECT_101: gfedcb|abcdefgh|gfedcba */
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * len - 2 - p;
    }
    return p;
}

/* 4-connected labelling, ids numbered by raster order of each component's first
 * pixel (scipy.ndimage.label with the default cross structure). Returns count. */
static int label4(const uint8_t *bin, int H, int W, int32_t *out, int32_t *stack) {
    int n = 0;
    memset(out, 0, sizeof(int32_t) * (size_t)H * W);
    for (int p0 = 0; p0 < H * W; ++p0) {
        if (!bin[p0] || out[p0]) continue;
        ++n;
        int sp = 0;
        stack[sp++] = p0;
        out[p0] = n;
        while (sp) {
            int p = stack[--sp];
            int y = p / W, x = p - y * W;
            if (y > 0 && bin[p - W] && !out[p - W]) { out[p - W] = n; stack[sp++] = p - W; }
            if (x > 0 && bin[p - 1] && !out[p - 1]) { out[p - 1] = n; stack[sp++] = p - 1; }
            if (x < W - 1 && bin[p + 1] && !out[p + 1]) { out[p + 1] = n; stack[sp++] = p + 1; }
            if (y < H - 1 && bin[p + W] && !out[p + W]) { out[p + W] = n; stack[sp++] = p + W; }
        }
    }
    return n;
}

/* tools.py:61-101: zero every label whose pixel count is < min_size. */
static void remove_small(int32_t *lab, int n_lab, int HW, int min_size) {
    if (min_size == 0) return;
    int32_t *cnt = (int32_t *)calloc((size_t)n_lab + 1, sizeof(int32_t));
    for (int p = 0; p < HW; ++p) cnt[lab[p]]++;
    for (int p = 0; p < HW; ++p)
        if (cnt[lab[p]] < min_size) lab[p] = 0;
    free(cnt);
}

/* cv2.normalize(alpha=0, beta=1, NORM_MINMAX, dtype=CV_32F) parameters as OpenCV 4.x computes them
 * (pinned against cv2 4.13): scale = (float)(1/(max-min)) (0 if max-min <= DBL_EPSILON),
 * shift = (float)0 - (float)(min * (double)scale). */
static void minmax_params(double mn, double mx, float *scale, float *shift) {
    double d = mx - mn;
    double sc = d > DBL_EPSILON ? 1.0 / d : 0.0;
    *scale = (float)sc;
    *shift = 0.0f - (float)(mn * (double)*scale);
}

/* f32 -> f32 convertTo: one fused multiply-add in float. */
static void minmax_norm_f32(const float *src, float *dst, int n) {
    double mn = src[0], mx = src[0];
    for (int i = 1; i < n; ++i) {
        if (src[i] < mn) mn = src[i];
        if (src[i] > mx) mx = src[i];
    }
    float a, b;
    minmax_params(mn, mx, &a, &b);
    for (int i = 0; i < n; ++i) dst[i] = fmaf(src[i], a, b);
}

/* f64 -> f32 convertTo: dst = (float)(src*(double)scale + (double)shift), product and sum rounded
 * separately in double. */
static void minmax_norm_f64_to_f32(const double *src, float *dst, int n) {
    double mn = src[0], mx = src[0];
    for (int i = 1; i < n; ++i) {
        if (src[i] < mn) mn = src[i];
        if (src[i] > mx) mx = src[i];
    }
    float af, bf;
    minmax_params(mn, mx, &af, &bf);
    double a = (double)af, b = (double)bf;
    for (int i = 0; i < n; ++i) {
        volatile double prod = src[i] * a;
        dst[i] = (float)(prod + b);
    }
}

/* cv::getSobelKernels for ksize > 7 (integer recurrences), order 0 or 1. */
static void sobel_kernel(int ksize, int order, double *k) {
    long long ker[64];
    memset(ker, 0, sizeof(ker));
    ker[0] = 1;
    for (int i = 0; i < ksize - order - 1; ++i) {
        long long oldv = ker[0];
        for (int j = 1; j <= ksize; ++j) {
            long long newv = ker[j] + ker[j - 1];
            ker[j - 1] = oldv;
            oldv = newv;
        }
    }
    for (int i = 0; i < order; ++i) {
        long long oldv = -ker[0];
        for (int j = 1; j <= ksize; ++j) {
            long long newv = ker[j - 1] - ker[j];
            ker[j - 1] = oldv;
            oldv = newv;
        }
    }
    for (int j = 0; j < ksize; ++j) k[j] = (double)ker[j];
}

/* cv2.sepFilter2D restated: generic row filter (ascending taps, separate mul/add),
 * then symmetric / anti-symmetric column filter (centre first, then +-j pairs).
 * src may be float (is_f32) or double. Border REFLECT_101. */
static void sep_filter(const void *src, int is_f32, int H, int W, const double *kx, const double *ky,
                       int ksize, int ky_antisym, double *tmp, double *dst) {
    int r = ksize / 2;
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            double acc = 0.0;
            for (int j = 0; j < ksize; ++j) {
                int xx = reflect101(x + j - r, W);
                double s = is_f32 ? (double)((const float *)src)[y * W + xx]
                                  : ((const double *)src)[y * W + xx];
                volatile double prod = kx[j] * s;
                acc = (j == 0) ? prod : acc + prod;
            }
            tmp[y * W + x] = acc;
        }
    }
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            double acc;
            if (!ky_antisym) {
                volatile double p0 = ky[r] * tmp[y * W + x];
                acc = p0;
                for (int j = 1; j <= r; ++j) {
                    double a = tmp[reflect101(y + j, H) * W + x];
                    double b = tmp[reflect101(y - j, H) * W + x];
                    volatile double s = a + b;
                    volatile double p = ky[r + j] * s;
                    acc += p;
                }
            } else {
                acc = 0.0;
                for (int j = 1; j <= r; ++j) {
                    double a = tmp[reflect101(y + j, H) * W + x];
                    double b = tmp[reflect101(y - j, H) * W + x];
                    volatile double s = a - b;
                    volatile double p = ky[r + j] * s;
                    acc += p;
                }
            }
            dst[y * W + x] = acc;
        }
    }
}

/* scipy.ndimage.binary_fill_holes, default structure: a background pixel is a hole
 * unless it is 4-connected to the image border through background. */
static void fill_holes(uint8_t *m, int H, int W, int32_t *stack) {
    uint8_t *reach = (uint8_t *)calloc((size_t)H * W, 1);
    int sp = 0;
#define SEED(p)                                  \
    if (!m[p] && !reach[p]) {                    \
        reach[p] = 1;                            \
        stack[sp++] = (p);                       \
    }
    for (int x = 0; x < W; ++x) { SEED(x); SEED((H - 1) * W + x); }
    for (int y = 0; y < H; ++y) { SEED(y * W); SEED(y * W + W - 1); }
    while (sp) {
        int p = stack[--sp];
        int y = p / W, x = p - y * W;
        if (y > 0) SEED(p - W);
        if (x > 0) SEED(p - 1);
        if (x < W - 1) SEED(p + 1);
        if (y < H - 1) SEED(p + W);
    }
#undef SEED
    for (int p = 0; p < H * W; ++p)
        if (!reach[p]) m[p] = 1;
    free(reach);
}

/* cv2.morphologyEx(MORPH_OPEN, getStructuringElement(MORPH_ELLIPSE,(5,5))):
 * erode then dilate; out-of-image taps are ignored by both passes. */
static const int8_t ELL5[5][5] = {{0, 0, 1, 0, 0}, {1, 1, 1, 1, 1}, {1, 1, 1, 1, 1}, {1, 1, 1, 1, 1}, {0, 0, 1, 0, 0}};
static void morph5(const uint8_t *src, uint8_t *dst, int H, int W, int dilate) {
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            uint8_t v = dilate ? 0 : 1;
            for (int dy = -2; dy <= 2; ++dy)
                for (int dx = -2; dx <= 2; ++dx) {
                    if (!ELL5[dy + 2][dx + 2]) continue;
                    int yy = y + dy, xx = x + dx;
                    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                    uint8_t s = src[yy * W + xx];
                    if (dilate) v |= s; else v &= s;
                }
            dst[y * W + x] = v;
        }
}

/* ------------------------------------------------------------ watershed (P7) */
typedef struct { double v; int32_t age; int32_t idx; } hitem;
static inline int hless(const hitem *a, const hitem *b) {
    if (a->v != b->v) return a->v < b->v;
    if (a->age != b->age) return a->age < b->age;
    return a->idx < b->idx;
}
static void hpush(hitem *h, int *n, hitem it) {
    int i = (*n)++;
    while (i > 0) {
        int par = (i - 1) >> 1;
        if (!hless(&it, &h[par])) break;
        h[i] = h[par];
        i = par;
    }
    h[i] = it;
}
static hitem hpop(hitem *h, int *n) {
    hitem top = h[0];
    hitem last = h[--(*n)];
    int i = 0, m = *n;
    for (;;) {
        int c = 2 * i + 1;
        if (c >= m) break;
        if (c + 1 < m && hless(&h[c + 1], &h[c])) ++c;
        if (!hless(&h[c], &last)) break;
        h[i] = h[c];
        i = c;
    }
    if (m > 0) h[i] = last;
    return top;
}

CVO_API void cvo_watershed(const double *image, const int32_t *markers, const uint8_t *mask, int H, int W,
                           int32_t *out) {
    int HW = H * W;
    hitem *heap = (hitem *)malloc(sizeof(hitem) * (size_t)(HW + 1));
    int hn = 0;
    int32_t age = 0;
    for (int p = 0; p < HW; ++p) out[p] = mask[p] ? markers[p] : 0;
    for (int p = 0; p < HW; ++p)
        if (out[p]) { hitem it = {image[p], 0, p}; hpush(heap, &hn, it); }
    while (hn) {
        hitem e = hpop(heap, &hn);
        int p = e.idx, y = p / W, x = p - y * W;
        int nb[4]; int nn = 0;
        if (y > 0) nb[nn++] = p - W;
        if (x > 0) nb[nn++] = p - 1;
        if (x < W - 1) nb[nn++] = p + 1;
        if (y < H - 1) nb[nn++] = p + W;
        for (int k = 0; k < nn; ++k) {
            int q = nb[k];
            if (!mask[q] || out[q]) continue;
            ++age;
            out[q] = out[p];
            hitem it = {image[q], age, q};
            hpush(heap, &hn, it);
        }
    }
    free(heap);
}

/* --------------------------------------------------------- full tile pipeline */

typedef struct {
    int32_t id;
    int32_t rmin, cmin, rmax, cmax; /* max exclusive (tools.py:31-33) */
    int32_t area;
    int32_t type;
    float type_prob_f;  /* unused padding keeps the row layout equal to the device table */
    double cx, cy;      /* centroid x,y = m10/m00 + cmin, m01/m00 + rmin (post_proc_cellvit.py:117-125) */
    double type_prob;   /* count/(area+1e-6) (post_proc_cellvit.py:149) */
    int32_t hist[8];
} cvo_inst_row;

/* Runs P1..P7 on one tile.
 *  np_bin  [H*W] u8  : argmax of the NP head (0/1)         (cellvit.py:372)
 *  hv      [2*H*W] f32: hv_map planes, h (x-map) then v     (cellvit.py:375)
 *  outputs (any may be NULL): labels int32, blb u8, dist f64, marker int32
 * magnification semantic is pre-resolved by the caller into object_size and ksize
 * (post_proc_cellvit.py:55-65). */
CVO_API int cvo_proc_np_hv(const uint8_t *np_bin, const float *hv, int H, int W, int object_size, int ksize,
                           int32_t *labels, uint8_t *blb_out, double *dist_out, int32_t *marker_out) {
    if (ksize > 31 || (ksize & 1) == 0 || H < 1 || W < 1) return -1;
    int HW = H * W;
    int32_t *stack = (int32_t *)malloc(sizeof(int32_t) * (size_t)HW);
    int32_t *lab = (int32_t *)malloc(sizeof(int32_t) * (size_t)HW);
    uint8_t *blb = (uint8_t *)malloc((size_t)HW);
    float *hn = (float *)malloc(sizeof(float) * (size_t)HW);
    float *vn = (float *)malloc(sizeof(float) * (size_t)HW);
    double *tmp = (double *)malloc(sizeof(double) * (size_t)HW);
    double *sh = (double *)malloc(sizeof(double) * (size_t)HW);
    double *sv = (double *)malloc(sizeof(double) * (size_t)HW);
    float *shn = (float *)malloc(sizeof(float) * (size_t)HW);
    float *svn = (float *)malloc(sizeof(float) * (size_t)HW);
    double *overall = (double *)malloc(sizeof(double) * (size_t)HW);
    double *dist = (double *)malloc(sizeof(double) * (size_t)HW);
    uint8_t *mk = (uint8_t *)malloc((size_t)HW);
    uint8_t *mk2 = (uint8_t *)malloc((size_t)HW);

    /* P1/P2  :179-183 */
    int n = label4(np_bin, H, W, lab, stack);
    remove_small(lab, n, HW, 10);
    for (int p = 0; p < HW; ++p) blb[p] = lab[p] > 0;

    /* P3  :185-200 */
    minmax_norm_f32(hv, hn, HW);
    minmax_norm_f32(hv + HW, vn, HW);

    /* P4  :205-206  Sobel(h,dx=1) and Sobel(v,dy=1), CV_64F */
    double kd[32], ks[32];
    sobel_kernel(ksize, 1, kd);
    sobel_kernel(ksize, 0, ks);
    sep_filter(hn, 1, H, W, kd, ks, ksize, 0, tmp, sh);
    sep_filter(vn, 1, H, W, ks, kd, ksize, 1, tmp, sv);

    /* P5  :208-235 */
    minmax_norm_f64_to_f32(sh, shn, HW);
    minmax_norm_f64_to_f32(sv, svn, HW);
    for (int p = 0; p < HW; ++p) {
        float a = 1.0f - shn[p], b = 1.0f - svn[p];
        float m = a > b ? a : b;                 /* np.maximum */
        double o = (double)m - (double)(1 - (int)blb[p]);
        if (o < 0) o = 0;
        overall[p] = o;
        dist[p] = (1.0 - o) * (double)blb[p];
    }
    {
        const double g[3] = {0.25, 0.5, 0.25};
        sep_filter(dist, 0, H, W, g, g, 3, 0, tmp, sh);
        for (int p = 0; p < HW; ++p) dist[p] = -sh[p];
    }

    /* P6  :237-245 */
    for (int p = 0; p < HW; ++p) {
        int ov = overall[p] >= 0.4;
        int m = (int)blb[p] - ov;
        mk[p] = m > 0;
    }
    fill_holes(mk, H, W, stack);
    morph5(mk, mk2, H, W, 0);
    morph5(mk2, mk, H, W, 1);
    n = label4(mk, H, W, lab, stack);
    remove_small(lab, n, HW, object_size);

    /* P7  :247 */
    if (marker_out) memcpy(marker_out, lab, sizeof(int32_t) * (size_t)HW);
    if (blb_out) memcpy(blb_out, blb, (size_t)HW);
    if (dist_out) memcpy(dist_out, dist, sizeof(double) * (size_t)HW);
    if (labels) cvo_watershed(dist, lab, blb, H, W, labels);

    free(stack); free(lab); free(blb); free(hn); free(vn); free(tmp); free(sh); free(sv);
    free(shn); free(svn); free(overall); free(dist); free(mk); free(mk2);
    return 0;
}

/* P8/P9 (post_proc_cellvit.py:95-151) minus the contour (cv2.findContours stays a
 * host call on the bbox crop, both in the reference and in the product's interim path).
 * Rows come out in ascending id order. Honours the reference quirk that
 * np.unique(label)[1:] drops the smallest value unconditionally (:95).
 * Returns the number of rows, or -needed if max_rows is too small. */
CVO_API int cvo_instance_table(const int32_t *labels, const int32_t *type_map, int H, int W, int nr_types,
                               cvo_inst_row *rows, int max_rows) {
    int HW = H * W;
    int32_t maxid = 0, minval = INT32_MAX;
    for (int p = 0; p < HW; ++p) {
        if (labels[p] > maxid) maxid = labels[p];
        if (labels[p] < minval) minval = labels[p];
    }
    cvo_inst_row *acc = (cvo_inst_row *)calloc((size_t)maxid + 1, sizeof(cvo_inst_row));
    int64_t *sx = (int64_t *)calloc((size_t)maxid + 1, sizeof(int64_t));
    int64_t *sy = (int64_t *)calloc((size_t)maxid + 1, sizeof(int64_t));
    for (int i = 0; i <= maxid; ++i) { acc[i].rmin = H; acc[i].cmin = W; acc[i].rmax = -1; acc[i].cmax = -1; }
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int32_t l = labels[y * W + x];
            cvo_inst_row *r = &acc[l];
            r->area++;
            if (y < r->rmin) r->rmin = y;
            if (y > r->rmax) r->rmax = y;
            if (x < r->cmin) r->cmin = x;
            if (x > r->cmax) r->cmax = x;
            sx[l] += x; sy[l] += y;
            if (type_map) {
                int t = type_map[y * W + x];
                if (t >= 0 && t < 8) r->hist[t]++;
            }
        }
    int nrows = 0;
    for (int32_t l = 0; l <= maxid; ++l) {
        if (acc[l].area == 0 || l == minval) continue; /* [1:] drops the smallest unique value */
        if (nrows < max_rows) {
            cvo_inst_row r = acc[l];
            r.id = l;
            r.rmax += 1; r.cmax += 1;
            /* cv2.moments on the bbox crop: integer sums, then m10/m00 in double, then + offset */
            double m00 = (double)r.area;
            r.cx = (double)(sx[l] - (int64_t)r.cmin * r.area) / m00 + (double)r.cmin;
            r.cy = (double)(sy[l] - (int64_t)r.rmin * r.area) / m00 + (double)r.rmin;
            /* :141-151 majority vote; ties -> smaller class id; 0 yields to the runner-up */
            int best = -1, second = -1;
            for (int t = 0; t < nr_types && t < 8; ++t) {
                if (r.hist[t] == 0) continue;
                if (best < 0 || r.hist[t] > r.hist[best]) { second = best; best = t; }
                else if (second < 0 || r.hist[t] > r.hist[second]) second = t;
            }
            int ty = best;
            if (ty == 0 && second >= 0) ty = second;
            r.type = ty;
            r.type_prob = ty >= 0 ? (double)r.hist[ty] / ((double)r.area + 1.0e-6) : 0.0;
            r.type_prob_f = (float)r.type_prob;
            rows[nrows] = r;
        }
        ++nrows;
    }
    free(acc); free(sx); free(sy);
    return nrows <= max_rows ? nrows : -nrows;
}

CVO_API int cvo_row_bytes(void) { return (int)sizeof(cvo_inst_row); }
