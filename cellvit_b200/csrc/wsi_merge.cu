// wsi_merge.cu -- exact pairwise polygon overlap for the WSI-level duplicate removal.
//
// Reference behaviour: CellPostProcessor._remove_overlap (cell_segmentation/inference/cell_detection.py:687-767) builds a
// shapely Polygon per margin cell, queries an STRtree for envelope hits and tests
//     intersection(a, b).area / a.area > 0.01  or  intersection(a, b).area / b.area > 0.01.
// shapely (GEOS) is a third-party dependency that is not installed here, so this is a restatement of the geometric
// quantity, not of GEOS: area(A ∩ B) of two simple polygons under the even-odd rule, computed exactly (up to fp64
// rounding) by slab decomposition -- between two consecutive critical abscissae (vertices of either polygon and
// edge-edge crossings) the chord length |A_x ∩ B_x| is linear in x, so its integral is width * value at the midpoint.
// One thread per candidate pair (contours have <= 128 points; the pair list comes from a bounding-box hash on the host).
#include "../../include/cellvit_b200.h"
#include "common.cuh"

namespace {

constexpr int PO_MAXP = 128;   // points per polygon
constexpr int PO_MAXX = 640;   // critical abscissae per pair (2 * 128 vertices + crossings)
constexpr int PO_MAXY = 24;    // boundary crossings of one polygon with one vertical line

__device__ __forceinline__ void sort_small(double* a, int n) {
    for (int i = 1; i < n; ++i) {
        const double v = a[i];
        int j = i - 1;
        while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; --j; }
        a[j + 1] = v;
    }
}

// crossings of the vertical line x = xm with the polygon boundary (xm is never a vertex abscissa)
__device__ __forceinline__ int crossings(const double* px, const double* py, int n, double xm, double* ys, bool* overflow) {
    int k = 0;
    for (int i = 0; i < n; ++i) {
        const int j = i + 1 == n ? 0 : i + 1;
        const double x1 = px[i], x2 = px[j];
        if ((x1 < xm) != (x2 < xm)) {
            if (k >= PO_MAXY) { *overflow = true; return k; }
            ys[k++] = py[i] + (xm - x1) * (py[j] - py[i]) / (x2 - x1);
        }
    }
    sort_small(ys, k);
    return k;
}

__global__ void __launch_bounds__(64)
polygon_overlap_kernel(const double* __restrict__ pts, const int* __restrict__ off, const int* __restrict__ pairs, int n_pairs,
                       double* __restrict__ inter) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pairs) return;
    const int ia = pairs[2 * t], ib = pairs[2 * t + 1];
    const int na = off[ia + 1] - off[ia], nb = off[ib + 1] - off[ib];
    if (na < 3 || nb < 3 || na > PO_MAXP || nb > PO_MAXP) { inter[t] = na < 3 || nb < 3 ? 0.0 : -1.0; return; }
    double ax[PO_MAXP], ay[PO_MAXP], bx[PO_MAXP], by[PO_MAXP], xs[PO_MAXX];
    // coordinates relative to A's first vertex: global WSI coordinates reach 1e5, contours span tens of pixels
    const double ox = pts[2 * (long long)off[ia]], oy = pts[2 * (long long)off[ia] + 1];
    int nx = 0;
    for (int i = 0; i < na; ++i) { ax[i] = pts[2 * (long long)(off[ia] + i)] - ox; ay[i] = pts[2 * (long long)(off[ia] + i) + 1] - oy; xs[nx++] = ax[i]; }
    for (int i = 0; i < nb; ++i) { bx[i] = pts[2 * (long long)(off[ib] + i)] - ox; by[i] = pts[2 * (long long)(off[ib] + i) + 1] - oy; xs[nx++] = bx[i]; }
    bool overflow = false;
    for (int i = 0; i < na && !overflow; ++i) {
        const int i2 = i + 1 == na ? 0 : i + 1;
        const double px = ax[i], py = ay[i], rx = ax[i2] - px, ry = ay[i2] - py;
        const double alx = fmin(px, ax[i2]), ahx = fmax(px, ax[i2]), aly = fmin(py, ay[i2]), ahy = fmax(py, ay[i2]);
        for (int j = 0; j < nb; ++j) {
            const int j2 = j + 1 == nb ? 0 : j + 1;
            const double qx = bx[j], qy = by[j], sx = bx[j2] - qx, sy = by[j2] - qy;
            if (fmax(qx, bx[j2]) < alx || fmin(qx, bx[j2]) > ahx || fmax(qy, by[j2]) < aly || fmin(qy, by[j2]) > ahy) continue;
            const double den = rx * sy - ry * sx;
            if (den == 0.0) continue;  // parallel / collinear edges do not change the interleaving order
            const double tt = ((qx - px) * sy - (qy - py) * sx) / den;
            const double uu = ((qx - px) * ry - (qy - py) * rx) / den;
            if (tt >= 0.0 && tt <= 1.0 && uu >= 0.0 && uu <= 1.0) {
                if (nx >= PO_MAXX) { overflow = true; break; }
                xs[nx++] = px + tt * rx;
            }
        }
    }
    if (overflow) { inter[t] = -1.0; return; }
    sort_small(xs, nx);
    double area = 0.0;
    double ya[PO_MAXY], yb[PO_MAXY];
    for (int s = 0; s + 1 < nx; ++s) {
        const double x0 = xs[s], x1 = xs[s + 1];
        if (!(x1 > x0)) continue;
        const double xm = 0.5 * (x0 + x1);
        if (!(xm > x0 && xm < x1)) continue;  // slab thinner than one ulp
        const int ka = crossings(ax, ay, na, xm, ya, &overflow);
        const int kb = crossings(bx, by, nb, xm, yb, &overflow);
        if (overflow) { inter[t] = -1.0; return; }
        // total length of the intersection of the even-odd interval sets [ya0,ya1] u [ya2,ya3] ... and likewise for B
        double len = 0.0;
        int i = 0, j = 0;
        while (i + 1 < ka && j + 1 < kb) {
            const double lo = fmax(ya[i], yb[j]), hi = fmin(ya[i + 1], yb[j + 1]);
            if (hi > lo) len += hi - lo;
            if (ya[i + 1] < yb[j + 1]) i += 2; else j += 2;
        }
        area += len * (x1 - x0);
    }
    inter[t] = area;
}

__global__ void polygon_area_kernel(const double* __restrict__ pts, const int* __restrict__ off, int n_poly, double* __restrict__ area) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_poly) return;
    const int o = off[t], n = off[t + 1] - o;
    if (n < 3) { area[t] = 0.0; return; }
    const double ox = pts[2 * (long long)o], oy = pts[2 * (long long)o + 1];
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        const int j = i + 1 == n ? 0 : i + 1;
        const double x1 = pts[2 * (long long)(o + i)] - ox, y1 = pts[2 * (long long)(o + i) + 1] - oy;
        const double x2 = pts[2 * (long long)(o + j)] - ox, y2 = pts[2 * (long long)(o + j) + 1] - oy;
        s += x1 * y2 - x2 * y1;
    }
    area[t] = 0.5 * fabs(s);
}

}  // namespace

#define CVB_API extern "C" __attribute__((visibility("default")))

CVB_API int cvb_polygon_overlap(const double* pts_xy, const int32_t* poly_off, int n_poly, const int32_t* pairs, int n_pairs,
                                double* poly_area, double* inter_area, void* stream) {
    CVB_CHECK(pts_xy && poly_off && n_poly >= 0 && n_pairs >= 0, CVB_EARG, "cvb_polygon_overlap: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (poly_area && n_poly > 0) polygon_area_kernel<<<(n_poly + 255) / 256, 256, 0, st>>>(pts_xy, poly_off, n_poly, poly_area);
    if (n_pairs > 0) {
        CVB_CHECK(pairs && inter_area, CVB_EARG, "cvb_polygon_overlap: null pair list or output");
        polygon_overlap_kernel<<<(n_pairs + 63) / 64, 64, 0, st>>>(pts_xy, poly_off, pairs, n_pairs, inter_area);
    }
    cvb_note_launches(2);
    CVB_CUDA(cudaGetLastError());
    return CVB_OK;
}
