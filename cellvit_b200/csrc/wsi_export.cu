// wsi_export.cu -- host-only: streams the WSI-level result files (cells.json, cell_detection.json, the two GeoJSON files)
// from the columnar per-slide cell store, replacing the per-cell Python dicts + json encoder of the reference's export
// (cell_segmentation/inference/cell_detection.py:352-409 per-cell records, :438-475 dumps, :538-597 convert_geojson).
// No CUDA in this file; it lives in the same library so that the WSI entry point has one native dependency.
//
// Output format = Python's json.dumps(obj, indent=indent) byte for byte (tests/test_wsi_export.py): ", " / ": " separators
// when compact, one element per line when indented, floats as float.__repr__ (shortest round-trip digits, exponent form
// outside 1e-4 <= |x| < 1e16), NaN / Infinity spelled as json does. The reference writes through ujson(indent=2); files are
// equal as parsed JSON, which is what its consumers (QuPath import, cell graph tooling) read.
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/cellvit_b200.h"
#include "common.cuh"

namespace {

class JsonOut {
  public:
    JsonOut(FILE* f, int indent) : f_(f), indent_(indent) { buf_.reserve(CAP + 4096); }
    bool ok() const { return ok_; }
    void flush() {
        if (!buf_.empty() && ok_) ok_ = fwrite(buf_.data(), 1, buf_.size(), f_) == buf_.size();
        buf_.clear();
    }
    void raw(const char* s, size_t n) { buf_.append(s, n); if (buf_.size() >= CAP) flush(); }
    void raw(const char* s) { raw(s, strlen(s)); }
    void set_depth(int d) { depth_ = d; }

    // containers: `first` is the caller's per-container flag
    void open(char c, bool& first) { buf_.push_back(c); ++depth_; first = true; }
    void close(char c, bool first) {
        --depth_;
        if (!first && indent_ >= 0) newline();
        buf_.push_back(c);
        if (buf_.size() >= CAP) flush();
    }
    // separator before an element (or a key) of the current container
    void item(bool& first) {
        if (first) { if (indent_ >= 0) newline(); }
        else if (indent_ >= 0) { buf_.push_back(','); newline(); }
        else { buf_.push_back(','); buf_.push_back(' '); }
        first = false;
    }
    void key(const char* k, bool& first) { item(first); buf_.push_back('"'); buf_.append(k); buf_.append("\": "); }

    void i64(long long v) {
        char t[24];
        auto r = std::to_chars(t, t + sizeof(t), v);
        buf_.append(t, r.ptr - t);
    }
    void boolean(bool b) { buf_.append(b ? "true" : "false"); }
    void null() { buf_.append("null"); }
    // float.__repr__ (Python/dtoa.c mode 0 + format_float_short 'r'): shortest digits that round-trip; fixed notation for
    // decimal exponents in [-4, 16), otherwise d[.ddd]e+XX with at least two exponent digits
    void f64(double v) {
        if (std::isnan(v)) { buf_.append("NaN"); return; }
        if (std::isinf(v)) { buf_.append(v < 0 ? "-Infinity" : "Infinity"); return; }
        if (v == 0.0) { buf_.append(std::signbit(v) ? "-0.0" : "0.0"); return; }
        char t[40];
        auto r = std::to_chars(t, t + sizeof(t), v, std::chars_format::scientific);  // [-]d[.ddd]e[+-]XX, shortest round-trip
        const char* p = t;
        if (*p == '-') { buf_.push_back('-'); ++p; }
        char digits[24];
        int nd = 0;
        while (p < r.ptr && *p != 'e') { if (*p != '.') digits[nd++] = *p; ++p; }
        int e = 0;
        std::from_chars(p + 1 + (p[1] == '+' ? 1 : 0), r.ptr, e);
        if (e < -4 || e >= 16) {
            buf_.push_back(digits[0]);
            if (nd > 1) { buf_.push_back('.'); buf_.append(digits + 1, nd - 1); }
            buf_.push_back('e');
            buf_.push_back(e < 0 ? '-' : '+');
            const int a = e < 0 ? -e : e;
            if (a < 10) buf_.push_back('0');
            i64(a);
        } else if (e < 0) {
            buf_.append("0.");
            buf_.append((size_t)(-e - 1), '0');
            buf_.append(digits, nd);
        } else {
            const int ip = e + 1;  // digits before the decimal point
            if (nd <= ip) { buf_.append(digits, nd); buf_.append((size_t)(ip - nd), '0'); buf_.append(".0"); }
            else { buf_.append(digits, ip); buf_.push_back('.'); buf_.append(digits + ip, nd - ip); }
        }
    }

  private:
    static constexpr size_t CAP = 1 << 20;
    void newline() { buf_.push_back('\n'); buf_.append((size_t)(indent_ * depth_), ' '); }
    FILE* f_;
    int indent_, depth_ = 0;
    bool ok_ = true;
    std::string buf_;
};

void pair_i64(JsonOut& o, long long a, long long b) {
    bool f;
    o.open('[', f);
    o.item(f); o.i64(a);
    o.item(f); o.i64(b);
    o.close(']', f);
}
void pair_f64(JsonOut& o, double a, double b) {
    bool f;
    o.open('[', f);
    o.item(f); o.f64(a);
    o.item(f); o.f64(b);
    o.close(']', f);
}

void bbox(JsonOut& o, const int64_t* b) {
    bool f;
    o.open('[', f);
    o.item(f); pair_i64(o, b[0], b[1]);
    o.item(f); pair_i64(o, b[2], b[3]);
    o.close(']', f);
}

// neighbour tiles of a border-touching cell (cell_detection.py:877-902), keyed by the [top, right, down, left] flags
int edge_patches(const int8_t* pos, int out[3][2]) {
    const int code = (pos[0] ? 8 : 0) | (pos[1] ? 4 : 0) | (pos[2] ? 2 : 0) | (pos[3] ? 1 : 0);
    static const int T[16][3][2] = {
        /*0000*/ {{9, 9}}, /*0001 left*/ {{0, -1}}, /*0010 down*/ {{1, 0}}, /*0011 down+left*/ {{1, 0}, {1, -1}, {0, -1}},
        /*0100 right*/ {{0, 1}}, /*0101*/ {{9, 9}}, /*0110 right+down*/ {{0, 1}, {1, 1}, {1, 0}}, /*0111*/ {{9, 9}},
        /*1000 top*/ {{-1, 0}}, /*1001 top+left*/ {{0, -1}, {-1, -1}, {-1, 0}}, /*1010*/ {{9, 9}}, /*1011*/ {{9, 9}},
        /*1100 top+right*/ {{-1, 0}, {-1, 1}, {0, 1}}, /*1101*/ {{9, 9}}, /*1110*/ {{9, 9}}, /*1111*/ {{9, 9}}};
    static const int N[16] = {0, 1, 1, 3, 1, 0, 3, 0, 1, 3, 0, 0, 3, 0, 0, 0};
    for (int k = 0; k < N[code]; ++k) { out[k][0] = T[code][k][0]; out[k][1] = T[code][k][1]; }
    return N[code];
}

void contour(JsonOut& o, const cvb_cell_columns& c, long long i, bool close_ring) {
    bool f;
    o.open('[', f);
    const long long a = c.contour_off[i], b = c.contour_off[i + 1];
    for (long long k = a; k < b; ++k) { o.item(f); pair_i64(o, c.contour_pts[2 * k], c.contour_pts[2 * k + 1]); }
    if (close_ring && b > a) { o.item(f); pair_i64(o, c.contour_pts[2 * a], c.contour_pts[2 * a + 1]); }
    o.close(']', f);
}

void cell_record(JsonOut& o, const cvb_cell_columns& c, long long i) {
    bool f;
    o.open('{', f);
    o.key("bbox", f); bbox(o, c.bbox + 4 * i);
    o.key("centroid", f); pair_f64(o, c.centroid[2 * i], c.centroid[2 * i + 1]);
    o.key("contour", f); contour(o, c, i, false);
    o.key("type_prob", f); o.f64(c.type_prob[i]);
    o.key("type", f); o.i64(c.type[i]);
    o.key("patch_coordinates", f); pair_i64(o, c.patch[2 * i], c.patch[2 * i + 1]);
    o.key("cell_status", f); o.i64(c.status[i]);
    o.key("offset_global", f); pair_i64(o, c.offset[2 * i], c.offset[2 * i + 1]);
    o.key("edge_position", f); o.boolean(c.edge[i] != 0);
    if (c.edge[i]) {
        o.key("edge_information", f);
        bool g;
        o.open('{', g);
        o.key("position", g);
        bool h;
        o.open('[', h);
        for (int k = 0; k < 4; ++k) { o.item(h); o.i64(c.position[4 * i + k]); }
        o.close(']', h);
        o.key("edge_patches", g);
        int ep[3][2];
        const int n = edge_patches(c.position + 4 * i, ep);
        if (n == 0) o.null();
        else {
            o.open('[', h);
            for (int k = 0; k < n; ++k) { o.item(h); pair_i64(o, c.patch[2 * i] + ep[k][0], c.patch[2 * i + 1] + ep[k][1]); }
            o.close(']', h);
        }
        o.close('}', g);
    }
    o.close('}', f);
}

void detection_record(JsonOut& o, const cvb_cell_columns& c, long long i) {
    bool f;
    o.open('{', f);
    o.key("bbox", f); bbox(o, c.bbox + 4 * i);
    o.key("centroid", f); pair_f64(o, c.centroid[2 * i], c.centroid[2 * i + 1]);
    o.key("type", f); o.i64(c.type[i]);
    o.close('}', f);
}

}  // namespace

extern "C" __attribute__((visibility("default")))
int cvb_export_json(const char* path, const cvb_cell_columns* cols, const cvb_json_section* sections, int n_sections, int indent) {
    CVB_CHECK(path && cols && (sections || n_sections == 0) && n_sections >= 0, CVB_EARG, "cvb_export_json: null argument");
    CVB_CHECK(cols->n >= 0 && (cols->n == 0 || (cols->bbox && cols->centroid && cols->contour_off && cols->type_prob && cols->type &&
                                                cols->patch && cols->status && cols->offset && cols->edge && cols->position)),
              CVB_EARG, "cvb_export_json: null column");
    for (long long i = 0; i < cols->n; ++i)
        CVB_CHECK(cols->contour_off[i] >= 0 && cols->contour_off[i] <= cols->contour_off[i + 1], CVB_EARG,
                  "cvb_export_json: contour offsets must be non-negative and non-decreasing (cell %lld)", i);
    CVB_CHECK(cols->n == 0 || cols->contour_off[cols->n] == 0 || cols->contour_pts != nullptr, CVB_EARG, "cvb_export_json: null contour points");
    for (int s = 0; s < n_sections; ++s) {
        const cvb_json_section& sec = sections[s];
        CVB_CHECK(sec.kind >= CVB_JSON_CELLS && sec.kind <= CVB_JSON_POINTS && sec.n_idx >= 0 && (sec.idx || sec.n_idx == 0), CVB_EARG,
                  "cvb_export_json: bad section %d", s);
        for (long long k = 0; k < sec.n_idx; ++k)
            CVB_CHECK(sec.idx[k] >= 0 && sec.idx[k] < cols->n, CVB_EARG, "cvb_export_json: cell index %lld out of range in section %d",
                      (long long)sec.idx[k], s);
    }
    FILE* f = fopen(path, "wb");
    CVB_CHECK(f != nullptr, CVB_EARG, "cvb_export_json: cannot open '%s' for writing", path);
    JsonOut o(f, indent);
    for (int s = 0; s < n_sections; ++s) {
        const cvb_json_section& sec = sections[s];
        if (sec.head) o.raw(sec.head);
        o.set_depth(sec.depth);
        bool first;
        o.open('[', first);
        for (long long k = 0; k < sec.n_idx; ++k) {
            const long long i = sec.idx[k];
            o.item(first);
            switch (sec.kind) {
                case CVB_JSON_CELLS: cell_record(o, *cols, i); break;
                case CVB_JSON_DETECTION: detection_record(o, *cols, i); break;
                case CVB_JSON_POLYGONS: {  // one polygon = [closed outer ring]
                    bool g;
                    o.open('[', g);
                    o.item(g);
                    contour(o, *cols, i, true);
                    o.close(']', g);
                    break;
                }
                default: pair_f64(o, cols->centroid[2 * i], cols->centroid[2 * i + 1]); break;
            }
        }
        o.close(']', first);
        if (sec.tail) o.raw(sec.tail);
    }
    o.flush();
    const bool ok = o.ok();
    const bool closed = fclose(f) == 0;
    CVB_CHECK(ok && closed, CVB_EARG, "cvb_export_json: write to '%s' failed", path);
    return CVB_OK;
}
