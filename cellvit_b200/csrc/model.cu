// model.cu -- the CellViT forward (ViT encoder + 3-branch U-Net decoder) sequenced over the tile engine and
// the helper kernels. One call = one batch of tiles; no allocation, no synchronisation, every launch on the
// caller's stream, so the whole forward can be captured into a CUDA graph.
//
// Reference behaviour followed (file:line in the reference tree):
//   models/segmentation/cell_segmentation/cellvit.py:153-244   CellViT.forward / _forward_upsample
//   models/segmentation/cell_segmentation/cellvit.py:586-644   CellViTSAM.forward
//   models/segmentation/cell_segmentation/utils.py:149-233      encoder wrappers (skip extraction, tissue logits)
//   models/encoders/VIT/SAM/image_encoder.py:177-392            SAM block / window attention / rel-pos
//   models/encoders/VIT/vits_histo.py:172-247,404-415           ViT-S block / token preparation
// The shared skip decoders decoder0..3 are evaluated once per batch (the reference re-evaluates them in each of
// the three branch calls, cellvit.py:235-241; identical results in eval mode).
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/cellvit_b200.h"
#include "ops.h"
#include "tc_gemm.h"

struct cvb_model {
    cvb_model_desc d;
    std::unordered_map<std::string, const void*> params;
    // option "attention_tc" (cvb_model_set_option): bit 0 = tcgen05 kernel for the global-attention blocks (flash_tc.cu), bit 1 =
    // tcgen05 kernel for the 14 x 14 windows (window_tc.cu). Both on by default; 0 selects the mma.sync kernels of
    // attention.cu (kept as the independent implementation the parity tests compare against).
    int attn_tc_mode = 3;
    // option "dynamic_tiles": 1 (default) = the persistent kernels (tile engine, window attention) claim their tiles from a
    // per-launch counter (SchedRing, common.cuh) instead of walking a static list; 0 = static lists (ablation)
    int dynamic_tiles = 1;
    // option "square_canvas": tiles whose token grid is not a native grid of the tile engine run their decoder on a zero-extended
    // canvas (see forward_impl). 0 (default) = each dimension extended on its own (a 272 x 400 tile runs on 512 x 512, a 208 x 1024
    // tile on 256 x 1024); 1 = both dimensions extended to the larger canvas edge (ablation / fallback)
    int square_canvas = 0;
    // option "window_pad_skip": 1 (default) = the QKV and attn-out GEMMs of a windowed block run over the real tokens only (QKV:
    // LayerNorm in raster order, rows scattered into window order by the epilogue, the padding rows filled with the bias they would
    // compute to; attn-out: the tcgen05 window attention writes its output un-partitioned); 0 = both GEMMs run over all
    // window-partitioned rows incl. the zero padding (16 % more rows on a 64 x 64 grid; ablation). Same results.
    int window_pad_skip = 1;
};

namespace {

struct Arena {
    uint8_t* base;
    size_t off, cap;
    bool dry;
    size_t peak = 0;  // the branch loop rewinds `off`: the workspace requirement is the high-water mark
    template <class T>
    T* alloc(size_t n) {
        off = align_up(off, 256);
        T* p = reinterpret_cast<T*>(base + off);
        off += n * sizeof(T);
        if (off > peak) peak = off;
        return p;
    }
};

struct Fwd {
    cvb_model& m;
    Arena& A;
    cudaStream_t st;
    int rc = CVB_OK;
    std::string missing;
    // per-launch tile counters of the persistent kernels: one int each, zeroed by ONE memset at the start of the forward
    static constexpr int N_COUNTERS = 1024;
    int* counters = nullptr;
    int n_counters = 0;
    int* next_counter() { return (counters && m.dynamic_tiles && n_counters < N_COUNTERS) ? counters + n_counters++ : nullptr; }
    TcEpilogue with_counter(const TcEpilogue& e) { TcEpilogue c = e; c.sched_counter = next_counter(); return c; }
    // canvas mode (forward_impl): the decoder runs on hc x wc token canvases of which the top-left h x w tokens are real
    bool canvas = false;
    int h = 0, w = 0, hc = 0, wc = 0;
    // re-zero the canvas margin of an NHWC fp16 layer [NB,H,W,C] (H = s * hc): the next 3x3 convolution must see zeros beyond
    // the real image border, and conv (+shift, ReLU) / ConvT (+bias) outputs are not zero there
    void mask(__half* buf, int NB, int H, int W, int C) {
        if (canvas && live()) chk(op_zero_margin(buf, NB, H, W, C * 2, h * (H / hc), w * (W / wc), st));
    }

    template <class T>
    const T* P(const std::string& name) {
        auto it = m.params.find(name);
        if (it == m.params.end() || it->second == nullptr) {
            if (!A.dry && rc == CVB_OK) {
                rc = CVB_EARG;
                cvb_set_error("cvb_forward: parameter '%s' was not registered", name.c_str());
            }
            return nullptr;
        }
        return reinterpret_cast<const T*>(it->second);
    }
    bool live() const { return !A.dry && rc == CVB_OK; }
    void chk(int r) { if (rc == CVB_OK && r != CVB_OK) rc = r; }

    static TcEpilogue epi0() { TcEpilogue e; memset(&e, 0, sizeof(e)); return e; }

    void gemm(const __half* a, int M, int K, const std::string& w, int N, const TcEpilogue& e) {
        const __half* wp = P<__half>(w);
        if (live()) chk(tc_gemm(a, M, K, K, wp, N, K, tc_pick_block_n(N), with_counter(e), st));
    }

    // Conv3x3 + folded BN + ReLU over (src0 || src1), NHWC fp16 -> NHWC fp16 [NB,H,W,N]
    __half* conv_bn_relu(const __half* s0, int C0, const __half* s1, int C1, int NB, int H, int W, const std::string& name, int N) {
        __half* out = A.alloc<__half>((size_t)NB * H * W * N);
        TcEpilogue e = epi0();
        e.kind = TC_EPI_F16; e.act = TC_ACT_RELU; e.out = out; e.ldc = N;
        e.scale = P<float>(name + ".scale"); e.shift = P<float>(name + ".shift");
        const __half* wp = P<__half>(name + ".w");
        if (live()) chk(tc_conv3x3(s0, C0, s1, C1, NB, H, W, wp, N, tc_pick_block_n(N), with_counter(e), st));
        mask(out, NB, H, W, N);
        return out;
    }
    // ConvTranspose2d k2 s2 (+bias): [NB,hin,win,Cin] -> [NB,2hin,2win,Cout]
    __half* conv_t(const __half* src, int Cin, int NB, int hin, int win, const std::string& name, int Cout) {
        __half* out = A.alloc<__half>((size_t)NB * 4 * hin * win * Cout);
        TcEpilogue e = epi0();
        e.kind = TC_EPI_CONVT; e.out = out; e.ldc = Cout; e.shift = P<float>(name + ".b");
        e.ct_cout = Cout; e.ct_hin = hin; e.ct_win = win;
        const __half* wp = P<__half>(name + ".w");
        if (live()) chk(tc_gemm(src, NB * hin * win, Cin, Cin, wp, 4 * Cout, Cin, tc_pick_block_n(4 * Cout), with_counter(e), st));
        mask(out, NB, 2 * hin, 2 * win, Cout);
        return out;
    }
    // Deconv2DBlock (utils.py:46-86): ConvT -> Conv3x3 -> BN -> ReLU
    __half* deconv_block(const __half* src, int Cin, int NB, int hin, int win, const std::string& name, int Cout) {
        __half* t = conv_t(src, Cin, NB, hin, win, name + ".ct", Cout);
        return conv_bn_relu(t, Cout, nullptr, 0, NB, 2 * hin, 2 * win, name + ".conv", Cout);
    }
};

int pad64(int c) { return (c + 63) / 64 * 64; }

// Native token grids of the tile engine's decoder tilings (128-pixel blocks at every level): 16, 32 or 64 tokens per edge.
int canvas_tokens(int t) { return t <= 16 ? 16 : (t <= 32 ? 32 : 64); }

int forward_impl(cvb_model& m, const float* x, int B, int H, int W, float* o_np, float* o_hv, float* o_nt, float* o_tissue,
                 float* o_tokens, uint8_t* o_np_arg, uint8_t* o_nt_arg, Arena& A, cudaStream_t st) {
    const cvb_model_desc& d = m.d;
    Fwd f{m, A, st};
    const int h = H / 16, w = W / 16, T = h * w, D = d.embed_dim, heads = d.num_heads, hd = D / heads;
    const bool sam = d.sam != 0;
    const int Tx = sam ? T : T + 1;  // ViT-S carries a cls token (vits_histo.py:404-415)
    const int skip = sam ? 0 : 1;
    const float scale = 1.0f / sqrtf((float)hd);
    const bool live = !A.dry;
    // Any H, W divisible by 16 (cellvit.py:170-175, 603-608). The encoder runs on the real h x w token grid. The decoder's
    // tilings want 16 / 32 / 64 tokens per edge: other grids run the decoder on the next larger canvas, the real tokens /
    // pixels top-left and zeros elsewhere, with the margin re-zeroed after every layer -- inside the real region every
    // convolution then sees exactly the zero padding of the real border, so the cropped result is the reference's.
    int hc = canvas_tokens(h), wc = canvas_tokens(w);
    if (m.square_canvas) hc = wc = (hc > wc ? hc : wc);
    const bool canvas = hc != h || wc != w;
    const int Hc = 16 * hc, Wc = 16 * wc, Tc = hc * wc;
    f.canvas = canvas; f.h = h; f.w = w; f.hc = hc; f.wc = wc;
    f.counters = A.alloc<int>(Fwd::N_COUNTERS);
    if (live) CVB_CUDA(cudaMemsetAsync(f.counters, 0, Fwd::N_COUNTERS * sizeof(int), st));

    // ------------------------------------------------------------------ patch embedding (+bias +pos)
    __half* a0 = A.alloc<__half>((size_t)B * T * 768);
    float* xs = A.alloc<float>((size_t)B * Tx * D);
    if (live) f.chk(op_patch_im2col(x, B, H, W, 16, a0, st));
    {
        TcEpilogue e = Fwd::epi0();
        e.kind = TC_EPI_RES_F32; e.out = xs; e.ldc = D; e.shift = f.P<float>("patch.b");
        e.res = f.P<float>("pos"); e.ldres = D; e.res_mod = T; e.res_off = skip;
        if (!sam) { e.row_map = TC_ROW_SEQ; e.row_seq = T; e.row_pad = 1; e.row_off = 1; }
        f.gemm(a0, B * T, 768, "patch.w", D, e);
        if (!sam) {
            const float* cp = f.P<float>("clspos");
            if (f.live())
                for (int b = 0; b < B; ++b)
                    CVB_CUDA(cudaMemcpyAsync(xs + (size_t)b * Tx * D, cp, (size_t)D * 4, cudaMemcpyDeviceToDevice, st));
        }
    }

    // ------------------------------------------------------------------ transformer blocks
    const int ws = sam ? d.window_size : 0;
    const int g = ws > 0 ? (h + ws - 1) / ws : 0;
    const int Tw = g * g * ws * ws;
    const size_t rows_max = (size_t)B * (Tw > Tx ? Tw : Tx);
    __half* ln = A.alloc<__half>(rows_max * D);
    __half* qkv = A.alloc<__half>(rows_max * 3 * D);
    __half* att = A.alloc<__half>(rows_max * D);
    __half* hid = A.alloc<__half>((size_t)B * Tx * 4 * D);
    __half* z[4];
    for (int k = 0; k < 4; ++k) z[k] = A.alloc<__half>((size_t)B * Tc * D);
    __half* zt = canvas ? A.alloc<__half>((size_t)B * T * D) : nullptr;  // compact skip tokens before they are embedded

    __half* ybuf = A.alloc<__half>(rows_max * D);
    // global attention (SAM's four global blocks, every block of ViT-S) runs on the tcgen05 attention kernel
    const bool attn_tc = (m.attn_tc_mode & 1) && op_attention_tc_supported(Tx, hd, sam ? reinterpret_cast<const __half*>(1) : nullptr, h, w);
    const bool win_tc = sam && (m.attn_tc_mode & 2) && ws > 0 && op_window_attention_tc_supported(ws * ws, hd, ws, ws);
    const size_t attn_ws_bytes = (attn_tc && sam) ? op_attention_tc_workspace_bytes(B, Tx, heads) : 0;   // rel-pos bias tables
    uint8_t* attn_ws = nullptr;
    if (attn_tc && sam) {
        attn_ws = A.alloc<uint8_t>(attn_ws_bytes + 1024);
        attn_ws = reinterpret_cast<uint8_t*>(((uintptr_t)attn_ws + 1023) & ~(uintptr_t)1023);
    }
    for (int i = 0; i < d.depth; ++i) {
        const std::string p = "b" + std::to_string(i);
        bool is_global = !sam;
        for (int k = 0; k < d.n_global; ++k) is_global |= (d.global_idx[k] == i);
        const bool win = sam && !is_global && ws > 0;
        const int rows = win ? B * Tw : B * Tx;
        const float* n1w = f.P<float>(p + ".n1.w");
        const float* n1b = f.P<float>(p + ".n1.b");
        // Windowed blocks pad the token grid to a multiple of 14 AFTER norm1 (image_encoder.py:180-184): the padding rows of the
        // QKV input are zeros, so their QKV rows equal the bias. The GEMM therefore runs over the B*T real tokens in raster order
        // and its epilogue scatters them to their window-partitioned rows; the padding rows are filled with the bias.
        const bool pad_skip = win && m.window_pad_skip && Tw != T;
        if (f.live()) f.chk(op_layernorm_f16(xs, n1w, n1b, 1e-6f, pad_skip ? B * Tx : rows, D, ln, (win && !pad_skip) ? 1 : 0, B, h, w, ws, g, st));
        {
            TcEpilogue e = Fwd::epi0();
            e.kind = TC_EPI_F16; e.out = qkv; e.ldc = 3 * D; e.shift = f.P<float>(p + ".qkv.b");
            if (pad_skip) { e.row_map = TC_ROW_TO_WINDOW; e.win_size = ws; e.win_grid = g; e.tok_h = h; e.tok_w = w; }
            f.gemm(ln, pad_skip ? B * Tx : rows, D, p + ".qkv.w", 3 * D, e);
            if (pad_skip && f.live()) f.chk(op_window_pad_fill(qkv, e.shift, B, 3 * D, h, w, ws, g, st));
        }
        // ... and the tcgen05 window attention writes its output un-partitioned (raster order, padding rows dropped), so that the
        // attn-out projection runs over the real tokens only as well
        const bool att_raster = pad_skip && win_tc;
        const int Gb = win ? B * g * g : B, S = win ? ws * ws : Tx, gh = win ? ws : h, gw = win ? ws : w;
        const __half* th = sam ? f.P<__half>(p + ".relh") : nullptr;
        const __half* tw = sam ? f.P<__half>(p + ".relw") : nullptr;
        if (f.live()) {
            if (attn_tc && !win) f.chk(op_attention_tc(qkv, Gb, S, heads, hd, scale, th, tw, gh, gw, att, attn_ws, attn_ws_bytes, st));
            else if (win_tc && win) f.chk(op_window_attention_tc(qkv, Gb, heads, hd, scale, f.P<__half>(p + ".relcat"), att, f.next_counter(), st,
                                                                 att_raster ? g : 0, h, w));
            else f.chk(op_attention(qkv, Gb, S, heads, hd, scale, th, tw, gh, gw, att, st));
        }
        {
            // attn-out projection writes fp16 y (+bias) in the row order of `att` (raster, or window order on the mma.sync /
            // no-pad-skip paths); the residual add x += y is fused into norm2 below (coalesced) -- with K = 1280 a scattered fp32
            // read-modify-write epilogue is slower than the main loop (131 -> 69 us), unlike lin2 (K = 5120) which keeps it.
            TcEpilogue e = Fwd::epi0();
            e.kind = TC_EPI_F16; e.out = ybuf; e.ldc = D; e.shift = f.P<float>(p + ".proj.b");
            f.gemm(att, att_raster ? B * Tx : rows, D, p + ".proj.w", D, e);
        }
        const float* n2w = f.P<float>(p + ".n2.w");
        const float* n2b = f.P<float>(p + ".n2.b");
        {
            LnFuse lf{ybuf, (win && !att_raster) ? 1 : 0, xs, nullptr, 0, Tx, 1};
            if (f.live()) f.chk(op_residual_ln(xs, lf, n2w, n2b, 1e-6f, B * Tx, D, ln, 0, B, h, w, ws, g, st));
        }
        {
            TcEpilogue e = Fwd::epi0();
            e.kind = TC_EPI_F16; e.act = TC_ACT_GELU; e.out = hid; e.ldc = 4 * D; e.shift = f.P<float>(p + ".fc1.b");
            f.gemm(ln, B * Tx, D, p + ".fc1.w", 4 * D, e);
        }
        {
            TcEpilogue e = Fwd::epi0();
            e.kind = TC_EPI_RES_F32; e.out = xs; e.ldc = D; e.res = xs; e.ldres = D; e.shift = f.P<float>(p + ".fc2.b");
            f.gemm(hid, B * Tx, 4 * D, p + ".fc2.w", D, e);
        }
        for (int k = 0; k < 4; ++k)
            if (d.extract[k] == i + 1) {
                if (f.live()) f.chk(op_cast_rows_f16(xs, B, Tx, skip, D, canvas ? zt : z[k], st));
                if (canvas && f.live()) f.chk(op_copy_planes(zt, h, w * (D / 8), z[k], hc, wc * (D / 8), B, 16, st));
                if (k == 3 && o_tokens && f.live()) f.chk(op_tokens_nchw(xs, B, Tx, skip, D, o_tokens, st));
            }
    }

    // ------------------------------------------------------------------ tissue classifier
    if (sam) {
        // neck: conv1x1 -> LayerNorm2d -> conv3x3 -> LayerNorm2d -> spatial mean -> Linear (image_encoder.py:97-113)
        __half* xf = A.alloc<__half>((size_t)B * T * D);
        float* n0 = A.alloc<float>((size_t)B * T * 256);
        __half* n1 = A.alloc<__half>((size_t)B * T * 256);
        float* n2 = A.alloc<float>((size_t)B * T * 256);
        if (f.live()) f.chk(op_cast_rows_f16(xs, B, Tx, 0, D, xf, st));
        TcEpilogue e = Fwd::epi0();
        e.kind = TC_EPI_RES_F32; e.out = n0; e.ldc = 256;
        f.gemm(xf, B * T, D, "neck.0.w", 256, e);
        const float* g1 = f.P<float>("neck.1.w");
        const float* b1 = f.P<float>("neck.1.b");
        if (f.live()) f.chk(op_layernorm_f16(n0, g1, b1, 1e-6f, B * T, 256, n1, 0, B, h, w, 0, 0, st));
        const __half* w2 = f.P<__half>("neck.2.w");
        if (!canvas) {
            e.out = n2;
            if (f.live()) f.chk(tc_conv3x3(n1, 256, nullptr, 0, B, h, w, w2, 256, 256, f.with_counter(e), st));
        } else {
            __half* n1c = A.alloc<__half>((size_t)B * Tc * 256);
            float* n2c = A.alloc<float>((size_t)B * Tc * 256);
            e.out = n2c;
            if (f.live()) {
                f.chk(op_copy_planes(n1, h, w * 32, n1c, hc, wc * 32, B, 16, st));
                f.chk(tc_conv3x3(n1c, 256, nullptr, 0, B, hc, wc, w2, 256, 256, f.with_counter(e), st));
                f.chk(op_copy_planes(n2c, hc, wc * 64, n2, h, w * 64, B, 16, st));
            }
        }
        const float* g3 = f.P<float>("neck.3.w");
        const float* b3 = f.P<float>("neck.3.b");
        const float* cw = f.P<float>("cls.w");
        const float* cb = f.P<float>("cls.b");
        float* lnm = A.alloc<float>(op_ln_mean_linear_scratch_floats(B, T, 256));
        if (f.live() && o_tissue) f.chk(op_ln_mean_linear(n2, B, T, 256, g3, b3, 1e-6f, cw, cb, d.n_tissue, o_tissue, lnm, st));
    } else {
        const float* gw_ = f.P<float>("norm.w");
        const float* gb_ = f.P<float>("norm.b");
        const float* hw = f.P<float>("head.w");
        const float* hb = f.P<float>("head.b");
        if (f.live() && o_tissue) f.chk(op_cls_head(xs, B, Tx, D, gw_, gb_, 1e-6f, hw, hb, d.n_tissue, o_tissue, st));
    }

    // ------------------------------------------------------------------ shared skip decoders (cellvit.py:116-131)
    const int s11 = d.skip11, s12 = d.skip12, bt = d.bott_pad;
    // from here on every layer lives on the canvas (hc x wc tokens = Hc x Wc pixels; == h, w, H, W for the native grids)
    __half* st0 = A.alloc<__half>((size_t)B * Hc * Wc * 64);
    {
        const float* sw = f.P<float>("decoder0.0.w");
        const float* sc = f.P<float>("decoder0.0.scale");
        const float* sh = f.P<float>("decoder0.0.shift");
        // the stem stencil pads at the real border itself: run it on the real tile, then embed its output
        __half* st0r = canvas ? A.alloc<__half>((size_t)B * H * W * 64) : st0;
        if (f.live()) f.chk(op_stem_conv(x, B, H, W, sw, sc, sh, st0r, 64, st));
        if (canvas && f.live()) f.chk(op_copy_planes(st0r, H, W * 8, st0, Hc, Wc * 8, B, 16, st));
    }
    __half* s0 = f.conv_bn_relu(st0, 64, nullptr, 0, B, Hc, Wc, "decoder0.1", 64);
    __half* s1 = f.deconv_block(z[0], D, B, hc, wc, "decoder1.0", s11);
    s1 = f.deconv_block(s1, s11, B, 2 * hc, 2 * wc, "decoder1.1", s12);
    s1 = f.deconv_block(s1, s12, B, 4 * hc, 4 * wc, "decoder1.2", 128);
    __half* s2 = f.deconv_block(z[1], D, B, hc, wc, "decoder2.0", s11);
    s2 = f.deconv_block(s2, s11, B, 2 * hc, 2 * wc, "decoder2.1", 256);
    __half* s3 = f.deconv_block(z[2], D, B, hc, wc, "decoder3.0", bt);

    // ------------------------------------------------------------------ three upsampling branches (cellvit.py:212-244)
    // arg: optional u8 arg-max plane of the branch (K12 fusion), over the first arg_nc classes (NP: the binary map's two, also when
    // the regression channels are present)
    struct Br { const char* name; float* out; int nc; uint8_t* arg; int arg_nc; };
    const Br branches[3] = {{"np", o_np, d.n_np_out, o_np_arg, 2}, {"hv", o_hv, 2, nullptr, 0}, {"nt", o_nt, d.n_nt, o_nt_arg, d.n_nt}};
    // CellViT: one upsampling trunk per output. The *Shared variants (cellvit_shared.py:147-231): ONE trunk ("dec") whose 64-channel
    // feature map feeds three 1x1 heads -- here the last trunk convolution (decoder0_header.1) runs once per head with that head
    // fused into its epilogue, so the feature map is never written (the same launches as a CellViT branch tail).
    const bool shared = d.shared_decoder != 0;
    const int n_trunks = shared ? 1 : 3;
    const size_t mark = A.off;
    for (int t = 0; t < n_trunks; ++t) {
        A.off = mark;  // trunk activations reuse the same arena region (stream order serialises the trunks)
        const std::string n = shared ? "dec" : branches[t].name;
        // a trunk none of whose outputs is asked for is not launched; its allocations still happen (the dry pass that sizes the
        // workspace and the live pass must agree), which is what a temporarily "dry" arena does
        bool wanted = false;
        for (int k = 0; k < 3; ++k) wanted |= (shared || k == t) && branches[k].out != nullptr;
        const bool saved_dry = A.dry;
        if (!wanted) A.dry = true;
        __half* b = f.conv_t(z[3], D, B, hc, wc, n + ".bottleneck", bt);
        b = f.conv_bn_relu(s3, bt, b, bt, B, 2 * hc, 2 * wc, n + ".d3.0", bt);
        b = f.conv_bn_relu(b, bt, nullptr, 0, B, 2 * hc, 2 * wc, n + ".d3.1", bt);
        b = f.conv_bn_relu(b, bt, nullptr, 0, B, 2 * hc, 2 * wc, n + ".d3.2", bt);
        b = f.conv_t(b, bt, B, 2 * hc, 2 * wc, n + ".d3.ct", 256);
        b = f.conv_bn_relu(s2, 256, b, 256, B, 4 * hc, 4 * wc, n + ".d2.0", 256);
        b = f.conv_bn_relu(b, 256, nullptr, 0, B, 4 * hc, 4 * wc, n + ".d2.1", 256);
        b = f.conv_t(b, 256, B, 4 * hc, 4 * wc, n + ".d2.ct", 128);
        b = f.conv_bn_relu(s1, 128, b, 128, B, 8 * hc, 8 * wc, n + ".d1.0", 128);
        b = f.conv_bn_relu(b, 128, nullptr, 0, B, 8 * hc, 8 * wc, n + ".d1.1", 128);
        b = f.conv_t(b, 128, B, 8 * hc, 8 * wc, n + ".d1.ct", 64);
        b = f.conv_bn_relu(s0, 64, b, 64, B, Hc, Wc, n + ".d0.0", 64);
        A.dry = saved_dry;
        for (int k = 0; k < 3; ++k) {
            if (!shared && k != t) continue;
            const Br& br = branches[k];
            // canvas mode: the fused head writes canvas-sized planes, which are cropped to the caller's [.., H, W] outputs
            float* head_out = br.out;
            uint8_t* head_arg = br.arg;
            if (canvas) {
                head_out = A.alloc<float>((size_t)B * br.nc * Hc * Wc);
                head_arg = A.alloc<uint8_t>((size_t)B * Hc * Wc);
                if (!br.arg) head_arg = nullptr;
            }
            TcEpilogue e = Fwd::epi0();
            e.kind = TC_EPI_HEAD;
            e.scale = f.P<float>(n + ".d0.1.scale"); e.shift = f.P<float>(n + ".d0.1.shift");
            e.head_w = f.P<float>(std::string(br.name) + ".head.w"); e.head_b = f.P<float>(std::string(br.name) + ".head.b");
            e.head_nc = br.nc; e.head_hw = Hc * Wc; e.head_out = head_out; e.head_argmax = head_arg; e.head_argmax_nc = br.arg_nc;
            const __half* wp = f.P<__half>(n + ".d0.1.w");
            if (f.live() && br.out) {
                f.chk(tc_conv3x3(b, 64, nullptr, 0, B, Hc, Wc, wp, 64, 64, f.with_counter(e), st));
                if (canvas) {
                    f.chk(op_copy_planes(head_out, Hc, Wc, br.out, H, W, (long long)B * br.nc, 4, st));
                    if (br.arg) f.chk(op_copy_planes(head_arg, Hc, Wc, br.arg, H, W, B, 1, st));
                }
            }
        }
    }
    (void)pad64;
    return f.rc;
}

}  // namespace

#define CVB_API extern "C" __attribute__((visibility("default")))

CVB_API int cvb_model_create(const cvb_model_desc* desc, cvb_model** out) {
    CVB_CHECK(desc && out, CVB_EARG, "cvb_model_create: null argument");
    CVB_CHECK(desc->embed_dim > 0 && desc->num_heads > 0 && desc->embed_dim % desc->num_heads == 0, CVB_EARG,
              "cvb_model_create: bad embed_dim/num_heads");
    const int hd = desc->embed_dim / desc->num_heads;
    CVB_CHECK(hd == 64 || hd == 80, CVB_ESHAPE, "cvb_model_create: head dim %d not supported (64, 80)", hd);
    CVB_CHECK(desc->embed_dim % 64 == 0 && desc->skip11 % 64 == 0 && desc->skip12 % 64 == 0 && desc->bott_pad % 64 == 0,
              CVB_ESHAPE, "cvb_model_create: channel widths must be padded to multiples of 64");
    CVB_CHECK(desc->n_np_out >= 1 && desc->n_np_out <= 8 && desc->n_nt >= 1 && desc->n_nt <= 8, CVB_ESHAPE,
              "cvb_model_create: head widths must be in 1..8");
    CVB_CHECK(desc->n_global >= 0 && desc->n_global <= 8, CVB_EARG, "cvb_model_create: n_global out of range");
    cvb_model* m = new cvb_model();
    m->d = *desc;
    *out = m;
    return CVB_OK;
}

CVB_API int cvb_model_set_param(cvb_model* m, const char* name, const void* dev_ptr) {
    CVB_CHECK(m && name, CVB_EARG, "cvb_model_set_param: null argument");
    m->params[name] = dev_ptr;
    return CVB_OK;
}

static int check_shape(const cvb_model* m, int B, int H, int W) {
    CVB_CHECK(m != nullptr && B > 0, CVB_EARG, "cvb_forward: null model or empty batch");
    CVB_CHECK(H % 16 == 0 && W % 16 == 0 && H > 0 && W > 0, CVB_ESHAPE, "Input images must be divisible by the patch size (%dx%d)", H, W);
    CVB_CHECK(H <= 1024 && W <= 1024, CVB_ESHAPE, "cvb_forward: tile %dx%d exceeds the 1024-pixel edge the engine is sized for", H, W);
    // SAM adds pos_embed[:, :h, :h, :] to the [B,h,w,D] tokens (utils.py:222-224): that only broadcasts for square token grids
    CVB_CHECK(!m->d.sam || H == W, CVB_ESHAPE, "cvb_forward: the SAM encoders take square tiles only (%dx%d)", H, W);
    return CVB_OK;
}

CVB_API int cvb_model_workspace_bytes(cvb_model* m, int B, int H, int W, size_t* out) {
    CVB_CHECK(out != nullptr, CVB_EARG, "cvb_model_workspace_bytes: null out");
    CVB_TRY(check_shape(m, B, H, W));
    Arena A{nullptr, 0, 0, true};
    forward_impl(*m, nullptr, B, H, W, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, A, nullptr);
    const size_t peak = A.peak;
    *out = peak + 4096;
    return CVB_OK;
}

static int forward_checked(cvb_model* m, const float* x, int B, int H, int W, float* np_logits, float* hv, float* nt_logits,
                           float* tissue, float* tokens, uint8_t* np_argmax, uint8_t* nt_argmax, void* workspace, size_t ws_bytes, void* stream) {
    CVB_TRY(check_shape(m, B, H, W));
    CVB_CHECK(x && workspace, CVB_EARG, "cvb_forward: null input or workspace");
    size_t need = 0;
    CVB_TRY(cvb_model_workspace_bytes(m, B, H, W, &need));
    CVB_CHECK(ws_bytes >= need, CVB_EWORKSPACE, "cvb_forward: workspace %zu < required %zu bytes", ws_bytes, need);
    CVB_CHECK(((uintptr_t)workspace & 255) == 0, CVB_EARG, "cvb_forward: workspace must be 256-byte aligned");
    Arena A{reinterpret_cast<uint8_t*>(workspace), 0, ws_bytes, false};
    return forward_impl(*m, x, B, H, W, np_logits, hv, nt_logits, tissue, tokens, np_argmax, nt_argmax, A, (cudaStream_t)stream);
}

CVB_API int cvb_forward(cvb_model* m, const float* x, int B, int H, int W, float* np_logits, float* hv, float* nt_logits,
                        float* tissue, float* tokens, void* workspace, size_t ws_bytes, void* stream) {
    return forward_checked(m, x, B, H, W, np_logits, hv, nt_logits, tissue, tokens, nullptr, nullptr, workspace, ws_bytes, stream);
}

CVB_API int cvb_forward_argmax(cvb_model* m, const float* x, int B, int H, int W, float* np_logits, float* hv, float* nt_logits,
                               float* tissue, float* tokens, uint8_t* np_argmax, uint8_t* nt_argmax, void* workspace, size_t ws_bytes,
                               void* stream) {
    CVB_CHECK(np_logits && nt_logits, CVB_EARG, "cvb_forward_argmax: the arg-max planes come out of the NP / NT heads, which need their logit outputs");
    return forward_checked(m, x, B, H, W, np_logits, hv, nt_logits, tissue, tokens, np_argmax, nt_argmax, workspace, ws_bytes, stream);
}

CVB_API void cvb_model_destroy(cvb_model* m) { delete m; }

CVB_API int cvb_model_set_option(cvb_model* m, const char* name, int value) {
    CVB_CHECK(m && name, CVB_EARG, "cvb_model_set_option: null argument");
    if (std::string(name) == "attention_tc") {
        CVB_CHECK(value >= 0 && value <= 3, CVB_EARG, "cvb_model_set_option: attention_tc must be in 0..3");
        m->attn_tc_mode = value;
        return CVB_OK;
    }
    if (std::string(name) == "dynamic_tiles") {
        m->dynamic_tiles = value != 0;
        return CVB_OK;
    }
    if (std::string(name) == "square_canvas") {
        m->square_canvas = value != 0;
        return CVB_OK;
    }
    if (std::string(name) == "window_pad_skip") {
        m->window_pad_skip = value != 0;
        return CVB_OK;
    }
    cvb_set_error("cvb_model_set_option: unknown option '%s'", name);
    return CVB_EARG;
}
