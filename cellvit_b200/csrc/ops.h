// ops.h -- host launchers of the non-GEMM kernels (layer norm, attention, layout transforms, stencils).
// Every launcher is asynchronous on `stream` and returns a CVB_* status.
#pragma once
#include "common.cuh"

// --- ops_misc.cu ------------------------------------------------------------------------------------
// Non-overlapping PxP patches of x [B,3,H,W] fp32 -> A [B*(H/P)*(W/P), 3*P*P] fp16, column = c*P*P + ky*P + kx
// (the flattening of the reference Conv2d weight, image_encoder.py:418-426 / vits_histo.py:273-280).
int op_patch_im2col(const float* x, int B, int H, int W, int P, __half* out, cudaStream_t stream);

// LayerNorm over the last dim (biased variance, eps inside the sqrt) fp32 -> fp16.
//   map 0: dst row r <- src row r;
//   map 1: window partition (image_encoder.py:263-288): dst rows are [B, g*g windows, ws*ws tokens]; tokens that
//          fall outside the tok_h x tok_w grid are written as zeros (padding happens AFTER norm1).
int op_layernorm_f16(const float* x, const float* gamma, const float* beta, float eps, int rows_dst, int D, __half* out,
                     int map, int B, int tok_h, int tok_w, int ws, int g, cudaStream_t stream);

// Residual-fused LayerNorm for the encoder blocks. Per destination row (mapping as in op_layernorm_f16):
//   v = x[src] (+ add[add_row])            add fp16: the previous GEMM's output (attn-out proj or MLP lin2 + bias)
//   x_out[src] = v      (fp32, if set)     -- the residual update  x = x + f(x)  happens here, fully coalesced,
//   cast_out[..] = v    (fp16, if set)        instead of as a scattered fp32 read-modify-write in the GEMM epilogue
//   out[dst] = LayerNorm(v)  (if norm != 0)
// add_map 0: add row = src token row; 1: add rows are window-partitioned [B, g*g, ws*ws] (image_encoder.py:291-318).
// cast_out drops the first `cast_skip` rows of every batch item of `tok_per_item` rows (ViT-S cls token).
struct LnFuse {
    const __half* add;
    int add_map;
    float* x_out;
    __half* cast_out;
    int cast_skip, tok_per_item;
    int norm;
};
int op_residual_ln(const float* x, const LnFuse& f, const float* gamma, const float* beta, float eps, int rows_dst, int D,
                   __half* out, int map, int B, int tok_h, int tok_w, int ws, int g, cudaStream_t stream);

// fp32 [B, T_src, D] rows (skipping `skip` leading rows per batch item) -> fp16 [B, T_src-skip, D].
int op_cast_rows_f16(const float* x, int B, int T_src, int skip, int D, __half* out, cudaStream_t stream);

// fp32 token rows [B, T_src(+skip), D] -> fp32 NCHW [B, D, T] (tokens output of the reference forward).
int op_tokens_nchw(const float* x, int B, int T_src, int skip, int D, float* out, cudaStream_t stream);

// decoder0.0: Conv3x3(3->32)+BN+ReLU on x [B,3,H,W] fp32 -> NHWC fp16 with `cpad` channels (>= 32, rest zero).
// w [32,3,3,3] fp32, scale/shift [32] (folded BN + bias).
int op_stem_conv(const float* x, int B, int H, int W, const float* w, const float* scale, const float* shift,
                 __half* out, int cpad, cudaStream_t stream);

// Padding rows of a window-partitioned fp16 matrix [B*g*g*ws*ws, N] <- bias (fp32 [N], rounded to fp16): the value a Linear layer
// has on the all-zero rows the reference pads with after norm1 (image_encoder.py:180-184), so that the QKV GEMM of a windowed
// block only runs over the real tokens (TC_ROW_TO_WINDOW scatters its rows into window order). Real rows are not touched.
int op_window_pad_fill(__half* out, const float* bias, int B, int N, int tok_h, int tok_w, int ws, int g, cudaStream_t stream);

// Canvas helpers (model.cu: tiles whose token grid is not a native grid of the tile engine run the decoder on a
// zero-extended canvas). copy_planes: dst[p][y][x] = (y < sH && x < sW) ? src[p][y][x] : 0 over `planes` planes of
// elements of 1, 4 or 16 bytes (crop when dst is smaller, zero-extending embed when larger).
// zero_margin: zero every pixel with y >= vH or x >= vW of NHWC planes [planes][H][W][bytes_per_px].
int op_copy_planes(const void* src, int sH, int sW, void* dst, int dH, int dW, long long planes, int elem_bytes, cudaStream_t stream);
int op_zero_margin(void* buf, long long planes, int H, int W, int bytes_per_px, int vH, int vW, cudaStream_t stream);

// SAM neck tail (image_encoder.py:110-113, utils.py:230-233, cellvit.py:613): per image LayerNorm2d over C of
// y [B, T, C] fp32, mean over T, then Linear(C -> n_out). out [B, n_out] fp32.
int op_ln_mean_linear(const float* y, int B, int T, int C, const float* gamma, const float* beta, float eps,
                      const float* w, const float* b, int n_out, float* out, float* scratch, cudaStream_t stream);
size_t op_ln_mean_linear_scratch_floats(int B, int T, int C);  // size of `scratch` (per-chunk partial sums)

// ViT-256 tissue head (utils.py:171-172): LayerNorm(x[:,0]) then Linear. x [B, T, D] fp32.
int op_cls_head(const float* x, int B, int T, int D, const float* gamma, const float* beta, float eps,
                const float* w, const float* b, int n_out, float* out, cudaStream_t stream);

// --- attention.cu -----------------------------------------------------------------------------------
// softmax(scale * q k^T + bias) v per (group, head); groups are windows or whole images.
// qkv fp16 [Gb*S, 3*D] (q | k | v, head-major inside each); out fp16 [Gb*S, D], head h at columns [h*hd, (h+1)*hd).
// Decomposed relative position (image_encoder.py:354-392): Rh / Rw fp16 tables [2*gh-1, hd] / [2*gw-1, hd]
// (row = q - k + g - 1), applied to the UNSCALED q inside the kernel; both null = no bias (ViT-S).
int op_attention(const __half* qkv, int Gb, int S, int heads, int hd, float scale, const __half* Rh, const __half* Rw,
                 int gh, int gw, __half* out, cudaStream_t stream);

// Decomposed rel-pos bias tables of a global-attention call (attention.cu, MMA path): bias_h / bias_w fp16
// [(g*heads + head)*S + q][64] = log2(e) * <q, R[qpos - k + g - 1]> for k < gh / gw (other entries zero).
int op_relpos_tables(const __half* qkv, int Gb, int S, int heads, int hd, const __half* Rh, const __half* Rw, int gh, int gw,
                     __half* bias_h, __half* bias_w, cudaStream_t stream);

// --- flash_tc.cu ------------------------------------------------------------------------------------
// tcgen05 version of op_attention for the SAM global-attention shape (head dim 80, 64-wide token grid, S % 128 == 0,
// rel-pos tables present). `workspace` (1024-byte aligned, op_attention_tc_workspace_bytes) holds V^T and the two
// decomposed rel-pos bias tables of the call.
bool op_attention_tc_supported(int S, int hd, const __half* Rh, int gh, int gw);
size_t op_attention_tc_workspace_bytes(int Gb, int S, int heads);
int op_attention_tc(const __half* qkv, int Gb, int S, int heads, int hd, float scale, const __half* Rh, const __half* Rw, int gh,
                    int gw, __half* out, void* workspace, size_t ws_bytes, cudaStream_t stream);

// --- window_tc.cu -----------------------------------------------------------------------------------
// tcgen05 attention for the 14 x 14 SAM windows (196 tokens, head dim 80), one kernel without pre-passes or workspace.
// qkv fp16 [n_items*196, 3*D] in window order, relcat fp16 [64, 80] = rel_h table rows at 0.., rel_w table rows at 32..
// (packing.py), out fp16 [n_items*196, D].
bool op_window_attention_tc_supported(int S, int hd, int gh, int gw);
// sched_counter: optional device int, zero at launch: the (window, head) items are then claimed dynamically (SchedRing).
// un_g > 0: the output is written UN-PARTITIONED (window_unpartition, image_encoder.py:291-318): out fp16 [B*un_h*un_w, D] in raster
// token order for un_g x un_g windows per image over an un_h x un_w token grid, rows of padding tokens dropped.
int op_window_attention_tc(const __half* qkv, int n_items, int heads, int hd, float scale, const __half* relcat, __half* out,
                           int* sched_counter, cudaStream_t stream, int un_g = 0, int un_h = 0, int un_w = 0);
