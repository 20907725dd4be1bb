// capi_ops.cu -- operator-level C-ABI entry points (declared in include/cellvit_b200.h).
// Thin wrappers: plain pointers and sizes in, status code out, asynchronous on the given stream.
#include "ops.h"
#include "tc_gemm.h"

#define CVB_API extern "C" __attribute__((visibility("default")))

CVB_API int cvb_op_gemm_f16(const void* A, int M, int K, long long lda, const void* W, int N, long long ldw,
                            int block_n, const TcEpilogue* epi, void* stream) {
    CVB_CHECK(epi != nullptr, CVB_EARG, "cvb_op_gemm_f16: null epilogue");
    if (block_n <= 0) block_n = tc_pick_block_n(N);
    return tc_gemm((const __half*)A, M, K, lda, (const __half*)W, N, ldw, block_n, *epi, (cudaStream_t)stream);
}

CVB_API int cvb_op_conv3x3_f16(const void* src0, int C0, const void* src1, int C1, int NB, int H, int W,
                               const void* Wp, int N, int block_n, const TcEpilogue* epi, void* stream) {
    CVB_CHECK(epi != nullptr, CVB_EARG, "cvb_op_conv3x3_f16: null epilogue");
    if (block_n <= 0) block_n = tc_pick_block_n(N);
    return tc_conv3x3((const __half*)src0, C0, (const __half*)src1, C1, NB, H, W, (const __half*)Wp, N, block_n, *epi,
                      (cudaStream_t)stream);
}

CVB_API int cvb_op_layernorm_f16(const float* x, const float* gamma, const float* beta, float eps, int rows_dst, int D,
                                 void* out, int map, int B, int tok_h, int tok_w, int ws, int g, void* stream) {
    return op_layernorm_f16(x, gamma, beta, eps, rows_dst, D, (__half*)out, map, B, tok_h, tok_w, ws, g, (cudaStream_t)stream);
}

CVB_API int cvb_op_attention(const void* qkv, int Gb, int S, int heads, int hd, float scale, const void* Rh,
                             const void* Rw, int gh, int gw, void* out, void* stream) {
    return op_attention((const __half*)qkv, Gb, S, heads, hd, scale, (const __half*)Rh, (const __half*)Rw, gh, gw, (__half*)out,
                        (cudaStream_t)stream);
}

CVB_API int cvb_op_attention_tc_workspace_bytes(int Gb, int S, int heads, size_t* out) {
    CVB_CHECK(out != nullptr && Gb > 0 && S > 0 && heads > 0, CVB_EARG, "cvb_op_attention_tc_workspace_bytes: bad arguments");
    *out = op_attention_tc_workspace_bytes(Gb, S, heads);
    return CVB_OK;
}
CVB_API int cvb_op_attention_tc(const void* qkv, int Gb, int S, int heads, int hd, float scale, const void* Rh, const void* Rw,
                                int gh, int gw, void* out, void* workspace, size_t ws_bytes, void* stream) {
    return op_attention_tc((const __half*)qkv, Gb, S, heads, hd, scale, (const __half*)Rh, (const __half*)Rw, gh, gw, (__half*)out,
                           workspace, ws_bytes, (cudaStream_t)stream);
}

CVB_API int cvb_op_window_attention_tc(const void* qkv, int n_items, int heads, int hd, float scale, const void* relcat, void* out,
                                       int* sched_counter, void* stream) {
    return op_window_attention_tc((const __half*)qkv, n_items, heads, hd, scale, (const __half*)relcat, (__half*)out, sched_counter,
                                  (cudaStream_t)stream);
}

CVB_API int cvb_op_patch_im2col(const float* x, int B, int H, int W, int P, void* out, void* stream) {
    return op_patch_im2col(x, B, H, W, P, (__half*)out, (cudaStream_t)stream);
}

CVB_API int cvb_op_stem_conv(const float* x, int B, int H, int W, const float* w, const float* scale, const float* shift,
                             void* out, int cpad, void* stream) {
    return op_stem_conv(x, B, H, W, w, scale, shift, (__half*)out, cpad, (cudaStream_t)stream);
}

CVB_API int cvb_op_copy_planes(const void* src, int sH, int sW, void* dst, int dH, int dW, long long planes, int elem_bytes, void* stream) {
    return op_copy_planes(src, sH, sW, dst, dH, dW, planes, elem_bytes, (cudaStream_t)stream);
}

CVB_API int cvb_op_zero_margin(void* buf, long long planes, int H, int W, int bytes_per_px, int vH, int vW, void* stream) {
    return op_zero_margin(buf, planes, H, W, bytes_per_px, vH, vW, (cudaStream_t)stream);
}

CVB_API int cvb_tc_epilogue_bytes(void) { return (int)sizeof(TcEpilogue); }
