// tc_gemm.h -- host API of the tcgen05/TMA tile engine (GEMM and implicit-GEMM 3x3 convolution).
#pragma once
#include "common.cuh"

enum TcEpiKind {
    TC_EPI_F16 = 0,     // out f16 [rows, ldc] = act(acc*scale[n] + shift[n])
    TC_EPI_RES_F32 = 1, // out f32 [rows, ldc] = acc + shift[n] + res[res_row, n]
    TC_EPI_CONVT = 2,   // ConvTranspose2d k2 s2 scatter: N = 4*Cout ordered (dy,dx,co), out NHWC f16 (+shift[co])
    TC_EPI_HEAD = 3,    // N == 64: relu(acc*scale+shift) then fused 1x1 head -> f32 NCHW planes
};
enum TcAct { TC_ACT_NONE = 0, TC_ACT_RELU = 1, TC_ACT_GELU = 2 };
enum TcRowMap {
    TC_ROW_IDENTITY = 0,
    TC_ROW_SEQ = 1,     // out row = m + (m / row_seq) * row_pad + row_off          (ViT-256 cls-token slot)
    TC_ROW_WINDOW = 2,  // m indexes window-partitioned tokens; out row = (b,y,x) raster, padded tokens dropped
    TC_ROW_TO_WINDOW = 3,  // m indexes (b,y,x) raster tokens; out row = their window-partitioned position (window_partition,
                           // image_encoder.py:263-288) -- the padded rows in between are not touched
};

// Plain-old-data: mirrored field by field by ctypes in tests (cellvit_b200/_lib.py).
struct TcEpilogue {
    int kind;
    int act;
    const float* scale;  // [N] or null (== 1)
    const float* shift;  // [N] or null (== 0)   (CONVT: [Cout])
    void* out;
    long long ldc;       // elements per output row
    const float* res;    // RES_F32: residual / additive table (may alias out)
    long long ldres;
    int res_mod;         // 0: res row == out row; >0: res row = (m % res_mod) + res_off
    int res_off;
    int row_map;         // TcRowMap
    int row_seq, row_pad, row_off;
    int win_size, win_grid, tok_h, tok_w;  // TC_ROW_WINDOW / TC_ROW_TO_WINDOW: window edge, windows per side, token grid
    int ct_cout, ct_hin, ct_win;           // CONVT geometry (input grid)
    const float* head_w;                   // HEAD: [nc, 64]
    const float* head_b;                   // HEAD: [nc]
    int head_nc;
    int head_hw;                           // HEAD: H*W of one image
    float* head_out;                       // HEAD: [NB, nc, H, W]
    uint8_t* head_argmax;                  // HEAD, optional: [NB, H, W] arg-max over the first head_argmax_nc classes (first maximum wins)
    int head_argmax_nc;
    int* sched_counter;                    // optional: device int, ZERO at launch -> tiles are claimed dynamically (SchedRing, common.cuh)
};

// C[M,N] = A[M,K] * W[N,K]^T.  A fp16 row-major (lda elements), W fp16 row-major [N, ldw] (K-major).
// N % block_n == 0, block_n % 16 == 0, 16 <= block_n <= 256; K % 64 == 0; lda, ldw multiples of 8.
int tc_gemm(const __half* A, int M, int K, long long lda, const __half* W, int N, long long ldw, int block_n,
            const TcEpilogue& epi, cudaStream_t stream);

// 3x3 / pad 1 / stride 1 convolution over NHWC fp16 activations, optionally over the channel-concatenation
// of two sources (src1 may be null). Weights packed [N, 9*(C0+C1)] with k = tap*(C0+C1) + channel.
// C0, C1 multiples of 64. Output rows are pixels (n,y,x) in raster order.
int tc_conv3x3(const __half* src0, int C0, const __half* src1, int C1, int NB, int H, int W, const __half* Wp,
               int N, int block_n, const TcEpilogue& epi, cudaStream_t stream);

// Largest block_n (multiple of 16, <= 256) that divides n; 0 if none.
int tc_pick_block_n(int n);
