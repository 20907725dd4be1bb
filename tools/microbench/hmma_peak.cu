// Peak rate of the legacy mma.sync.m16n8k16 (fp16 in, fp32 accumulate) path on sm_100a: decides whether the
// attention kernels (mma.sync based) are near their tensor ceiling. Build: nvcc -arch=sm_100a -O3 hmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void k(float* out, int iters) {
    float d[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 0x3c003c00u, 0x3c003c00u}, b0 = 0x3c003c00u, b1 = 0x38003800u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 16 * 1024 * 4);
    for (int warps = 4; warps <= 32; warps *= 2) {
        int iters = 20000;
        k<<<148, warps * 32>>>(out, 100); cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0); k<<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 148.0 * warps * iters * 8 * (16.0 * 8 * 16 * 2);
        printf("warps/SM %2d: %.1f TFLOP/s (%.0f FMA/clk/SM at 1.9 GHz)\n", warps, flops / ms / 1e9, flops / 2 / (ms * 1e-3) / 148 / 1.9e9);
    }
    return 0;
}
