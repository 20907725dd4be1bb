// common.cuh -- shared helpers for the sm_100a kernels (error plumbing + PTX wrappers).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

// ---------------------------------------------------------------- status codes (include/cellvit_b200.h)
#define CVB_OK 0
#define CVB_EARG -1
#define CVB_ESHAPE -2
#define CVB_ECUDA -3
#define CVB_EWORKSPACE -4
#define CVB_EOVERFLOW -5

void cvb_set_error(const char* fmt, ...);

#define CVB_CUDA(call)                                                                     \
    do {                                                                                   \
        cudaError_t _e = (call);                                                           \
        if (_e != cudaSuccess) {                                                           \
            cvb_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return CVB_ECUDA;                                                              \
        }                                                                                  \
    } while (0)

#define CVB_CHECK(cond, code, ...)            \
    do {                                      \
        if (!(cond)) {                        \
            cvb_set_error(__VA_ARGS__);       \
            return (code);                    \
        }                                     \
    } while (0)

#define CVB_TRY(expr)                  \
    do {                               \
        int _rc = (expr);              \
        if (_rc != CVB_OK) return _rc; \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int cvb_num_sms();
int cvb_current_device();
void cvb_note_launches(int n);
// 2-D fp16 tensor map, 128B swizzle, zero fill outside the tensor (core.cu); `tm` is a CUtensorMap*
int cvb_tmap_2d_f16(void* tm, const void* base, uint64_t cols, uint64_t rows, uint64_t row_stride_bytes, uint32_t box_cols,
                    uint32_t box_rows);

#ifdef __CUDACC__
// ---------------------------------------------------------------- device-side PTX wrappers
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// One lane of the (fully active) warp: lets a whole warp run a producer / MMA-issue loop in uniform control flow -- so
// that descriptors and coordinates live in uniform registers -- while a single thread issues the asynchronous op.
// (A loop entered by one lane only makes ptxas wrap every UTCHMMA / UTMALDG in an ELECT + R2UR "waterfall".)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps after ~2 s instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("cellvit_b200: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x,
                   threadIdx.x, bar, parity);
            __trap();
        }
    }
}

// ---- TMA
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tmap) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"((uint64_t)tmap), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"((uint64_t)tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// L2 prefetch of a tile (no shared-memory destination, no barrier): warms the line fill ahead of the real load
__device__ __forceinline__ void tma_prefetch_2d(const void* tmap, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"((uint64_t)tmap), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d(const void* tmap, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];"
                 ::"l"((uint64_t)tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// ---- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], fp16 operands, fp32 accumulate, issued by one thread.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same with the A operand in tensor memory (128 lanes = rows, fp16 pairs packed in 32-bit columns: 8 columns per K = 16 step)
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier once all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives lane (base_lane+i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// register -> TMEM stores (same lane / column mapping as the loads)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
          "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
          "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- CTA-pair (cta_group::2) variants: two CTAs of a cluster share one 256-row MMA tile
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> leader CTA
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta_rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(bar), "r"(cta_rank)
        : "memory");
}
// cluster-scope release / acquire variants (dynamic tile scheduler: a 4-byte tile index travels with the barrier)
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t bar, uint32_t cta_rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(bar), "r"(cta_rank)
        : "memory");
}
__device__ __forceinline__ void st_shared_remote_u32(uint32_t addr, uint32_t cta_rank, uint32_t v) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "st.shared::cluster.u32 [ra], %2;\n\t}"
        ::"r"(addr), "r"(cta_rank), "r"(v)
        : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_acq_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_acq_cluster(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait_acq_cluster(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait_acq_cluster(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("cellvit_b200: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void st_shared_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_shared_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

// ---- dynamic tile scheduler of the persistent kernels. One warp claims tile indices from a global counter (zeroed by the
// host before the launch) and publishes them through a ring of SCHED_DEPTH slots in shared memory; every other role of the
// CTA (or CTA pair) consumes the same sequence. A persistent kernel whose CTA i walks tiles i, i + grid, ... takes twice as
// long as soon as ONE of its CTAs starts late -- which is what happens whenever a CTA of the concurrent post-processing
// stream occupies an SM at the kernel boundary; with claimed tiles the late CTA simply finds less work.
constexpr int SCHED_DEPTH = 4;
constexpr int SCHED_BYTES = 2 * SCHED_DEPTH * (8 + 8 + 4) + 8 + 8;
struct SchedRing {
    // shared-memory addresses: full[SCHED_DEPTH], empty[SCHED_DEPTH] barriers, the `go` barrier, int tile[SCHED_DEPTH], and the
    // same again (lfull, lempty, ltile) for the peer CTA's local copy of the sequence in CTA-pair mode
    uint32_t full, empty, go, tile, lfull, lempty, ltile;
    __device__ __forceinline__ void carve(uint32_t base) {
        full = base; empty = base + 8u * SCHED_DEPTH; go = base + 16u * SCHED_DEPTH; tile = go + 8u;
        lfull = tile + 4u * SCHED_DEPTH; lempty = lfull + 8u * SCHED_DEPTH; ltile = lempty + 8u * SCHED_DEPTH;
    }
    // one thread, before the CTA-wide barrier. consumers = consumer WARPS of this CTA (one elected arrival each);
    // pair: one more arrival per slot from the peer CTA's relay warp
    __device__ __forceinline__ void init(int consumers, bool pair) {
        for (int s = 0; s < SCHED_DEPTH; ++s) {
            mbar_init(full + 8u * s, 1); mbar_init(empty + 8u * s, (uint32_t)(consumers + (pair ? 1 : 0)));
            mbar_init(lfull + 8u * s, 1); mbar_init(lempty + 8u * s, (uint32_t)consumers);
        }
        mbar_init(go, 1);
    }
    // producer warp of the leader CTA (all lanes): claims until the counter runs out; publishes -1 as the terminator.
    // Everything the role warps of a CTA touch stays at CTA scope: a cluster-scope acquire on every wait costs several
    // hundred cycles per tile and role (measured: +7..15 % on the kernels' duration). In CTA-pair mode the index travels to
    // the peer CTA once per tile (remote store + cluster-scope release), where an otherwise idle warp relays it (relay()).
    template <bool PAIR>
    __device__ __forceinline__ void produce(int* counter, int n_tiles) {
        const int lane = threadIdx.x & 31;
        for (int k = 0;; ++k) {
            const int slot = k & (SCHED_DEPTH - 1);
            const uint32_t eph = (uint32_t)(((k / SCHED_DEPTH) & 1) ^ 1);
            if (PAIR) mbar_wait_acq_cluster(empty + 8u * slot, eph); else mbar_wait(empty + 8u * slot, eph);
            // claim tile k only once the leading consumer (the load warp) has STARTED tile k - 1: a CTA never holds more than one
            // claimed-but-unstarted tile, so the tail of the kernel balances like the static lists do, and the claim's
            // round trip to L2 still has a whole tile to complete
            if (k >= 1) mbar_wait(go, (uint32_t)((k - 1) & 1));
            int t = 0;
            if (lane == 0) t = atomicAdd(counter, 1);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t >= n_tiles) t = -1;
            if (lane == 0) {
                st_shared_u32(tile + 4u * slot, (uint32_t)t);
                mbar_arrive(full + 8u * slot);
                if (PAIR) {
                    st_shared_remote_u32(tile + 4u * slot, 1, (uint32_t)t);
                    mbar_arrive_remote_release(full + 8u * slot, 1);
                }
            }
            __syncwarp();
            if (t < 0) break;
        }
    }
    // peer CTA of a pair (all lanes of one idle warp): receives the leader's sequence at cluster scope, frees the leader's slot,
    // and republishes it to the peer's own role warps at CTA scope
    __device__ __forceinline__ void relay() {
        const int lane = threadIdx.x & 31;
        for (int k = 0;; ++k) {
            const int slot = k & (SCHED_DEPTH - 1);
            const uint32_t ph = (uint32_t)((k / SCHED_DEPTH) & 1);
            mbar_wait_acq_cluster(full + 8u * slot, ph);
            const int t = (int)ld_shared_u32(tile + 4u * slot);
            __syncwarp();
            mbar_wait(lempty + 8u * slot, ph ^ 1u);
            if (lane == 0) {
                mbar_arrive_remote_release(empty + 8u * slot, 0);
                st_shared_u32(ltile + 4u * slot, (uint32_t)t);
                mbar_arrive(lfull + 8u * slot);
            }
            __syncwarp();
            if (t < 0) break;
        }
    }
    // role warp (all lanes): k-th tile of the sequence, -1 when the work is exhausted. `leads`: this is the load warp of
    // the leader CTA, whose progress paces the claims (see produce)
    __device__ __forceinline__ int consume(int k, uint32_t rank, bool leads) {
        const int slot = k & (SCHED_DEPTH - 1);
        const uint32_t ph = (uint32_t)((k / SCHED_DEPTH) & 1);
        const uint32_t f = rank == 0 ? full : lfull, e = rank == 0 ? empty : lempty, tl = rank == 0 ? tile : ltile;
        mbar_wait(f + 8u * slot, ph);
        const int t = (int)ld_shared_u32(tl + 4u * slot);
        __syncwarp();
        if (elect_one()) {
            mbar_arrive(e + 8u * slot);
            if (leads) mbar_arrive(go);
        }
        return t;
    }
};

__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"((uint64_t)tmap), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2,
                                                 int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"((uint64_t)tmap), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive (once all previously issued MMAs completed) on the barrier at this offset in every CTA of `cta_mask`.
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(cta_mask) : "memory");
}

// ---- Ampere-style pieces still useful on sm_100a (attention kernel)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
    int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
// 2^x, single MUFU (exp2f without -use_fast_math adds range/denormal handling: ~8 instructions)
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// two 2^x in one MUFU op, fp16 in / fp16 out (softmax weights that are rounded to fp16 for the P V MMA anyway)
__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t x) {
    uint32_t y;
    asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t addr, uint32_t& a, uint32_t& b) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

}  // namespace ptx

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// Exact-erf GELU with a branch-free erf (Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7 -- far below the fp16
// rounding of the stored activation): ~15 instructions instead of erff's ~35, which matters because the MLP
// epilogue applies it to 128 x 5120 values per tile row block.
__device__ __forceinline__ float gelu_erf_fast(float x) {
    // gelu(x) = relu(x) - |x|/2 * erfc(|x|/sqrt2), erfc(z) = poly(t) * t * exp(-z^2), t = 1 / (1 + p z): 15 issue slots
    const float z = fabsf(x) * 0.70710678118654752440f;  // |x| / sqrt(2); |x| / 2 = z * 0.7071...
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    float ex;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-1.4426950408889634f * z * z));
    const float y = (p * t) * ex;                          // erfc(z)
    return fmaf(-(z * 0.70710678118654752440f), y, fmaxf(x, 0.0f));
}

// Two softmax weights as one packed fp16 pair for a P V MMA: returns (fp16(2^(y0 + 112)), fp16(2^(y1 + 112))) for arguments that the
// caller has ALREADY rebiased by -112 (y <= -97, i.e. weights up to 2^15). `ex2.approx.f16x2` is two MUFU ops plus a conversion
// (F2FP) and a PRMT on this chip, and the F2FP shares the MUFU's issue queue, which made the exponential passes of the
// attention kernels queue-bound. Here: fp32 MUFU, whose result 2^(y) has the fp16 exponent in the fp32 exponent field, so
// (bits + 0x1000) >> 13 IS the fp16 encoding (round half up) -- integer ops on the ALU / FMA pipes instead of a third queue
// slot. fp32 MUFU flushes below 2^-126, i.e. weights below 2^-14: callers scale their weights by 2^8 (maximum 256, sums in
// fp32) so that the flush threshold is 2^-22 of the maximum -- fp16 denormals of unscaled weights would still have carried the
// collective mass of a peaked row's tail (measured: 2x the error without the scaling).
__device__ __forceinline__ uint32_t exp2_pair_f16(float y0, float y1) {
    const uint32_t b0 = __float_as_uint(ptx::ex2(y0)), b1 = __float_as_uint(ptx::ex2(y1));
    return ((b0 + 0x1000u) >> 13) | (((b1 + 0x1000u) << 3) & 0xFFFF0000u);
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
#endif  // __CUDACC__
