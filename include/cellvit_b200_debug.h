/*
 * cellvit_b200_debug.h -- test / profiling hooks of libcellvit_b200.so. NOT part of the drop-in boundary (cellvit_b200.h).
 *
 * Everything here is PROCESS-GLOBAL state or instrumentation: ablation switches for the parity tests, counters and
 * timers for bench.py, a clock-stamp trace for one kernel. A production caller never needs them; models configure their
 * engine per handle through cvb_model_set_option (cellvit_b200.h). Not thread-safe.
 */
#ifndef CELLVIT_B200_DEBUG_H
#define CELLVIT_B200_DEBUG_H

#ifdef __cplusplus
extern "C" {
#endif

/* Number of kernels this library has launched since the last reset (bench.py reports it as gpu_launches). */
long long cvb_launch_count(int reset);

/* Per-launch timing of the tile engine (tc_kernel / conv_patch_kernel): CUDA event pairs around every launch on the
 * launching stream between begin and end. end synchronises the events; total_ms = sum of the launch durations,
 * flops = executed 2*M*N*K summed (bench.py roofline leg). */
int cvb_tc_profile_begin(int max_launches);
int cvb_tc_profile_end(double* total_ms, int* n_launches, double* flops);

/* Ablation switches of the tile engine (op level, tests/test_gpu_tc.py):
 *   pair_mode 0 disables the CTA-pair (cta_group::2) path; conv_patch_mode 0 forces the k-block convolution;
 *   max_ctas restricts the persistent grid (0 = all SMs); l2_prefetch = A k-blocks prefetched into L2 ahead of the ring. */
void cvb_tc_set_pair_mode(int on);
void cvb_tc_set_conv_patch_mode(int mode);
void cvb_tc_set_max_ctas(int n);
void cvb_tc_set_l2_prefetch(int k);

/* Timing experiments on the post-processing (bench.py with CVB_SKIP_FLOOD / CVB_POST_CTAS): skip the watershed floods (label maps
 * are then incomplete), cap the grid of the map kernels. */
void cvb_debug_postproc_skip_flood(int on);
void cvb_debug_postproc_max_ctas(int n);
/* Grid of the big-blob flood launch (200 KB of shared memory per CTA; 0 = one CTA per SM). */
void cvb_debug_flood_large_ctas(int n);

/* Clock-stamp timeline of window_tc_kernel (tools/trace_window.py): dev_buf = 8 x 64 int64 on the device, or NULL to stop. */
void cvb_debug_window_trace(void* dev_buf);
/* Clocks by which query group 1 of window_tc_kernel is delayed once, to run the two groups in anti-phase (default 3000; 0 = lockstep). */
void cvb_debug_window_skew(int clocks);

#ifdef __cplusplus
}
#endif
#endif /* CELLVIT_B200_DEBUG_H */
