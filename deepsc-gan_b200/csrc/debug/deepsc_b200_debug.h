/*
 * deepsc_b200_debug.h - entry points of libdeepsc_b200_debug.so (built by `python deepsc-gan_b200/build.py --debug`,
 * i.e. the product sources compiled with -DDSC_DEBUG_TOOLS=1 plus csrc/debug/*.cu).  Developer tools only: nothing here
 * is part of the product ABI (include/deepsc_b200.h) and libdeepsc_b200.so does not export these symbols.
 */
#ifndef DEEPSC_B200_DEBUG_H
#define DEEPSC_B200_DEBUG_H
#ifdef __cplusplus
extern "C" {
#endif

/* Micro-benchmark: cycles for iters x 8 back-to-back tcgen05.mma (M = 128, K = 16, width n) with the A operand in
 * tensor memory (ts_mode != 0) or shared memory; result in cycles_dev[0]. */
int dsc_umma_probe(int ts_mode, int n, int iters, long long* cycles_dev, void* stream);

/* Timeline of dsc_star_cycles_tc: registers a device buffer of 768 uint64 (NULL = off); CTA 0 of every later launch
 * stamps clock64() at the hand-offs of its first tile (layout: dsc_star_fused.cu). */
int dsc_debug_star_trace(void* device_buffer_768_u64);

#ifdef __cplusplus
}
#endif
#endif
