/*
 * nl_oracle.h — CPU restatement of numpyllvm's pipeline execution.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: it
 * may be imported / linked / executed only by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs, and there only as the
 * checker or the timed CPU baseline.  The product path (libnlkernels, jit)
 * never routes through it and has no CPU fallback.
 *
 * Parity status: PINNED for int64 `*`, `>=`, `[mask]`, `sum`, scalar operands,
 * materialised intermediates and the 8-chunk gather by the golden vectors of
 * the reference's own Tests scripts (tests/golden/).  "PARITY UNPINNED" (no
 * reference test or fixture exists; pinned against NumPy only) for: max/min
 * (never called by a reference test and unfinished), every float
 * dtype, int32, `> <= < == !=`, `+ -`, boolean ops and masked assignment.
 *
 * The reference itself cannot be built in this image (CPython 2 C-API,
 * LLVM 3.9 ORC headers, macOS GCD semaphores — see DESIGN.md), so there is no
 * oracle/_ref.
 */
#ifndef NL_ORACLE_H
#define NL_ORACLE_H

#include "nl_abi.h"   /* struct definitions only */

#ifdef __cplusplus
extern "C" {
#endif

/* Run one pipeline the way the reference does:
 *   ScheduleFunction  (scheduler.cpp:92-149): zero-filled upper-bound outputs,
 *       [0,size) cut into `thread_count` contiguous chunks (reference: 8);
 *   the JIT'd loop    (compiler.cpp:309-538) per chunk, evaluated by recursion
 *       over the operation tree exactly like PerformOperation
 *       (compiler.cpp:202-269);
 *   JITFunctionDECREF (compiler.cpp:18-79): memmove compaction of the chunk
 *       runs, or finalize_data for reductions (thunk_methods.cpp:133-151).
 * n_workers = 1 is the faithful configuration (create_threads starts ONE
 * worker, scheduler.cpp:227-231); > 1 runs the chunks on that many pthreads.
 *
 * outputs[k]: caller-allocated, `size` elements of the output dtype (1 element
 * for a reduction).  NL_OP_WHERE_STORE outputs are updated in place and are
 * not zero-filled.  out_counts[k] receives the patched DIMS[0]
 * (compiler.cpp:48).  Returns NL_OK or a negative nl error code. */
int nlo_run(const nl_expr_node *nodes, int n_nodes,
            const nl_input_desc *in, int n_inputs,
            const void *const *inputs, void *const *outputs, int n_outputs,
            int64_t size, int thread_count, int n_workers, int64_t *out_counts);

/* In-place assignment restatement: NumPy setitem on the storage
 * (thunk_as_mapping.cpp:84) for the index forms the engine implements. */
int nlo_assign_fill(void *dst, int dtype, int64_t lo, int64_t step, int64_t count, uint64_t value_bits);
int nlo_assign_copy(void *dst, int dtype, int64_t lo, int64_t step, int64_t count, const void *src);

/* The reference partition (scheduler.cpp:133-147). */
void nlo_partition(int64_t size, int thread_count, int j, int64_t *start, int64_t *end);

/* ---- compiled fixed-shape loops: what LLVM -O3 would make of the benchmark
 *      pipelines; used as the timed CPU baseline and checked against nlo_run
 *      in tests.  All use the reference's chunking + gather. ---------------- */
int nlo_fast_mul2_f32(const float *a, const float *b, float *c, int64_t n, int thread_count, int n_workers);
int nlo_fast_mul6_f64(const double *const *in6, double *out, int64_t n, int thread_count, int n_workers);
int nlo_fast_filter_gt_f64(const double *a, double t, double *out, int64_t n, int thread_count, int n_workers, int64_t *count);
int nlo_fast_max_f64(const double *a, int64_t n, int thread_count, int n_workers, double *out);
int nlo_fast_max_i64(const int64_t *a, int64_t n, int thread_count, int n_workers, int64_t *out);
int nlo_fast_sum_i64(const int64_t *a, int64_t n, int thread_count, int n_workers, int64_t *out);
int nlo_fast_filter_mul_max_f32(const float *a, float t, float k, int64_t n, int thread_count, int n_workers, float *out);

/* Deterministic synthetic inputs shared with the device generator
 * (BASELINE.md §4): x[i] = f(splitmix64(seed ^ i)).
 * kind 0: ints uniform in [1,100) ; 1: full-range ints ; 2: floats in [1,2) ;
 * 3: floats in [0,1). */
int nlo_fill(void *dst, int dtype, int64_t first_index, int64_t count, uint64_t seed, int kind);

#ifdef __cplusplus
}
#endif
#endif
