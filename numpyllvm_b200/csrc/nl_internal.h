// nl_internal.h — declarations shared by the translation units of libnlkernels.
#pragma once
#include <stdint.h>
#include <stddef.h>

#include "nl_abi.h"

namespace nl {

void set_error(const char *fmt, ...);

// elements of one "wide" row of a thread: 16 bytes for 4- and 8-byte types; the 1- and 2-byte types use rows of
// four elements (one predicate bit per element and 8-bit packed row counters bound a tile to 32 elements per thread)
inline int nl_vec_elems(int dtype) {
  const size_t s = nl_dtype_size(dtype);
  return s >= 4 ? (int)(16 / s) : 4;
}

// every value type the pipeline kernels are instantiated for: X(nl_dtype, C type, suffix)
#define NL_FOR_VALUE_DTYPES(X)                                                                       \
  X(NL_I32, int32_t, i32) X(NL_I64, long long, i64) X(NL_F32, float, f32) X(NL_F64, double, f64)     \
  X(NL_I8, signed char, i8) X(NL_I16, short, i16) X(NL_U8, unsigned char, u8) X(NL_U16, unsigned short, u16) \
  X(NL_U32, unsigned int, u32) X(NL_U64, unsigned long long, u64)

// ---- kernel parameters of the pipeline kernel (passed as __grid_constant__) -------
struct KParams {
  const void *in[NL_MAX_INPUTS];
  uint64_t imm[NL_MAX_INPUTS];
  void *out[NL_MAX_OUTPUTS];
  int64_t start, end;         // element range of this launch (indices into in[]/out[])
  int64_t n_tiles;
  int64_t *out_sizes;         // device, n_outputs entries
  // scratch (per device+stream arena)
  unsigned long long *tile_desc;  // decoupled look-back descriptors, n_tiles entries, zeroed
  unsigned int *tile_ticket;      // dynamic tile counter, zeroed
  unsigned int *sm_rank;          // per SM (%smid): CTAs of this launch that have started there, zeroed (compaction)
  // super-tile kernel: the arena (ticket, ranks, descriptors) the PREVIOUS launch used; this launch zeroes it while it starts
  unsigned long long *clean_desc;
  unsigned int *clean_ticket;
  unsigned int *clean_rank;
  int64_t clean_count;            // descriptors to zero in clean_desc (0: nothing to clean)
  unsigned long long *red_partials;  // one 8-byte partial per CTA
  unsigned int *red_ticket;       // self-resetting
  uint32_t step[NL_MAX_STEPS];
  int32_t n_steps;
  int32_t n_outputs;
  int32_t thread_nr;
  int32_t reduce_rop;             // ROp of the root reduction
  int32_t reduce_out;             // its output slot
  uint8_t in_kind[NL_MAX_INPUTS];
  uint8_t out_space[NL_MAX_OUTPUTS];
  // staged compaction pipeline (nl_compact_pipe.cuh)
  uint8_t stage_col[NL_MAX_INPUTS];   // input k -> staged column index, 0xff: read from global
  int32_t n_stage_cols;
  int32_t n_inputs;
  int32_t n_slots;                    // ring depth
  int32_t lag;                        // tiles between the count phase and the store phase
  int32_t filter_step;                // index of the S_FILTER step
  int32_t side_cap;                   // survivors per warp and tile that may wait in the side buffer (0: none)
  int32_t pipe_rows;                  // 16-byte rows per consumer thread and tile of the chosen instantiation
  int32_t super_k;                    // > 0: super-tile compaction kernel (nl_compact_super.cuh), tiles per super-tile
  int32_t super_rows;                 // ... and its rows per thread
  int32_t super_keep_ctas;            // a CTA that was not among the first `super_keep_ctas` on its SM retires when it meets a dense super-tile
  uint32_t poll_sleep_ns;             // back-off between two polls of a look-back descriptor (0: none)
  int32_t store_prefix;               // the store pass re-runs the value steps before the filter (nl_step.h: store_pass_needs_prefix)
  // fused multi-GPU finish of a reduction over peer memory (NVLink mailboxes, nl_pipeline.cuh)
  unsigned long long *const *peer_box;   // device array: mailbox of every rank, mapped here (NULL: no fused finish)
  int32_t peer_world, peer_rank;
  unsigned long long peer_epoch;      // > 0, the same on every rank for one reduction
  long long *peer_offsets;            // filters: exclusive scan of every rank's count, peer_world + 1 entries (or NULL)
  unsigned long long peer_timeout_ns; // bound on every wait for a peer's slot
  unsigned long long *peer_status;    // pinned host word of this device, raised when an exchange was given up
};

// geometry of the pipeline kernel
constexpr int kBlock = 256;       // threads per CTA (8 warps)
constexpr int kRows = 4;          // U: vectors per thread per tile
#ifndef NL_DESC_STRIDE
#define NL_DESC_STRIDE 16          // 8-byte words between consecutive tile descriptors (16 = one per 128-byte line)
#endif
constexpr int kMinBlocksCompact = 4;  // resident CTAs per SM of the build-time compaction shapes
#ifndef NL_ROWS_STREAM
#define NL_ROWS_STREAM 8
#endif
constexpr int kRowsStream = NL_ROWS_STREAM;    // U of the build-time elementwise / reduction shapes
constexpr int kRowsCompact = 8;   // U of the build-time compaction shapes
#ifndef NL_MIN_BLOCKS
#define NL_MIN_BLOCKS 2
#endif
constexpr int kMinBlocksLight = 3; // ... of the interpreter variant without value temporaries
constexpr int kMinBlocks = NL_MIN_BLOCKS;     // resident CTAs per SM the register allocation must allow

struct LaunchInfo {
  char kernel[96];
  int grid, block, vec_elems;
};

// nl_kernels.cu
int launch_pipeline_kernel(const nl_plan *plan, const KParams &kp, bool vec16, int sm_count,
                           void *stream /*cudaStream_t*/, LaunchInfo *info);
// elements per tile of the kernel that nl_launch_pipeline would select for this plan and size;
// fills the staged-pipeline fields of *kp when that kernel is chosen
int64_t pipeline_tile_elems(const nl_plan *plan, bool vec16, int64_t n_elems, KParams *kp, int sm_count);
int launch_finalize_partials(void *stream, int rop, int dtype, const void *partials, int n, void *out);
int launch_assign_fill(void *stream, int sm_count, void *dst, int dtype, int64_t lo, int64_t step, int64_t count, uint64_t bits);
int launch_assign_copy(void *stream, int sm_count, void *dst, int dtype, int64_t lo, int64_t step, int64_t count, const void *src);
int launch_scan_offsets(void *stream, const int64_t *counts, int n, int64_t *offsets);
int launch_fill(void *stream, int sm_count, void *dst, int dtype, int64_t first_index, int64_t count, uint64_t seed, int kind);
int launch_empty_range(void *stream, const nl_plan *plan, const KParams &kp);
int launch_gather(void *stream, int sm_count, int dtype, void *out, const int64_t *idx, int64_t count, const void *const *shard_ptr,
                  const int64_t *shard_start, int n_shards, int64_t total, int64_t *bad);
// nl_cast.cu
int launch_cast(void *stream, int sm_count, void *dst, int dst_dtype, const void *src, int src_dtype, int64_t n);
// nl_checksum.cu
int launch_checksum(void *stream, int sm_count, const void *ptr, size_t nbytes, unsigned long long *out);
// nl_sort.cu
int launch_sort(void *stream, int sm_count, int dtype, void *keys, void *tmp, int64_t n);
int launch_sort_sample(void *stream, int dtype, const void *keys, int64_t n, int k, unsigned long long *host_out);
int launch_sort_split(void *stream, int sm_count, int mode, int dtype, const void *keys, void *out, int64_t n, const unsigned long long *splitters, int m,
                      int64_t *host_counts, void *const *run_dst);
constexpr int kSortMaxSplitters = 15;
int describe_pipeline_kernel(const nl_plan *plan, bool vec16, char *name, size_t name_len, int *is_static);
void set_static_kernels_enabled(int on);
void set_compact_pipe_enabled(int on);
void set_compact_pipe_side(int on);
void set_compact_pipe_dynamic(int on);
void set_compact_prefer_pipe(int on);
void set_max_ctas_per_sm(int n);
void set_poll_sleep_ns(int v);
void set_sort_algo(int a);
void set_sort_portion_log2(int l);
void set_compact_pipe_lag(int l);

void count_launch(int n = 1);

}  // namespace nl
