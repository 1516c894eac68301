/*
 * nl_abi.h — C ABI of libnlkernels, the B200 (sm_100a) execution engine for
 * numpyllvm's lazy thunk pipelines.
 *
 * This header is the drop-in boundary.  It replaces the seam where the
 * reference hands a parsed Pipeline to LLVM and later calls the JIT'd loop
 * (all citations are relative to the reference checkout):
 *
 *   reference interface                              replaced by
 *   -----------------------------------------------  -------------------------
 *   CompilePipeline()            compiler.cpp:309    nl_plan_compile()
 *   CompilableOperation()        compiler.cpp:271    nl_plan_compile() rc
 *   jit_function call            compiler.cpp:87     nl_launch_pipeline()
 *     (typedef thread.hpp:25)
 *   ScheduleFunction() alloc     scheduler.cpp:107   nl_malloc()/nl_memset()
 *   ScheduleFunction() partition scheduler.cpp:133   nl_partition()
 *   JITFunctionDECREF() gather   compiler.cpp:30-49  nl_launch_pipeline_allgather(),
 *                                                    nl_finish_filter()
 *   llvm_finalize_sum_data()     thunk_methods.cpp:133  nl_launch_pipeline_allreduce(),
 *                                                    nl_finish_reduce(),
 *                                                    nl_finalize_partials()
 *   thunk_assign_subscript()     thunk_as_mapping.cpp:80  nl_assign_fill(),
 *     (NumPy setitem on storage)                     nl_assign_copy()
 *   _thunk_sort base (NumPy)     thunk_methods.cpp:37  nl_sort()
 *   PyThunk_FromArray() copy     thunk.cpp:30        nl_memcpy_h2d()
 *   PyThunk_AsArray()            thunk.cpp:119       nl_memcpy_d2h()
 *   create_threads()/RunThread   scheduler.cpp:161,227  nl_init() (streams)
 *
 * Conventions: plain C structs and pointers, `int` return codes (0 = NL_OK,
 * negative = error, message via nl_last_error()), no exceptions across the
 * boundary, the caller owns every buffer.  One host thread may drive all
 * local devices; calls are asynchronous with respect to the device unless the
 * name says sync.  There is NO CPU fallback: every compute entry point fails
 * with NL_ERR_NO_DEVICE when no sm_100 device was initialised.
 */
#ifndef NL_ABI_H
#define NL_ABI_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NL_ABI_VERSION 1

#if defined(__GNUC__)
#define NL_API __attribute__((visibility("default")))
#else
#define NL_API
#endif

/* ---- limits of one fused pipeline (over-long chains are split by the host
 *      plan optimiser with a materialised temporary) ------------------------ */
#define NL_MAX_INPUTS   8
#define NL_MAX_OUTPUTS  4
#define NL_MAX_NODES    48
#define NL_MAX_STEPS    64
#define NL_MAX_TEMPS    2     /* value temporaries of the accumulator machine */
#define NL_MAX_PTEMPS   3     /* predicate temporaries                        */
#define NL_MAX_DEVICES  16
#define NL_N_STREAMS    3     /* per device: 0 compute, 1 host->device, 2 device->host */

/* ---- error codes ---------------------------------------------------------- */
enum {
  NL_OK               = 0,
  NL_ERR_INVALID      = -1,  /* malformed argument / plan                      */
  NL_ERR_UNSUPPORTED  = -2,  /* well-formed but outside the kernel library     */
  NL_ERR_SPLIT        = -3,  /* expression exceeds NL_MAX_*: host must split   */
  NL_ERR_NO_DEVICE    = -4,  /* no CUDA device initialised (no CPU fallback)   */
  NL_ERR_CUDA         = -5,  /* a CUDA runtime call failed                     */
  NL_ERR_NCCL         = -6,  /* NCCL missing or a collective failed            */
  NL_ERR_NOMEM        = -7,
  NL_ERR_SPLIT_FILTER = -8   /* a column is combined with the result of a filter of the same
                                pipeline (a[m] * b[m], f[f < 10] with f = a[a > 5]): the host
                                must give the filters below the root pipelines of their own */
};

/* ---- element types (getLLVMType, compiler.cpp:156-182: bool/i8, i16, i32, i64, f32, f64; the
 *      reference leaves the unsigned types as a todo, :168-172 — here they compute with unsigned
 *      compares and wrap like NumPy's).  The build-time shapes and shape families exist for the
 *      first four value types; the others run on the step interpreter.  f16 is not implemented. */
typedef enum {
  NL_BOOL = 0,   /* NPY_BOOL, one byte, canonical 0x00/0x01 on store (Q4)      */
  NL_I32  = 1,
  NL_I64  = 2,
  NL_F32  = 3,
  NL_F64  = 4,
  NL_I8   = 5,
  NL_I16  = 6,
  NL_U8   = 7,
  NL_U16  = 8,
  NL_U32  = 9,
  NL_U64  = 10
} nl_dtype;
#define NL_N_DTYPES 11

/* ---- operators: the per-operator "emitters" of the reference -------------- */
typedef enum {
  NL_OP_INPUT  = 0,  /* leaf: ObjectOperation / PipelineOperation (parser.hpp:43-58) */
  /* T x T -> T */
  NL_OP_MUL    = 1,  /* llvm_generate_multiply, thunk_as_number.cpp:45 (ints wrap)   */
  NL_OP_ADD    = 2,
  NL_OP_SUB    = 3,
  NL_OP_DIV    = 4,  /* true division, float columns only (IEEE, correctly rounded)   */
  NL_OP_FLOORDIV = 5, /* NumPy floor_divide: ints floor (x // 0 == 0), floats npy_divmod */
  /* T x T -> bool */
  NL_OP_GE     = 8,  /* llvm_generate_ge, thunk_as_number.cpp:9 (signed / ordered)   */
  NL_OP_GT     = 9,
  NL_OP_LE     = 10,
  NL_OP_LT     = 11,
  NL_OP_EQ     = 12,
  NL_OP_NE     = 13,
  /* bool x bool -> bool, bool -> bool */
  NL_OP_AND    = 16,
  NL_OP_OR     = 17,
  NL_OP_XOR    = 18,
  NL_OP_NOT    = 19,
  /* lhs = value (T), rhs = mask (bool) -> T; skips the element when !mask and
   * redirects every later store to a compacted index
   * (llvm_generate_subscript, thunk_as_mapping.cpp:9-21) */
  NL_OP_FILTER = 24,
  /* T -> T, cardinality 1, pipeline root only ("parallelbreaker") */
  NL_OP_MAX    = 32, /* llvm_generate_max, thunk_methods.cpp:70 */
  NL_OP_MIN    = 33,
  NL_OP_SUM    = 34, /* llvm_generate_sum, thunk_methods.cpp:114 */
  /* lhs = value (T), rhs = mask (bool): `if (mask) out[i] = value` at the loop
   * index, root only — in-place masked assignment (thunk_as_mapping.cpp:80-86
   * delegated this to NumPy setitem) */
  NL_OP_WHERE_STORE = 40
} nl_op;

/* ---- one node of a pipeline's operation tree (parser.hpp:25-87).  Nodes are
 *      stored in evaluation (post) order: children precede parents; the last
 *      node is the pipeline root. -------------------------------------------- */
typedef struct nl_expr_node {
  int32_t op;       /* nl_op                                                   */
  int32_t lhs;      /* node index of LHS (unary: the operand), -1 for a leaf   */
  int32_t rhs;      /* node index of RHS, -1 when unary / leaf                 */
  int32_t input;    /* NL_OP_INPUT: index into the input table, else -1        */
  int32_t output;   /* >= 0: Operation::materialize (parser.hpp:37): the value
                       is stored to outputs[output]; -1: stays in registers    */
} nl_expr_node;

typedef enum {
  NL_IN_ARRAY      = 0,  /* DataSource with size > 1: streamed column          */
  NL_IN_SCALAR_IMM = 1,  /* size-1 source baked as an immediate
                            (getLLVMConstant, compiler.cpp:125-154)            */
  NL_IN_SCALAR_DEV = 2   /* size-1 source left on the device (result of a child
                            reduction pipeline; avoids the host sync that
                            baking an immediate would need)                    */
} nl_input_kind;

typedef struct nl_input_desc {
  int32_t  dtype;     /* nl_dtype of the source                                */
  int32_t  kind;      /* nl_input_kind                                         */
  uint64_t imm_bits;  /* NL_IN_SCALAR_IMM: the value, already converted to the
                         plan dtype, as a bit pattern in the low bytes         */
} nl_input_desc;

typedef enum {
  NL_IDX_LOOP    = 0,  /* out[i], i the loop index                              */
  NL_IDX_COMPACT = 1,  /* out[start + k], k-th element that passed the filter   */
  NL_IDX_SHARD   = 2,  /* out[thread_nr] — reduction partial                    */
  NL_IDX_WHERE   = 3   /* out[i] only where the mask holds (in-place assign)    */
} nl_index_space;

typedef struct nl_output_desc {
  int32_t dtype;       /* nl_dtype stored                                       */
  int32_t index_space; /* nl_index_space                                        */
} nl_output_desc;

/* ---- compiled plan: program of the accumulator machine that the pipeline
 *      kernel interprets.  Produced by nl_plan_compile(); opaque to callers
 *      except for the descriptive fields. ------------------------------------ */
typedef struct nl_plan {
  int32_t        abi_version;
  int32_t        dtype;         /* value type T of the register file           */
  int32_t        n_inputs;
  int32_t        n_outputs;
  int32_t        n_steps;
  int32_t        flags;         /* NL_PLAN_* below                              */
  int32_t        reduce_op;     /* nl_op of the root reduction or 0             */
  int32_t        n_temps;       /* value temporaries used                       */
  nl_input_desc  in[NL_MAX_INPUTS];
  nl_output_desc out[NL_MAX_OUTPUTS];
  uint32_t       step[NL_MAX_STEPS];
} nl_plan;

enum {
  NL_PLAN_HAS_FILTER  = 1,   /* contains NL_OP_FILTER                           */
  NL_PLAN_HAS_COMPACT = 2,   /* some output uses NL_IDX_COMPACT (needs the scan)*/
  NL_PLAN_HAS_REDUCE  = 4,
  NL_PLAN_HAS_WHERE   = 8
};

/* ======================= host-only entry points ============================ */

NL_API int         nl_abi_version(void);
NL_API const char *nl_last_error(void);            /* thread-local, never NULL         */
NL_API size_t      nl_dtype_size(int dtype);

/* Lower an operation tree to a plan ("CompilePipeline", compiler.cpp:309).
 * Pure host code; needs no device.  NL_ERR_SPLIT: too many inputs / temps /
 * steps — split the tree at a node and materialise it.  NL_ERR_SPLIT_FILTER: an
 * array load after a filter / a store between two filters — materialise the
 * filter results first.  NL_ERR_UNSUPPORTED: e.g. mixed array dtypes. */
NL_API int nl_plan_compile(const nl_expr_node *nodes, int n_nodes,
                    const nl_input_desc *inputs, int n_inputs,
                    nl_plan *plan_out);

/* Human-readable dump of a plan (debug_printer.cpp's role). Returns bytes
 * written (excluding NUL) or a negative error. */
NL_API int nl_plan_dump(const nl_plan *plan, char *buf, size_t buflen);

/* The reference's range partition (scheduler.cpp:133-147): chunk j of n_chunks
 * over [0,size); `align` (elements, power of two, 0/1 = none) rounds interior
 * boundaries down so 128-bit accesses stay aligned.  align = 1 reproduces the
 * reference exactly. */
NL_API int nl_partition(int64_t size, int n_chunks, int64_t align, int chunk,
                 int64_t *start_out, int64_t *end_out);

/* ======================= runtime ========================================== */

/* Create per-device streams and scratch.  device_ids == NULL: devices
 * 0..n_devices-1.  n_devices == 0: every visible device.  Fails with
 * NL_ERR_NO_DEVICE when there is none. */
NL_API int nl_init(int n_devices, const int *device_ids);
NL_API int nl_shutdown(void);
NL_API int nl_device_count(void);                  /* local devices after nl_init, 0 before */
NL_API int nl_device_info(int dev, char *name, size_t name_len, int *sm_count,
                   int *cc_major, int *cc_minor, size_t *total_mem);

NL_API int nl_device_cuda_id(int dev);              /* CUDA ordinal of local device `dev` (DLPack / __cuda_array_interface__ consumers) */

/* CUDA graphs: everything enqueued on (dev, stream 0) between nl_graph_begin and nl_graph_end — pipeline
 * launches, assignments, fills on FIXED buffers — is captured instead of run and can then be replayed
 * with one call (small pipelines are launch-bound: the reference's per-pipeline JIT cost has no
 * counterpart here, but the per-launch host cost remains).  No allocation, synchronisation, fused
 * multi-GPU finish or sort inside a capture. */
typedef struct nl_graph_s *nl_graph;
NL_API int nl_graph_begin(int dev);
NL_API int nl_graph_end(int dev, nl_graph *graph_out);
NL_API int nl_graph_launch(nl_graph graph);
NL_API int nl_graph_destroy(nl_graph graph);

NL_API int nl_malloc(int dev, size_t bytes, void **ptr_out);
NL_API int nl_free(int dev, void *ptr);
/* Give the free blocks of the device's stream-ordered pool back to the driver (synchronises the
 * device).  After many allocations of different sizes the pool's virtual ranges are backed by small
 * scattered physical chunks; large columns allocated afterwards then stream a few per cent slower. */
NL_API int nl_pool_trim(int dev);
NL_API int nl_memset(int dev, int stream, void *ptr, int value, size_t bytes);
NL_API int nl_host_alloc(size_t bytes, void **ptr_out);   /* pinned */
NL_API int nl_host_free(void *ptr);
NL_API int nl_memcpy_h2d(int dev, int stream, void *dst, const void *src, size_t bytes);
NL_API int nl_memcpy_d2h(int dev, int stream, void *dst, const void *src, size_t bytes);
NL_API int nl_memcpy_d2d(int dev, int stream, void *dst, const void *src, size_t bytes);
/* copy between two local devices (NVLink peer copy), ordered on dst's stream */
NL_API int nl_memcpy_peer(int dst_dev, int stream, void *dst, int src_dev, const void *src, size_t bytes);
NL_API int nl_stream_sync(int dev, int stream);
NL_API int nl_device_sync(int dev);
/* make `waiter` stream wait for everything enqueued so far on `signaler` */
NL_API int nl_stream_wait(int dev, int waiter, int signaler_dev, int signaler);

/* 1 when `ptr` is page-locked host memory the DMA engines can read directly (nl_host_alloc or any
 * registered block), 0 for pageable memory.  Host-only query. */
NL_API int nl_host_is_pinned(const void *ptr);

/* CUDA events: timing on the launching stream, and the dependencies between the three streams of a
 * device that replace the reference's task queue (a task becomes runnable when its children are
 * done, scheduler.cpp:13-50) — e.g. the kernel over chunk i waits for the upload of chunk i only. */
typedef struct nl_event_s *nl_event;
NL_API int nl_event_create(int dev, nl_event *ev_out);
NL_API int nl_event_create_sync(int dev, nl_event *ev_out);   /* ordering only (no timestamps): cheaper */
NL_API int nl_event_query(nl_event ev);                       /* 1 complete, 0 still pending, < 0 error */
/* make (dev, stream) wait for `ev` (recorded on any local device's stream) */
NL_API int nl_stream_wait_event(int dev, int stream, nl_event ev);
NL_API int nl_event_destroy(nl_event ev);
NL_API int nl_event_record(nl_event ev, int stream);
NL_API int nl_event_sync(nl_event ev);
NL_API int nl_event_elapsed_ms(nl_event start, nl_event stop, float *ms_out);

/* ======================= the hot path ===================================== */

/* Run one fused pipeline over [start,end) of its inputs on device `dev`
 * (replaces the call of the JIT'd loop, compiler.cpp:87; argument order and
 * meaning follow the IR prototype, compiler.cpp:340-347):
 *   outputs[k]      device pointer of output k (caller-allocated; NL_IDX_COMPACT
 *                   outputs need room for `end` elements — the reference
 *                   allocates the upper bound too, thunk_as_mapping.cpp:50)
 *   inputs[k]       device pointer of input k (NULL for NL_IN_SCALAR_IMM)
 *   output_sizes    DEVICE pointer to n_outputs int64: final index counter of
 *                   each output — `end`, `start + count`, or `thread_nr`-th
 *                   slot written => 1 (compiler.cpp:491-508)
 *   thread_nr       shard / chunk number: slot of a reduction partial
 * Asynchronous on (dev, stream). */
NL_API int nl_launch_pipeline(const nl_plan *plan, void *const *outputs,
                       const void *const *inputs, int64_t start, int64_t end,
                       int64_t *output_sizes, int thread_nr, int dev, int stream);

/* The same launch for a plan that ends in a reduction, with the multi-GPU
 * finish FUSED into the kernel: the CTA that folds this device's partials
 * writes the result into a mailbox on every peer GPU (stores over NVLink
 * peer / IPC mappings), waits for the peers' slots and folds them in rank
 * order, so out[0] on every rank holds the global result when the kernel
 * ends — no collective launch follows (replaces llvm_finalize_sum_data,
 * thunk_methods.cpp:133-151, across devices).  Every rank must issue the same
 * sequence of these calls (like a collective); stream 0, slot 0.  Requires
 * nl_comm_peer_ready(). */
NL_API int nl_launch_pipeline_allreduce(const nl_plan *plan, void *const *outputs,
                                 const void *const *inputs, int64_t start, int64_t end,
                                 int64_t *output_sizes, int dev);

/* The filter's twin: a compaction pipeline whose kernel also posts the shard's
 * count to every peer and leaves in `offsets` (DEVICE, world_size + 1 int64)
 * the exclusive scan of all ranks' counts — where each shard's run starts in
 * the concatenated result, and the total (the memmove gather of
 * compiler.cpp:36-48 without moving data).  Same calling rules as above. */
NL_API int nl_launch_pipeline_allgather(const nl_plan *plan, void *const *outputs,
                                 const void *const *inputs, int64_t start, int64_t end,
                                 int64_t *output_sizes, int64_t *offsets, int dev);

/* Combine `n` reduction partials in order j = 0..n-1 into out[0]
 * (llvm_finalize_sum_data, thunk_methods.cpp:133-151; same for max/min).
 * Device-side, asynchronous. */
NL_API int nl_finalize_partials(int dev, int stream, int reduce_op, int dtype,
                         const void *partials, int n, void *out);

/* Multi-device finish of a reduction: NCCL allreduce (max/min/sum) of ONE
 * element of `dtype` at scalar[d] on every local device d, in place, across
 * the whole communicator (all ranks).  Requires nl_comm_init(). */
NL_API int nl_finish_reduce(int reduce_op, int dtype, void *const *scalar /*[n_local]*/, int stream);

/* Multi-device finish of a filter: allgather the per-shard counts and take the
 * exclusive scan (the memmove gather of compiler.cpp:36-48 turned into
 * offsets).  count[d] is a DEVICE pointer to the shard's int64 count on local
 * device d; offsets[d] is a DEVICE pointer to world_size+1 int64 on device d
 * receiving the exclusive scan of all shard counts (offsets[world] = total). */
NL_API int nl_finish_filter(const int64_t *const *count, int64_t *const *offsets, int stream);

/* In-place assignment on an evaluated thunk's storage
 * (thunk_as_mapping.cpp:80-86): dst[lo + k*step] = value for k in [0,count). */
NL_API int nl_assign_fill(int dev, int stream, void *dst, int dtype,
                   int64_t lo, int64_t step, int64_t count, uint64_t value_bits);
/* dst[lo + k*step] = src[k] for k in [0,count); src on the same device. */
NL_API int nl_assign_copy(int dev, int stream, void *dst, int dtype,
                   int64_t lo, int64_t step, int64_t count, const void *src);

/* Synthetic column generator for benchmarks and tests, bit-identical to the
 * oracle's nlo_fill: x[i] = f(splitmix64(seed ^ (first_index + i))).
 * kind 0: ints uniform in [1,100); 1: full-range ints; 2: floats in [1,2);
 * 3: floats in [0,1).   */
NL_API int nl_fill(int dev, int stream, void *dst, int dtype, int64_t first_index,
            int64_t count, uint64_t seed, int kind);

/* dst[i] = (dst_dtype)src[i], i in [0, count): the materialise-store cast of compiler.cpp:256-260 as a pass of
 * its own, used where two columns of different types meet (NumPy promotion; the reference gives the result
 * its LEFT operand's type and leaves "todo: correct type", thunk_as_number.cpp:78).  Any value type to any
 * value type, bool as a source only; C conversion rules (integer narrowing wraps, float -> int truncates). */
NL_API int nl_cast(int dev, int stream, void *dst, int dst_dtype, const void *src, int src_dtype, int64_t count);

/* Three 64-bit checksums of `nbytes` of device memory read as little-endian 64-bit words w_i (a
 * trailing partial word zero-padded; `ptr` 8-byte aligned):
 *   out[0] = sum w_i,  out[1] = xor w_i,  out[2] = sum w_i * (2 i + 1)      (mod 2^64)
 * `out`: DEVICE pointer to 3 uint64.  The device half of the full-size parity check: the host forms
 * the same sums chunk by chunk from the oracle's output (bench.py).  Asynchronous on (dev, stream)
 * unless nbytes is not a multiple of 8. */
NL_API int nl_checksum(int dev, int stream, const void *ptr, size_t nbytes, uint64_t *out);

/* Integer-array indexing (declared but never implemented in the reference: `a[idx]` with an int64
 * index has a cardinality function but no emitter, thunk_as_mapping.cpp:51-55):
 *   out[i] = column[idx[i]],  i in [0, count)
 * `column` is passed as its shards — device pointers that may live on other local GPUs (read with
 * peer loads over NVLink; requires nl_comm_init) — with the global index of each shard's first
 * element.  Negative indices wrap as in NumPy; *bad (DEVICE int64, caller-zeroed) counts the
 * indices outside [-total, total). */
NL_API int nl_gather(int dev, int stream, int dtype, void *out, const int64_t *idx, int64_t count,
              const void *const *shard_ptr, const int64_t *shard_start, int n_shards, int64_t total,
              int64_t *bad);

/* Sort `count` keys of `dtype` on the device, ascending, NaNs last (NumPy order) — the `base`
 * function of the sort() full breaker (PyArray_Sort on a copy, thunk_methods.cpp:37-62; run by
 * compiler.cpp:93-122).  In place in `keys`; `tmp` is scratch of the same size.  LSD radix
 * sort; digit positions on which all keys agree cost no pass.  Synchronises the stream once. */
NL_API int nl_sort(int dev, int stream, int dtype, void *keys, void *tmp, int64_t count);

/* The two device steps of a sort across GPUs (sample sort; the reference sorts the whole array with one host call,
 * thunk_methods.cpp:37-62).  Keys are compared through nl_sort's order-preserving bijection of their bit patterns
 * ("order keys", widened to 64 bits: signed integers with the sign bit flipped, floats with NaNs last).
 *   nl_sort_sample  order keys of `n_samples` regularly spaced elements of `keys` -> host_keys.
 *   nl_sort_split   cuts `keys` at `n_splitters` (<= 15) ascending order keys: `out` receives the keys grouped, stably,
 *                   into n_splitters + 1 runs — run j holds the keys in (splitter[j-1], splitter[j]] — and
 *                   host_counts[j] the length of run j (one counting pass + one sweep of nl_sort's kernel with the run
 *                   number as its digit).
 * Both are stream-ordered: the host arrays are valid after the stream has been synchronised, and with page-locked
 * host arrays (nl_host_alloc) the calls return at once, so several devices work side by side. */
NL_API int nl_sort_sample(int dev, int stream, int dtype, const void *keys, int64_t count, int n_samples, uint64_t *host_keys);
NL_API int nl_sort_split(int dev, int stream, int dtype, const void *keys, void *out, int64_t count, const uint64_t *splitters,
                         int n_splitters, int64_t *host_counts);
/* The same in two steps, so that the runs can go straight to their destinations — buffers on OTHER GPUs included: the
 * sweep stores run j, stably, to run_dst[j] (device addresses: local memory or peer memory mapped into this device,
 * nl_comm_init), i.e. the exchange of a sort across GPUs happens inside the kernel, over NVLink, while it regroups.
 * run_dst[j] may be null for an empty run. */
NL_API int nl_sort_split_count(int dev, int stream, int dtype, const void *keys, int64_t count, const uint64_t *splitters,
                               int n_splitters, int64_t *host_counts);
NL_API int nl_sort_split_scatter(int dev, int stream, int dtype, const void *keys, int64_t count, const uint64_t *splitters,
                                 int n_splitters, void *const *run_dst);

/* ======================= communicator (NCCL over NVLink) ================== */

#define NL_UNIQUE_ID_BYTES 128
NL_API int nl_comm_unique_id(void *id_out /*NL_UNIQUE_ID_BYTES*/);
/* Local device d becomes rank first_rank + d of a world_size communicator.
 * Single process driving all GPUs: id from nl_comm_unique_id(), first_rank 0,
 * world_size = nl_device_count().  One process per GPU: rank 0 creates the id,
 * the host plumbing broadcasts it. */
NL_API int nl_comm_init(const void *id, int world_size, int first_rank);
NL_API int nl_comm_world_size(void);               /* 0 when no communicator            */
NL_API int nl_comm_destroy(void);

/* NVLink mailboxes: the peer memory behind nl_launch_pipeline_allreduce().
 * A process that drives every rank itself gets them from nl_comm_init() (CUDA
 * peer access).  With one process per GPU each rank exports the IPC handle of
 * its mailbox, the host plumbing all-gathers the handles (rank order,
 * NL_MAILBOX_HANDLE_BYTES each) and every rank imports the lot. */
#define NL_MAILBOX_HANDLE_BYTES 64
NL_API int nl_comm_mailbox_export(int dev, void *handle_out /*NL_MAILBOX_HANDLE_BYTES*/);
NL_API int nl_comm_mailbox_import(const void *handles /*world_size x NL_MAILBOX_HANDLE_BYTES*/);
NL_API int nl_comm_peer_ready(void);               /* 1 when the fused finish can be used */
/* The rank on local device `dev` cannot take part in the next fused finish (its own error path): post an
 * abort for that epoch into every peer's mailbox, so that kernels already waiting for this rank stop at
 * once instead of running into the peer timeout; their next synchronisation returns NL_ERR_NCCL.  The
 * library does the same by itself when a fused launch fails after its epoch was committed.  The
 * communicator is unusable afterwards (nl_comm_destroy + nl_comm_init). */
NL_API int nl_comm_abort(int dev);

/* ======================= introspection ==================================== */
/* Kernels launched by this library since load (the bench's gpu_launches). */
NL_API uint64_t nl_launch_count(void);
/* Kernel-template selection (the plan optimiser's last stage): name of the kernel a plan
 * would run — a build-time instantiated shape ("static:<shape>") or the dynamic
 * interpreter.  Host-only.  vec16: rows of 128 bits (aligned operands) or scalar rows. */
NL_API int nl_plan_kernel(const nl_plan *plan, int vec16, char *name, size_t name_len, int *is_static);
/* Which tiers of kernel-template selection a launch may use (tests run every tier on the same
 * programs): 1 (default) exact build-time shapes, then shape families, then the step interpreter;
 * 2 families and interpreter; 0 interpreter only. */
NL_API int nl_set_static_kernels(int mode);
/* Named tuning knobs (benchmarking / tests; the defaults are what the measurements in DESIGN.md use):
 *   "compact_kernel"     0 chosen by size: super-tile kernel, staged TMA pipeline, plain kernel (default);
 *                        1 plain kernel only; 2 staged pipeline before the super-tile kernel
 *   "pipe_dynamic"       1: the staged pipeline may also run under the step interpreter (default 0)
 *   "pipe_side_buffers"  0: no side buffers in the staged pipeline (default 1)
 *   "pipe_lag"           tiles between count and store phase of the staged pipeline (0: half the ring)
 *   "max_ctas_per_sm"    cap on resident CTAs per SM of the streaming kernels (0: by occupancy)
 *   "sort_algo"          radix sort of 4- and 8-byte keys: 0 one sweep per digit position (default: decoupled look-back over
 *                        tile histograms, 16 B per 8-byte key and position); 1 count + scatter per position (24 B; always
 *                        used for 1- and 2-byte keys)
 *   "compact_self_clean" 0: clear the compaction arena (ticket, descriptors) with a memset in front of every launch; default 1:
 *                        two arenas alternate and every super-tile launch zeroes the one its predecessor used
 *   "sort_portion_log2"  keys per launch of the one-sweep form (13..28; 0: default 28) — the tests shrink it to chain launches
 *   "poll_sleep_ns"      back-off between two polls of a look-back descriptor in the super-tile kernel (default 0)
 *   "peer_timeout_ms"    how long a fused multi-GPU finish waits for a peer's mailbox slot before it gives up
 *                        and the next synchronisation returns NL_ERR_NCCL (default 10000; 0: default)
 *   "reset"              all of the above back to their defaults (value ignored)
 * Unknown keys and out-of-range values return NL_ERR_INVALID. */
NL_API int nl_tune(const char *key, int value);
/* Name and grid of the most recent pipeline launch (for logs / tests). */
NL_API int nl_last_launch_info(char *kernel_name, size_t name_len, int *grid, int *block, int *vec_elems);

#ifdef __cplusplus
}
#endif
#endif /* NL_ABI_H */
