// nl_kernels.cu — host-side kernel dispatch of libnlkernels plus the small kernels
// (finalize, assignment fill/copy, count scan, synthetic fill).  The fused pipeline
// kernel template lives in nl_pipeline.cuh; this file instantiates its dynamic
// (interpreter) variants and selects between them and the build-time shapes of
// nl_static.cuh.
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <limits.h>
#include <stdio.h>
#include <string.h>

#include "nl_compact_pipe.cuh"
#include "nl_compact_super.cuh"

namespace nl {

// A launch over an empty range still has to define its outputs: cardinalities and the
// reduction identity (llvm_initialize_max / _sum with zero iterations).
template <typename T>
__global__ void empty_range_kernel(const __grid_constant__ KParams p) {
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  for (int o = 0; o < p.n_outputs; o++) {
    if (p.out_space[o] == NL_IDX_SHARD) {
      using Acc = typename TT<T>::Acc;
      // identity of the reduction; with a fused multi-GPU finish the empty shard still takes part
      unsigned long long bits = (p.reduce_rop == R_SUM) ? to_bits<Acc>(Acc(0))
                                                        : to_bits<T>(p.reduce_rop == R_MAX ? TT<T>::lowest() : TT<T>::highest());
      if (p.peer_box != nullptr) bits = peer_allreduce<T>(p, bits, p.reduce_rop, lane);
      if (lane == 0) {
        p.out_sizes[o] = 1;
        T *out = reinterpret_cast<T *>(p.out[o]);
        out[p.thread_nr] = (p.reduce_rop == R_SUM) ? (T)from_bits<Acc>(bits) : from_bits<T>(bits);
      }
    } else if (lane == 0) {
      p.out_sizes[o] = (p.out_space[o] == NL_IDX_COMPACT) ? p.start : p.end;
    }
  }
  if (p.peer_offsets != nullptr) peer_scan_counts(p, 0ull, lane);   // an empty shard keeps nothing but still posts its slot
}

// ------------------------------------------------------------------ small kernels
template <typename T>
__global__ void finalize_partials_kernel(int rop, const T *__restrict__ partials, int n, T *__restrict__ out) {
  // llvm_finalize_sum_data (thunk_methods.cpp:133-151): fold the partials in order j = 0..n-1
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  T d = (rop == R_SUM) ? T(0) : (rop == R_MAX ? TT<T>::lowest() : TT<T>::highest());
  for (int j = 0; j < n; j++) {
    if (rop == R_SUM) d = op_add<T>(d, partials[j]);
    else if (rop == R_MAX) d = op_max<T>(d, partials[j]);
    else d = op_min<T>(d, partials[j]);
  }
  out[0] = d;
}

template <typename T>
__global__ void __launch_bounds__(256) assign_fill_kernel(T *__restrict__ dst, int64_t lo, int64_t step, int64_t count, T value) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) dst[lo + k * step] = value;
}
// contiguous fill: 128-bit stores on the aligned body
template <typename T>
__global__ void __launch_bounds__(256) assign_fill_vec_kernel(T *__restrict__ dst, int64_t count, T value) {
  constexpr int V = 16 / sizeof(T);
  const uintptr_t addr = reinterpret_cast<uintptr_t>(dst);
  int64_t head = (int64_t)(((16 - (addr & 15)) & 15) / sizeof(T));
  if (head > count) head = count;
  const int64_t nvec = (count - head) / V;
  const int64_t tail0 = head + nvec * V;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  union { uint4 q; T t[V]; } u;
#pragma unroll
  for (int v = 0; v < V; v++) u.t[v] = value;
  uint4 *body = reinterpret_cast<uint4 *>(dst + head);
  for (int64_t k = gtid; k < nvec; k += stride) __stcs(body + k, u.q);
  if (gtid < head) dst[gtid] = value;
  if (gtid < count - tail0) dst[tail0 + gtid] = value;
}
template <typename T>
__global__ void __launch_bounds__(256) assign_copy_kernel(T *__restrict__ dst, int64_t lo, int64_t step, int64_t count, const T *__restrict__ src) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) dst[lo + k * step] = __ldcs(src + k);
}

__global__ void scan_offsets_kernel(const int64_t *__restrict__ counts, int n, int64_t *__restrict__ offsets) {
  // the memmove gather of compiler.cpp:36-48, reduced to where each shard's run starts
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int64_t run = 0;
  for (int j = 0; j < n; j++) { offsets[j] = run; run += counts[j]; }
  offsets[n] = run;
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
// synthetic inputs, bit-identical to the oracle's nlo_fill
template <typename T, bool kBool = false>
__global__ void __launch_bounds__(256) fill_kernel(T *__restrict__ dst, int64_t first_index, int64_t count, uint64_t seed, int kind) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
    const uint64_t z = splitmix64(seed ^ (uint64_t)(first_index + k));
    T v;
    if constexpr (kBool) {
      v = (T)(z & 1);
    } else if constexpr (!TT<T>::is_float) {
      v = (kind == 0) ? (T)(1 + (((z >> 32) * 99ULL) >> 32)) : (T)(long long)z;
    } else if constexpr (sizeof(T) == 4) {
      const float f = __uint_as_float(0x3f800000u | (uint32_t)(z >> 41));
      v = (kind == 3) ? f - 1.0f : f;
    } else {
      const double d = __longlong_as_double((long long)(0x3ff0000000000000ULL | (z >> 12)));
      v = (kind == 3) ? d - 1.0 : d;
    }
    dst[k] = v;
  }
}

// ------------------------------------------------------------------ host-side dispatch
// The interpreter-driven kernels are instantiated per dtype in nl_dynamic_<dtype>.cu (one translation
// unit each, so the build spreads over the cores); plans that never save a value temporary take the
// light variant (fewer registers, one more resident CTA per SM: measured +10-17 %).
static pipeline_fn pick_dynamic(int dtype, bool vec16, bool compact, bool reduce, bool uses_temps) {
  switch (dtype) {
#define NL_CASE(dt, T, sfx) case dt: return dynamic_kernel_##sfx(vec16, compact, reduce, uses_temps);
    NL_FOR_VALUE_DTYPES(NL_CASE)
#undef NL_CASE
  }
  return nullptr;
}

// kernel-template selection: a build-time instantiation when the plan's step sequence is in
// the library (nl_static.cuh), the interpreter otherwise
static int g_static_enabled = 1;
void set_static_kernels_enabled(int on) { g_static_enabled = on; }

// three tiers: an exact build-time shape (nl_static.cuh), a shape family (nl_family.cuh: operators and
// inputs are run-time fields), else the step interpreter.  g_static_enabled: 1 all tiers, 2 families and
// interpreter only, 0 interpreter only (tests run every tier on the same programs).
static const StaticEntry *find_static(const nl_plan *plan, bool vec16) {
  if (!g_static_enabled || !vec16) return nullptr;
  for (int tier = (g_static_enabled == 2 ? 2 : 0); tier < 4; tier++) {          // each list is compiled in two halves per dtype
    int n = 0;
    const StaticEntry *tab = nullptr;
    const int half = tier & 1;
    if (tier < 2) {
      switch (plan->dtype) {
        case NL_I32: tab = half ? static_table_i32_b(&n) : static_table_i32(&n); break;
        case NL_I64: tab = half ? static_table_i64_b(&n) : static_table_i64(&n); break;
        case NL_F32: tab = half ? static_table_f32_b(&n) : static_table_f32(&n); break;
        case NL_F64: tab = half ? static_table_f64_b(&n) : static_table_f64(&n); break;
        default: return nullptr;
      }
    } else {
      switch (plan->dtype) {
        case NL_I32: tab = half ? family_table_i32_b(&n) : family_table_i32(&n); break;
        case NL_I64: tab = half ? family_table_i64_b(&n) : family_table_i64(&n); break;
        case NL_F32: tab = half ? family_table_f32_b(&n) : family_table_f32(&n); break;
        case NL_F64: tab = half ? family_table_f64_b(&n) : family_table_f64(&n); break;
        default: return nullptr;
      }
    }
    for (int i = 0; i < n; i++)
      if (tab[i].matches(plan)) return &tab[i];
  }
  return nullptr;
}

static pipeline_fn pick(const nl_plan *plan, bool vec16, const char **static_name) {
  const bool compact = (plan->flags & NL_PLAN_HAS_COMPACT) != 0;
  const bool reduce = (plan->flags & NL_PLAN_HAS_REDUCE) != 0;
  const StaticEntry *e = find_static(plan, vec16);
  if (static_name) *static_name = e ? e->name : nullptr;
  return e ? e->fn : pick_dynamic(plan->dtype, vec16, compact, reduce, plan->n_temps > 0);
}

static const char *dt_name(int d) {
  static const char *n[] = {"bool", "i32", "i64", "f32", "f64", "i8", "i16", "u8", "u16", "u32", "u64"};
  return (d >= 0 && d < NL_N_DTYPES) ? n[d] : "?";
}

#define CUDA_TRY(expr)                                                             \
  do {                                                                             \
    cudaError_t e_ = (expr);                                                       \
    if (e_ != cudaSuccess) {                                                       \
      set_error("%s failed: %s", #expr, cudaGetErrorString(e_));                   \
      return NL_ERR_CUDA;                                                          \
    }                                                                              \
  } while (0)

// the interpreter's staged compaction pipeline, one per dtype
static pipeline_fn pick_dynamic_pipe(int dtype, int rows) {
  switch (dtype) {
#define NL_CASE(dt, T, sfx) case dt: return dynamic_pipe_##sfx(rows);
    NL_FOR_VALUE_DTYPES(NL_CASE)
#undef NL_CASE
  }
  return nullptr;
}

static int g_pipe_enabled = 1;
static int g_pipe_dynamic = 0;     // allow the interpreter-driven staged pipeline (tests)
void set_compact_pipe_dynamic(int on) { g_pipe_dynamic = on; }
static int g_pipe_no_side = 0;
static int g_max_ctas_per_sm = 0;   // tuning: cap on resident CTAs per SM of the streaming kernels (0: occupancy)
void set_max_ctas_per_sm(int n) { g_max_ctas_per_sm = n; }
static int g_pipe_lag = 0;       // 0: slots - 2
void set_compact_pipe_enabled(int on) { g_pipe_enabled = on; }
void set_compact_pipe_side(int on) { g_pipe_no_side = !on; }
void set_compact_pipe_lag(int l) { g_pipe_lag = l; }

// Can this plan run on the staged compaction pipeline?  One filter, nothing stored before it,
// at least one value column to stage, enough tiles to fill the machine.  Fills the staged fields.
// dynamic shared memory of compact_pipe_kernel (must mirror the carve-up at the top of the kernel)
static size_t pipe_smem_bytes(int slots, int lag, int cols, bool side, size_t es, int rows) {
  const size_t col_bytes = (size_t)kPipeConsumers * rows * 16;
  const size_t entries = (size_t)lag + 1;
  return ((sizeof(PipeShared) + 127) / 128) * 128 + entries * kPipeConsumers * (((size_t)rows * 16 / es <= 16) ? 2 : 4) +
         (side ? entries * kPipeConsumerWarps * pipe_side_bytes(rows) : 0) + (size_t)slots * cols * col_bytes;
}

// A plan can be split at its filter into "count" and "store" halves (staged pipeline, super-tile
// kernel) when it has exactly one filter, stores nothing before it and reads no predicate
// temporary after it (the store half skips the predicate steps).  Returns the filter's step or -1.
static int split_point(const nl_plan *plan) {
  if (!(plan->flags & NL_PLAN_HAS_COMPACT) || (plan->flags & (NL_PLAN_HAS_REDUCE | NL_PLAN_HAS_WHERE))) return -1;
  int filter = -1;
  for (int i = 0; i < plan->n_steps; i++) {
    const uint32_t op = step_op(plan->step[i]);
    if (op == S_FILTER) {
      if (filter >= 0) return -1;
      filter = i;
    } else if ((op == S_STORE || op == S_PSTORE || op == S_WHERE) && filter < 0) {
      return -1;
    } else if ((op == S_PLOADT || op == S_PBIN || op == S_PSTORE) && filter >= 0) {
      return -1;
    }
  }
  return filter;
}

static int g_poll_sleep_ns = 0;     // tuning: back-off in the look-back poll of the super-tile kernel
void set_poll_sleep_ns(int v) { g_poll_sleep_ns = v < 0 ? 0 : v; }
static int g_prefer_pipe = 0;      // tuning: staged pipeline before the super-tile kernel
void set_compact_prefer_pipe(int on) { g_prefer_pipe = on; }

// the interpreter's super-tile kernel
static pipeline_fn pick_dynamic_super(int dtype, bool uses_temps) {
  switch (dtype) {
#define NL_CASE(dt, T, sfx) case dt: return dynamic_super_##sfx(uses_temps);
    NL_FOR_VALUE_DTYPES(NL_CASE)
#undef NL_CASE
  }
  return nullptr;
}

// super-tile compaction kernel for this plan?  Fills kp->super_k / super_rows / filter_step.
static bool super_eligible(const nl_plan *plan, bool vec16, int64_t n_elems, KParams *kp, int sm_count) {
  if (!g_pipe_enabled || !vec16) return false;        // "plain kernel only" switches both off
  const int filter = split_point(plan);
  if (filter < 0) return false;
  const StaticEntry *se = find_static(plan, true);
  const int rows = (se && se->super) ? se->super_rows : kRows;
  const int64_t tile = (int64_t)kBlock * rows * (int64_t)nl_vec_elems(plan->dtype);
  if (n_elems < tile * kSuperK * sm_count) return false;      // at least one super-tile per SM
  if (kp) {
    kp->super_k = kSuperK;
    kp->super_rows = rows;
    kp->filter_step = filter;
    kp->n_inputs = plan->n_inputs;
    kp->store_prefix = store_pass_needs_prefix(plan->step, plan->n_steps, filter) ? 1 : 0;
    kp->poll_sleep_ns = (uint32_t)g_poll_sleep_ns;
  }
  return true;
}

static bool pipe_eligible(const nl_plan *plan, bool vec16, int64_t n_elems, KParams *kp, int sm_count) {
  if (!g_pipe_enabled || !vec16) return false;
  const int filter = split_point(plan);
  if (filter < 0) return false;
  const size_t es = nl_dtype_size(plan->dtype);
  int cols = 0;
  uint8_t map[NL_MAX_INPUTS];
  for (int k = 0; k < NL_MAX_INPUTS; k++) map[k] = 0xff;
  for (int k = 0; k < plan->n_inputs; k++)
    if (plan->in[k].kind == NL_IN_ARRAY && plan->in[k].dtype == plan->dtype) map[k] = (uint8_t)cols++;
  if (cols == 0) return false;
  // one staged column: 32 KB slots (8 rows per thread); two or three: the 4-row instantiation with
  // 16 KB per column and slot, so the ring keeps at least four slots.  Build-time shapes exist for
  // the 8-row geometry only.
  const StaticEntry *se = find_static(plan, true);
  const bool has_static_pipe = se && se->pipe;
  // The interpreter-driven pipeline is kept for testing only: with eight consumer warps per SM the
  // step interpreter is issue-bound (measured 0.8-1.0 TB/s vs 1.2-1.9 for the plain interpreter
  // kernel and 5-6 TB/s for the build-time shapes), so plans without a build-time shape stay on the
  // plain kernel unless the tuning flag asks for it.
  if (!has_static_pipe && !g_pipe_dynamic) return false;
  if (se && !se->pipe) return false;          // a build-time shape whose plain kernel is the fast one
  const int rows = (cols == 1) ? kPipeRows : kPipeRowsNarrow;
  if (has_static_pipe && se->pipe_rows != rows) return false;   // e.g. a scalar slot of the shape bound to a column
  if (es < 4) return false;                    // the TMA ring is laid out in 16-byte rows of 4- and 8-byte elements
  const int64_t tile = (int64_t)kPipeConsumers * rows * (int64_t)(16 / es);
  if (n_elems < tile * 2 * sm_count) return false;
  // the side buffer can take a survivor only if everything after the filter works on a bare
  // value: scalar operands and compacted stores
  bool sparse_ok = true;
  for (int i = filter + 1; i < plan->n_steps; i++) {
    const uint32_t st = plan->step[i], op = step_op(st);
    if (op == S_BIN) {
      const uint32_t src = step_b(st);
      if ((src & SRC_TEMP) || plan->in[src].kind == NL_IN_ARRAY) sparse_ok = false;
    } else if (op == S_STORE) {
      if (step_b(st) == NL_IDX_LOOP) sparse_ok = false;
    } else {
      sparse_ok = false;
    }
  }
  if (g_pipe_no_side) sparse_ok = false;   // tuning switch
  // deepest ring that fits: [PipeShared | keep bits and side buffers per lag entry | slots]
  int slots = 0, lag = 0;
  for (int s_try = kPipeMaxSlots; s_try >= 4; s_try--) {
    int d = (g_pipe_lag > 0) ? g_pipe_lag : s_try / 2;   // half the ring hides the look-back, half is prefetch (measured best)
    if (d > s_try - 2) d = s_try - 2;                    // a parked tile must be stored before the producer needs its slot
    if (s_try + d + 1 > kPipeMeta) continue;
    if (pipe_smem_bytes(s_try, d, cols, sparse_ok, es, rows) <= (size_t)227 * 1024 - 1024) { slots = s_try; lag = d; break; }
  }
  if (slots == 0) return false;
  if (kp) {
    for (int k = 0; k < NL_MAX_INPUTS; k++) kp->stage_col[k] = map[k];
    kp->n_stage_cols = cols;
    kp->n_slots = slots;
    kp->lag = lag;
    kp->filter_step = filter;
    kp->n_inputs = plan->n_inputs;
    kp->pipe_rows = rows;
    kp->side_cap = sparse_ok ? (int32_t)(pipe_side_bytes(rows) / es) : 0;
  }
  return true;
}

int64_t pipeline_tile_elems(const nl_plan *plan, bool vec16, int64_t n_elems, KParams *kp, int sm_count) {
  if (sm_count < 1) sm_count = 148;        // host-only callers (plan descriptions): a B200
  const int v = vec16 ? nl_vec_elems(plan->dtype) : 1;
  KParams local;
  KParams *q = kp ? kp : &local;
  q->super_k = 0;
  q->n_slots = 0;
  if (g_prefer_pipe && pipe_eligible(plan, vec16, n_elems, q, sm_count)) return (int64_t)kPipeConsumers * q->pipe_rows * v;
  if (super_eligible(plan, vec16, n_elems, q, sm_count)) return (int64_t)kBlock * q->super_rows * v * q->super_k;
  if (!g_prefer_pipe && pipe_eligible(plan, vec16, n_elems, q, sm_count)) return (int64_t)kPipeConsumers * q->pipe_rows * v;
  const StaticEntry *e = find_static(plan, vec16);
  const int rows = e ? e->rows : kRows;
  return (int64_t)kBlock * rows * v;
}

int describe_pipeline_kernel(const nl_plan *plan, bool vec16, char *name, size_t name_len, int *is_static) {
  const char *sname = nullptr;
  pipeline_fn fn = pick(plan, vec16, &sname);
  if (!fn) { set_error("no pipeline kernel for dtype %d", plan->dtype); return NL_ERR_UNSUPPORTED; }
  const bool compact = (plan->flags & NL_PLAN_HAS_COMPACT) != 0;
  const bool reduce = (plan->flags & NL_PLAN_HAS_REDUCE) != 0;
  const int v = vec16 ? nl_vec_elems(plan->dtype) : 1;
  if (name && name_len) {
    if (sname) snprintf(name, name_len, "pipeline_kernel<%s,V=%d,%s%s>", dt_name(plan->dtype), v, strncmp(sname, "family:", 7) ? "static:" : "", sname);
    else snprintf(name, name_len, "pipeline_kernel<%s,V=%d,compact=%d,reduce=%d,dynamic%s>", dt_name(plan->dtype), v, (int)compact, (int)reduce,
                  plan->n_temps > 0 ? "" : "-light");
  }
  if (is_static) *is_static = sname != nullptr;
  return NL_OK;
}

static int launch_compact_pipe(const nl_plan *plan, const KParams &kp, int sm_count, void *stream, LaunchInfo *info) {
  const StaticEntry *e = find_static(plan, true);
  pipeline_fn fn = (e && e->pipe && kp.pipe_rows == e->pipe_rows) ? e->pipe : pick_dynamic_pipe(plan->dtype, kp.pipe_rows);
  if (!fn) { set_error("no compaction pipeline for dtype %d", plan->dtype); return NL_ERR_UNSUPPORTED; }
  const size_t smem = pipe_smem_bytes(kp.n_slots, kp.lag, kp.n_stage_cols, kp.side_cap != 0, nl_dtype_size(plan->dtype), kp.pipe_rows);
  CUDA_TRY(cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t grid = sm_count;
  if (grid > kp.n_tiles) grid = kp.n_tiles;
  fn<<<(unsigned)grid, kPipeThreads, smem, (cudaStream_t)stream>>>(kp);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  if (info) {
    const int v = nl_vec_elems(plan->dtype);
    snprintf(info->kernel, sizeof(info->kernel), "compact_pipe_kernel<%s,V=%d,U=%d,%s%s,slots=%d,lag=%d,side=%d>", dt_name(plan->dtype), v, kp.pipe_rows,
             (e && e->pipe) ? "static:" : "dynamic", (e && e->pipe) ? e->name : "", kp.n_slots, kp.lag, kp.side_cap);
    info->grid = (int)grid;
    info->block = kPipeThreads;
    info->vec_elems = v;
  }
  return NL_OK;
}

static int launch_compact_super(const nl_plan *plan, const KParams &kp, int sm_count, void *stream, LaunchInfo *info) {
  const StaticEntry *e = find_static(plan, true);
  const bool is_static = e && e->super && e->super_rows == kp.super_rows;
  pipeline_fn fn = is_static ? e->super : pick_dynamic_super(plan->dtype, plan->n_temps > 0);
  if (!fn) { set_error("no super-tile compaction kernel for dtype %d", plan->dtype); return NL_ERR_UNSUPPORTED; }
  int bps = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, fn, kBlock, 0));
  if (bps < 1) bps = 1;
  if (g_max_ctas_per_sm > 0 && bps > g_max_ctas_per_sm) bps = g_max_ctas_per_sm;
  int64_t grid = (int64_t)sm_count * bps;
  if (grid > kp.n_tiles) grid = kp.n_tiles;
  if (grid < 1) grid = 1;
  KParams kq = kp;
  // on dense input every SM keeps three resident CTAs: whichever started fourth on its SM (counted per %smid by
  // the kernel itself — no assumption about how the hardware places CTAs) retires
  kq.super_keep_ctas = (bps >= 4 && grid == (int64_t)sm_count * bps) ? 3 : 0x7fffffff;
  fn<<<(unsigned)grid, kBlock, 0, (cudaStream_t)stream>>>(kq);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  if (info) {
    const int v = nl_vec_elems(plan->dtype);
    snprintf(info->kernel, sizeof(info->kernel), "compact_super_kernel<%s,V=%d,U=%d,K=%d,%s%s>", dt_name(plan->dtype), v, kp.super_rows, kp.super_k,
             is_static ? (strncmp(e->name, "family:", 7) ? "static:" : "") : (plan->n_temps > 0 ? "dynamic" : "dynamic-light"), is_static ? e->name : "");
    info->grid = (int)grid;
    info->block = kBlock;
    info->vec_elems = v;
  }
  return NL_OK;
}

int launch_pipeline_kernel(const nl_plan *plan, const KParams &kp, bool vec16, int sm_count, void *stream, LaunchInfo *info) {
  if (kp.super_k > 0) return launch_compact_super(plan, kp, sm_count, stream, info);
  if (kp.n_slots > 0) return launch_compact_pipe(plan, kp, sm_count, stream, info);
  pipeline_fn fn = pick(plan, vec16, nullptr);
  if (!fn) { set_error("no pipeline kernel for dtype %d", plan->dtype); return NL_ERR_UNSUPPORTED; }
  int bps = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, fn, kBlock, 0));
  if (bps < 1) bps = 1;
  // streaming kernels that also write prefer three resident CTAs per SM (measured on B200, c = a*b at
  // 2^30 f32: 2 -> 6.25, 3 -> 6.40, 4 -> 6.31 TB/s); pure readers (reductions) take all they can get
  if (!(plan->flags & (NL_PLAN_HAS_REDUCE | NL_PLAN_HAS_COMPACT)) && bps > 3) bps = 3;
  if (g_max_ctas_per_sm > 0 && bps > g_max_ctas_per_sm) bps = g_max_ctas_per_sm;
  int64_t grid = (int64_t)sm_count * bps;
  if (grid > kp.n_tiles) grid = kp.n_tiles;
  if (grid < 1) grid = 1;
  fn<<<(unsigned)grid, kBlock, 0, (cudaStream_t)stream>>>(kp);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  if (info) {
    describe_pipeline_kernel(plan, vec16, info->kernel, sizeof(info->kernel), nullptr);
    info->grid = (int)grid;
    info->block = kBlock;
    info->vec_elems = vec16 ? nl_vec_elems(plan->dtype) : 1;
  }
  return NL_OK;
}

int launch_empty_range(void *stream, const nl_plan *plan, const KParams &kp) {
  switch (plan->dtype) {
#define NL_CASE(dt, T, sfx) case dt: empty_range_kernel<T><<<1, 32, 0, (cudaStream_t)stream>>>(kp); break;
    NL_FOR_VALUE_DTYPES(NL_CASE)
#undef NL_CASE
    default: set_error("bad dtype"); return NL_ERR_INVALID;
  }
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return NL_OK;
}

int launch_finalize_partials(void *stream, int rop, int dtype, const void *partials, int n, void *out) {
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
#define NL_CASE(dt, T, sfx) case dt: finalize_partials_kernel<T><<<1, 32, 0, st>>>(rop, (const T *)partials, n, (T *)out); break;
    NL_FOR_VALUE_DTYPES(NL_CASE)
#undef NL_CASE
    default: set_error("bad dtype"); return NL_ERR_INVALID;
  }
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return NL_OK;
}

// ---- gather: out[i] = column[idx[i]] --------------------------------------------------------
// The column is given as its shards; a shard may live on another GPU of the box and is then read
// with peer loads over NVLink (the mappings exist once the communicator is up).  Negative indices
// wrap like NumPy's; indices outside [-total, total) are counted in *bad and read element 0.
struct GatherShards {
  const void *ptr[NL_MAX_DEVICES];
  long long start[NL_MAX_DEVICES + 1];   // first global index of each shard; start[n] = total
  int n;
};

template <typename T>
__global__ void __launch_bounds__(256) gather_kernel(T *__restrict__ out, const long long *__restrict__ idx, long long count,
                                                      const __grid_constant__ GatherShards sh, unsigned long long *bad) {
  const long long total = sh.start[sh.n];
  unsigned int my_bad = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    long long j = __ldcs(idx + i);
    if (j < 0) j += total;
    if (j < 0 || j >= total) { my_bad++; j = 0; }
    int s = 0;
#pragma unroll 1
    while (s + 1 < sh.n && j >= sh.start[s + 1]) s++;
    out[i] = total > 0 ? reinterpret_cast<const T *>(sh.ptr[s])[j - sh.start[s]] : T(0);
  }
  my_bad = __reduce_add_sync(0xffffffffu, my_bad);
  if ((threadIdx.x & 31) == 0 && my_bad) atomicAdd(bad, (unsigned long long)my_bad);
}

int launch_gather(void *stream, int sm_count, int dtype, void *out, const int64_t *idx, int64_t count, const void *const *shard_ptr,
                  const int64_t *shard_start, int n_shards, int64_t total, int64_t *bad) {
  GatherShards sh;
  memset(&sh, 0, sizeof(sh));
  sh.n = n_shards;
  for (int k = 0; k < n_shards; k++) { sh.ptr[k] = shard_ptr[k]; sh.start[k] = shard_start[k]; }
  sh.start[n_shards] = total;
  cudaStream_t st = (cudaStream_t)stream;
  int64_t blocks = (count + 255) / 256;
  const int64_t cap = (int64_t)sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  switch (nl_dtype_size(dtype)) {
    case 4: gather_kernel<uint32_t><<<(unsigned)blocks, 256, 0, st>>>((uint32_t *)out, (const long long *)idx, count, sh, (unsigned long long *)bad); break;
    case 8: gather_kernel<unsigned long long><<<(unsigned)blocks, 256, 0, st>>>((unsigned long long *)out, (const long long *)idx, count, sh, (unsigned long long *)bad); break;
    case 1: gather_kernel<unsigned char><<<(unsigned)blocks, 256, 0, st>>>((unsigned char *)out, (const long long *)idx, count, sh, (unsigned long long *)bad); break;
    case 2: gather_kernel<unsigned short><<<(unsigned)blocks, 256, 0, st>>>((unsigned short *)out, (const long long *)idx, count, sh, (unsigned long long *)bad); break;
    default: set_error("nl_gather: bad dtype"); return NL_ERR_INVALID;
  }
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return NL_OK;
}

static unsigned grid_for(int64_t work_items, int sm_count) {
  int64_t blocks = (work_items + 255) / 256;
  const int64_t cap = (int64_t)sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

template <typename T>
static void fill_dispatch(cudaStream_t st, int sm_count, void *dst, int64_t lo, int64_t step, int64_t count, uint64_t bits) {
  T value;
  memcpy(&value, &bits, sizeof(T));
  if (step == 1 && count >= 1024) {
    assign_fill_vec_kernel<T><<<grid_for(count / (16 / sizeof(T)), sm_count), 256, 0, st>>>((T *)dst + lo, count, value);
  } else {
    assign_fill_kernel<T><<<grid_for(count, sm_count), 256, 0, st>>>((T *)dst, lo, step, count, value);
  }
}

int launch_assign_fill(void *stream, int sm_count, void *dst, int dtype, int64_t lo, int64_t step, int64_t count, uint64_t bits) {
  cudaStream_t st = (cudaStream_t)stream;
  if (count <= 0) return NL_OK;
  switch (nl_dtype_size(dtype)) {            // a fill moves bit patterns: the item size is all that matters
    case 1: fill_dispatch<unsigned char>(st, sm_count, dst, lo, step, count, bits); break;
    case 2: fill_dispatch<unsigned short>(st, sm_count, dst, lo, step, count, bits); break;
    case 4: fill_dispatch<uint32_t>(st, sm_count, dst, lo, step, count, bits); break;
    case 8: fill_dispatch<unsigned long long>(st, sm_count, dst, lo, step, count, bits); break;
    default: set_error("bad dtype"); return NL_ERR_INVALID;
  }
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return NL_OK;
}

int launch_assign_copy(void *stream, int sm_count, void *dst, int dtype, int64_t lo, int64_t step, int64_t count, const void *src) {
  cudaStream_t st = (cudaStream_t)stream;
  if (count <= 0) return NL_OK;
  const unsigned g = grid_for(count, sm_count);
  switch (nl_dtype_size(dtype)) {
    case 1: assign_copy_kernel<unsigned char><<<g, 256, 0, st>>>((unsigned char *)dst, lo, step, count, (const unsigned char *)src); break;
    case 2: assign_copy_kernel<unsigned short><<<g, 256, 0, st>>>((unsigned short *)dst, lo, step, count, (const unsigned short *)src); break;
    case 4: assign_copy_kernel<uint32_t><<<g, 256, 0, st>>>((uint32_t *)dst, lo, step, count, (const uint32_t *)src); break;
    case 8: assign_copy_kernel<unsigned long long><<<g, 256, 0, st>>>((unsigned long long *)dst, lo, step, count, (const unsigned long long *)src); break;
    default: set_error("bad dtype"); return NL_ERR_INVALID;
  }
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return NL_OK;
}

int launch_scan_offsets(void *stream, const int64_t *counts, int n, int64_t *offsets) {
  scan_offsets_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(counts, n, offsets);
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return NL_OK;
}

int launch_fill(void *stream, int sm_count, void *dst, int dtype, int64_t first_index, int64_t count, uint64_t seed, int kind) {
  cudaStream_t st = (cudaStream_t)stream;
  if (count <= 0) return NL_OK;
  const unsigned g = grid_for(count, sm_count);
  switch (dtype) {
    case NL_BOOL: fill_kernel<unsigned char, true><<<g, 256, 0, st>>>((unsigned char *)dst, first_index, count, seed, kind); break;
#define NL_CASE(dt, T, sfx) case dt: fill_kernel<T><<<g, 256, 0, st>>>((T *)dst, first_index, count, seed, kind); break;
    NL_FOR_VALUE_DTYPES(NL_CASE)
#undef NL_CASE
    default: set_error("bad dtype"); return NL_ERR_INVALID;
  }
  CUDA_TRY(cudaGetLastError());
  count_launch();
  return NL_OK;
}

}  // namespace nl
