/*
 * nl_oracle.c — CPU restatement of numpyllvm's pipeline execution.
 * TEST INFRASTRUCTURE ONLY (see nl_oracle.h).  Follows, file:line relative to
 * the reference checkout:
 *   loop skeleton, index counters, result_sizes   compiler.cpp:309-538
 *   operand load / immediates / materialise store compiler.cpp:202-269, 125-182
 *   `*`                                           thunk_as_number.cpp:45-47
 *   `>=`                                          thunk_as_number.cpp:9-12
 *   `[mask]` filter                               thunk_as_mapping.cpp:9-35
 *   max / sum                                     thunk_methods.cpp:70-162
 *   range partition                               scheduler.cpp:133-147
 *   output allocation (zero-filled upper bound)   scheduler.cpp:107-122
 *   gather / finalize                             compiler.cpp:30-49, thunk_methods.cpp:133-151
 * Deliberate departures (documented in DESIGN.md): the gather uses the real
 * item size instead of the hard-coded 8 (compiler.cpp:39); max/min
 * get sum's partial + finalize structure and dtype-lowest/highest identities
 * (Q6); float ops use fmul/fcmp (Q1); boolean stores are 0x01 (Q4).
 */
#include "nl_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
#define T int32_t
#define UT uint32_t
#define ACC int64_t
#define SFX i32
#define IS_FLOAT 0
#define IS_UNSIGNED 0
#define T_LOWEST INT32_MIN
#define T_HIGHEST INT32_MAX
#include "nl_oracle_eval.inc"

#define T int64_t
#define UT uint64_t
#define ACC int64_t
#define SFX i64
#define IS_FLOAT 0
#define IS_UNSIGNED 0
#define T_LOWEST INT64_MIN
#define T_HIGHEST INT64_MAX
#include "nl_oracle_eval.inc"

#define T float
#define UT uint32_t
#define ACC double
#define SFX f32
#define IS_FLOAT 1
#define IS_UNSIGNED 0
#define T_LOWEST (-INFINITY)
#define T_HIGHEST (INFINITY)
#include "nl_oracle_eval.inc"

#define T double
#define UT uint64_t
#define ACC double
#define SFX f64
#define IS_FLOAT 1
#define IS_UNSIGNED 0
#define T_LOWEST (-INFINITY)
#define T_HIGHEST (INFINITY)
#include "nl_oracle_eval.inc"

/* the narrow and the unsigned integer types */
#define T int8_t
#define UT uint8_t
#define ACC int64_t
#define SFX i8
#define IS_FLOAT 0
#define IS_UNSIGNED 0
#define T_LOWEST INT8_MIN
#define T_HIGHEST INT8_MAX
#include "nl_oracle_eval.inc"

#define T int16_t
#define UT uint16_t
#define ACC int64_t
#define SFX i16
#define IS_FLOAT 0
#define IS_UNSIGNED 0
#define T_LOWEST INT16_MIN
#define T_HIGHEST INT16_MAX
#include "nl_oracle_eval.inc"

#define T uint8_t
#define UT uint8_t
#define ACC uint64_t
#define SFX u8
#define IS_FLOAT 0
#define IS_UNSIGNED 1
#define T_LOWEST 0
#define T_HIGHEST UINT8_MAX
#include "nl_oracle_eval.inc"

#define T uint16_t
#define UT uint16_t
#define ACC uint64_t
#define SFX u16
#define IS_FLOAT 0
#define IS_UNSIGNED 1
#define T_LOWEST 0
#define T_HIGHEST UINT16_MAX
#include "nl_oracle_eval.inc"

#define T uint32_t
#define UT uint32_t
#define ACC uint64_t
#define SFX u32
#define IS_FLOAT 0
#define IS_UNSIGNED 1
#define T_LOWEST 0
#define T_HIGHEST UINT32_MAX
#include "nl_oracle_eval.inc"

#define T uint64_t
#define UT uint64_t
#define ACC uint64_t
#define SFX u64
#define IS_FLOAT 0
#define IS_UNSIGNED 1
#define T_LOWEST 0
#define T_HIGHEST UINT64_MAX
#include "nl_oracle_eval.inc"

/* ------------------------------------------------------------------------- */
static size_t dsize(int dtype) {
  switch (dtype) {
    case NL_BOOL: case NL_I8: case NL_U8: return 1;
    case NL_I16: case NL_U16: return 2;
    case NL_I32: case NL_U32: case NL_F32: return 4;
    case NL_I64: case NL_U64: case NL_F64: return 8;
  }
  return 0;
}

void nlo_partition(int64_t size, int thread_count, int j, int64_t *start, int64_t *end) {
  /* increment = size / thread_count; last chunk takes the remainder (scheduler.cpp:133-139) */
  int64_t inc = size / thread_count;
  *start = inc * j;
  *end = (j == thread_count - 1) ? size : inc * (j + 1);
}

/* Emission-order walk that records which index counter governs each
 * materialised output (DataElement::index_addr, compiler.cpp:254). */
typedef struct { int space; int counter_node; } idx_state;
static void walk_spaces(const nl_expr_node *nodes, int n, idx_state *st,
                        int *out_space, int *out_counter_node, int *out_dtype, int *out_op,
                        const nl_input_desc *in, int value_dtype) {
  const nl_expr_node *nd = &nodes[n];
  if (nd->op != NL_OP_INPUT) {
    if (nd->lhs >= 0) walk_spaces(nodes, nd->lhs, st, out_space, out_counter_node, out_dtype, out_op, in, value_dtype);
    if (nd->rhs >= 0) walk_spaces(nodes, nd->rhs, st, out_space, out_counter_node, out_dtype, out_op, in, value_dtype);
    if (nd->op == NL_OP_FILTER) { st->space = NL_IDX_COMPACT; st->counter_node = n; }
    if (nd->op == NL_OP_MAX || nd->op == NL_OP_MIN || nd->op == NL_OP_SUM) st->space = NL_IDX_SHARD;
  }
  if (nd->output >= 0) {
    int is_bool = (nd->op >= NL_OP_GE && nd->op <= NL_OP_NOT) ||
                  (nd->op == NL_OP_INPUT && in[nd->input].dtype == NL_BOOL);
    out_space[nd->output] = (nd->op == NL_OP_WHERE_STORE) ? NL_IDX_WHERE : st->space;
    out_counter_node[nd->output] = st->counter_node;
    out_dtype[nd->output] = is_bool ? NL_BOOL : value_dtype;
    out_op[nd->output] = nd->op;
  }
}

typedef struct {
  const nl_expr_node *nodes; int n_nodes; const nl_input_desc *in;
  const void *const *inputs; void *const *outputs;
  const int *out_space; const int *out_counter_node; int n_outputs;
  int value_dtype; int64_t size; int thread_count;
  int64_t *output_starts;            /* [thread_count]            */
  int64_t (*output_ends)[64];        /* [n_outputs][thread_count] */
  int first, stride;                 /* chunks handled by this worker */
} job_t;

static void run_chunk(job_t *jb, int j) {
  int64_t s, e, sizes[NL_MAX_OUTPUTS];
  nlo_partition(jb->size, jb->thread_count, j, &s, &e);
#define CHUNK(sfx) chunk_##sfx(jb->nodes, jb->n_nodes, jb->in, jb->inputs, jb->outputs, jb->out_space, jb->out_counter_node, jb->n_outputs, s, e, j, sizes)
  switch (jb->value_dtype) {
    case NL_I32: CHUNK(i32); break;
    case NL_I64: CHUNK(i64); break;
    case NL_F32: CHUNK(f32); break;
    case NL_I8: CHUNK(i8); break;
    case NL_I16: CHUNK(i16); break;
    case NL_U8: CHUNK(u8); break;
    case NL_U16: CHUNK(u16); break;
    case NL_U32: CHUNK(u32); break;
    case NL_U64: CHUNK(u64); break;
    default:     CHUNK(f64); break;
  }
#undef CHUNK
  /* ExecuteFunction bookkeeping (compiler.cpp:88-91) */
  jb->output_starts[j] = s;
  for (int k = 0; k < jb->n_outputs; k++) jb->output_ends[k][j] = sizes[k];
}

static void *worker(void *arg) {
  job_t *jb = (job_t *)arg;
  for (int j = jb->first; j < jb->thread_count; j += jb->stride) run_chunk(jb, j);
  return NULL;
}

int nlo_run(const nl_expr_node *nodes, int n_nodes, const nl_input_desc *in, int n_inputs,
            const void *const *inputs, void *const *outputs, int n_outputs,
            int64_t size, int thread_count, int n_workers, int64_t *out_counts) {
  if (!nodes || n_nodes <= 0 || n_nodes > NL_MAX_NODES || n_outputs > NL_MAX_OUTPUTS ||
      n_inputs > NL_MAX_INPUTS || thread_count <= 0 || thread_count > 64 || size < 0)
    return NL_ERR_INVALID;
  /* value dtype of the pipeline = dtype of its non-bool inputs (Q3: equal dtypes required) */
  int value_dtype = -1;
  for (int k = 0; k < n_inputs; k++) {          /* arrays decide; two arrays must agree */
    if (in[k].dtype == NL_BOOL || in[k].kind != NL_IN_ARRAY) continue;
    if (value_dtype >= 0 && in[k].dtype != value_dtype) return NL_ERR_UNSUPPORTED;
    value_dtype = in[k].dtype;
  }
  for (int k = 0; k < n_inputs && value_dtype < 0; k++)   /* scalars only: they are pre-converted */
    if (in[k].dtype != NL_BOOL) value_dtype = in[k].dtype;
  if (value_dtype < 0) value_dtype = NL_I64;

  int out_space[NL_MAX_OUTPUTS], out_counter_node[NL_MAX_OUTPUTS], out_dtype[NL_MAX_OUTPUTS], out_op[NL_MAX_OUTPUTS];
  for (int k = 0; k < NL_MAX_OUTPUTS; k++) { out_space[k] = NL_IDX_LOOP; out_counter_node[k] = 0; out_dtype[k] = value_dtype; out_op[k] = 0; }
  idx_state st = { NL_IDX_LOOP, 0 };
  walk_spaces(nodes, n_nodes - 1, &st, out_space, out_counter_node, out_dtype, out_op, in, value_dtype);

  /* ScheduleFunction: allocate outputs (scheduler.cpp:107-122).  Reductions get a
   * per-thread partial array (llvm_initialize_sum_data, thunk_methods.cpp:153-162). */
  void *eff_outputs[NL_MAX_OUTPUTS];
  void *partials[NL_MAX_OUTPUTS];
  for (int k = 0; k < n_outputs; k++) {
    partials[k] = NULL;
    eff_outputs[k] = outputs[k];
    if (out_space[k] == NL_IDX_SHARD) {
      partials[k] = calloc((size_t)thread_count, 8);
      if (!partials[k]) return NL_ERR_NOMEM;
      switch (value_dtype) {
        case NL_I32: init_partials_i32(out_op[k], partials[k], thread_count); break;
        case NL_I64: init_partials_i64(out_op[k], partials[k], thread_count); break;
        case NL_F32: init_partials_f32(out_op[k], partials[k], thread_count); break;
        case NL_I8: init_partials_i8(out_op[k], partials[k], thread_count); break;
        case NL_I16: init_partials_i16(out_op[k], partials[k], thread_count); break;
        case NL_U8: init_partials_u8(out_op[k], partials[k], thread_count); break;
        case NL_U16: init_partials_u16(out_op[k], partials[k], thread_count); break;
        case NL_U32: init_partials_u32(out_op[k], partials[k], thread_count); break;
        case NL_U64: init_partials_u64(out_op[k], partials[k], thread_count); break;
        default:     init_partials_f64(out_op[k], partials[k], thread_count); break;
      }
      eff_outputs[k] = partials[k];
    } else if (out_space[k] != NL_IDX_WHERE) {
      memset(outputs[k], 0, (size_t)size * dsize(out_dtype[k]));   /* PyArray_ZEROS */
    }
  }

  int64_t output_starts[64];
  int64_t output_ends[NL_MAX_OUTPUTS][64];
  job_t base = { nodes, n_nodes, in, inputs, eff_outputs, out_space, out_counter_node, n_outputs,
                 value_dtype, size, thread_count, output_starts, output_ends, 0, 1 };
  if (n_workers <= 1) {
    worker(&base);   /* the single worker drains the queue in order (scheduler.cpp:161-219) */
  } else {
    if (n_workers > thread_count) n_workers = thread_count;
    pthread_t th[64]; job_t jobs[64];
    for (int w = 0; w < n_workers; w++) { jobs[w] = base; jobs[w].first = w; jobs[w].stride = n_workers; pthread_create(&th[w], NULL, worker, &jobs[w]); }
    for (int w = 0; w < n_workers; w++) pthread_join(th[w], NULL);
  }

  /* JITFunctionDECREF: gather (compiler.cpp:30-49) */
  for (int k = 0; k < n_outputs; k++) {
    if (out_space[k] == NL_IDX_SHARD) {
      switch (value_dtype) {
        case NL_I32: finalize_i32(out_op[k], partials[k], thread_count, outputs[k]); break;
        case NL_I64: finalize_i64(out_op[k], partials[k], thread_count, outputs[k]); break;
        case NL_F32: finalize_f32(out_op[k], partials[k], thread_count, outputs[k]); break;
        case NL_I8: finalize_i8(out_op[k], partials[k], thread_count, outputs[k]); break;
        case NL_I16: finalize_i16(out_op[k], partials[k], thread_count, outputs[k]); break;
        case NL_U8: finalize_u8(out_op[k], partials[k], thread_count, outputs[k]); break;
        case NL_U16: finalize_u16(out_op[k], partials[k], thread_count, outputs[k]); break;
        case NL_U32: finalize_u32(out_op[k], partials[k], thread_count, outputs[k]); break;
        case NL_U64: finalize_u64(out_op[k], partials[k], thread_count, outputs[k]); break;
        default:     finalize_f64(out_op[k], partials[k], thread_count, outputs[k]); break;
      }
      free(partials[k]);
      if (out_counts) out_counts[k] = 1;
    } else {
      char *data = (char *)outputs[k];
      size_t es = dsize(out_dtype[k]);     /* reference hard-codes sizeof(long long) (Q5) */
      for (int j = 1; j < thread_count; j++) {
        if (output_ends[k][j - 1] != output_starts[j]) {
          memmove(data + output_ends[k][j - 1] * es, data + output_starts[j] * es,
                  (size_t)(output_ends[k][j] - output_starts[j]) * es);
          output_ends[k][j] = output_ends[k][j - 1] + (output_ends[k][j] - output_starts[j]);
        }
      }
      if (out_counts) out_counts[k] = output_ends[k][thread_count - 1];
    }
  }
  return NL_OK;
}

/* ---- assignment: NumPy setitem on the storage (thunk_as_mapping.cpp:84) -- */
int nlo_assign_fill(void *dst, int dtype, int64_t lo, int64_t step, int64_t count, uint64_t value_bits) {
  size_t es = dsize(dtype);
  if (!es) return NL_ERR_INVALID;
  for (int64_t k = 0; k < count; k++) memcpy((char *)dst + (lo + k * step) * (int64_t)es, &value_bits, es);
  return NL_OK;
}
int nlo_assign_copy(void *dst, int dtype, int64_t lo, int64_t step, int64_t count, const void *src) {
  size_t es = dsize(dtype);
  if (!es) return NL_ERR_INVALID;
  for (int64_t k = 0; k < count; k++) memcpy((char *)dst + (lo + k * step) * (int64_t)es, (const char *)src + k * (int64_t)es, es);
  return NL_OK;
}

/* ---- synthetic inputs ---------------------------------------------------- */
static inline uint64_t splitmix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
int nlo_fill(void *dst, int dtype, int64_t first_index, int64_t count, uint64_t seed, int kind) {
  for (int64_t k = 0; k < count; k++) {
    uint64_t z = splitmix64(seed ^ (uint64_t)(first_index + k));
    switch (dtype) {
      case NL_I32: case NL_I64: case NL_I8: case NL_I16: case NL_U8: case NL_U16: case NL_U32: case NL_U64: {
        int64_t v = (kind == 0) ? (int64_t)(1 + (((z >> 32) * 99ULL) >> 32)) : (int64_t)z;
        switch (dsize(dtype)) {          /* the low bytes: the same bits for the signed and the unsigned type */
          case 1: ((int8_t *)dst)[k] = (int8_t)v; break;
          case 2: ((int16_t *)dst)[k] = (int16_t)v; break;
          case 4: ((int32_t *)dst)[k] = (int32_t)v; break;
          default: ((int64_t *)dst)[k] = v; break;
        }
        break;
      }
      case NL_F32: {
        uint32_t bits = 0x3f800000u | (uint32_t)(z >> 41);
        float f; memcpy(&f, &bits, 4);
        ((float *)dst)[k] = (kind == 3) ? f - 1.0f : f;
        break;
      }
      case NL_F64: {
        uint64_t bits = 0x3ff0000000000000ULL | (z >> 12);
        double d; memcpy(&d, &bits, 8);
        ((double *)dst)[k] = (kind == 3) ? d - 1.0 : d;
        break;
      }
      case NL_BOOL: ((uint8_t *)dst)[k] = (uint8_t)(z & 1); break;
      default: return NL_ERR_INVALID;
    }
  }
  return NL_OK;
}

/* ======================================================================== *
 * Fixed-shape loops: the scalar loop the reference would JIT for each       *
 * benchmark pipeline, written out so gcc -O3 can vectorise it the way       *
 * LLVM's LoopVectorize pass would (optimizer.cpp:76).  Same chunking and    *
 * gather as above.                                                          *
 * ======================================================================== */
typedef void (*chunk_fn)(void *arg, int j, int64_t s, int64_t e);
typedef struct { chunk_fn fn; void *arg; int64_t n; int thread_count; int first, stride; } fast_job;
static void *fast_worker(void *p) {
  fast_job *jb = (fast_job *)p;
  for (int j = jb->first; j < jb->thread_count; j += jb->stride) {
    int64_t s, e; nlo_partition(jb->n, jb->thread_count, j, &s, &e);
    jb->fn(jb->arg, j, s, e);
  }
  return NULL;
}
static void fast_run(chunk_fn fn, void *arg, int64_t n, int thread_count, int n_workers) {
  fast_job base = { fn, arg, n, thread_count, 0, 1 };
  if (n_workers <= 1) { fast_worker(&base); return; }
  if (n_workers > thread_count) n_workers = thread_count;
  pthread_t th[256]; fast_job jobs[256];
  if (n_workers > 256) n_workers = 256;
  for (int w = 0; w < n_workers; w++) { jobs[w] = base; jobs[w].first = w; jobs[w].stride = n_workers; pthread_create(&th[w], NULL, fast_worker, &jobs[w]); }
  for (int w = 0; w < n_workers; w++) pthread_join(th[w], NULL);
}

typedef struct { const float *a, *b; float *c; } mul2_arg;
static void mul2_chunk(void *p, int j, int64_t s, int64_t e) {
  (void)j; mul2_arg *g = (mul2_arg *)p;
  const float *restrict a = g->a; const float *restrict b = g->b; float *restrict c = g->c;
  for (int64_t i = s; i < e; i++) c[i] = a[i] * b[i];
}
int nlo_fast_mul2_f32(const float *a, const float *b, float *c, int64_t n, int thread_count, int n_workers) {
  mul2_arg g = { a, b, c };
  fast_run(mul2_chunk, &g, n, thread_count, n_workers);
  return NL_OK;
}

typedef struct { const double *const *in; double *out; } mul6_arg;
static void mul6_chunk(void *p, int j, int64_t s, int64_t e) {
  (void)j; mul6_arg *g = (mul6_arg *)p;
  const double *restrict a = g->in[0]; const double *restrict b = g->in[1]; const double *restrict c = g->in[2];
  const double *restrict d = g->in[3]; const double *restrict ee = g->in[4]; const double *restrict f = g->in[5];
  double *restrict o = g->out;
  /* (d*e*f) * (a*b*c), Tests/multiplication.py:26-27 with the sort left out; in6 = {d,e,f,a,b,c} */
  for (int64_t i = s; i < e; i++) o[i] = ((a[i] * b[i]) * c[i]) * ((d[i] * ee[i]) * f[i]);
}
int nlo_fast_mul6_f64(const double *const *in6, double *out, int64_t n, int thread_count, int n_workers) {
  mul6_arg g = { in6, out };
  fast_run(mul6_chunk, &g, n, thread_count, n_workers);
  return NL_OK;
}

typedef struct { const double *a; double t; double *out; int64_t starts[256], ends[256]; } filt_arg;
static void filt_chunk(void *p, int j, int64_t s, int64_t e) {
  filt_arg *g = (filt_arg *)p;
  const double *restrict a = g->a; double *restrict o = g->out; const double t = g->t;
  int64_t cnt = s;
  for (int64_t i = s; i < e; i++) { if (!(a[i] > t)) continue; o[cnt++] = a[i]; }
  g->starts[j] = s; g->ends[j] = cnt;
}
int nlo_fast_filter_gt_f64(const double *a, double t, double *out, int64_t n, int thread_count, int n_workers, int64_t *count) {
  if (thread_count > 256) return NL_ERR_INVALID;
  filt_arg *g = (filt_arg *)malloc(sizeof(filt_arg));
  if (!g) return NL_ERR_NOMEM;
  g->a = a; g->t = t; g->out = out;
  memset(out, 0, (size_t)n * 8);                       /* PyArray_ZEROS (scheduler.cpp:116) */
  fast_run(filt_chunk, g, n, thread_count, n_workers);
  for (int j = 1; j < thread_count; j++) {             /* compiler.cpp:36-46 */
    if (g->ends[j - 1] != g->starts[j]) {
      memmove(out + g->ends[j - 1], out + g->starts[j], (size_t)(g->ends[j] - g->starts[j]) * 8);
      g->ends[j] = g->ends[j - 1] + (g->ends[j] - g->starts[j]);
    }
  }
  *count = g->ends[thread_count - 1];
  free(g);
  return NL_OK;
}

typedef struct { const void *a; double pd[256]; int64_t pi[256]; float t, k; float pf[256]; } red_arg;
static void maxf64_chunk(void *p, int j, int64_t s, int64_t e) {
  red_arg *g = (red_arg *)p; const double *restrict a = (const double *)g->a;
  double cur = -INFINITY;
  for (int64_t i = s; i < e; i++) { double l = a[i]; if (l > cur || l != l) cur = l; }
  g->pd[j] = cur;
}
int nlo_fast_max_f64(const double *a, int64_t n, int thread_count, int n_workers, double *out) {
  if (thread_count > 256) return NL_ERR_INVALID;
  red_arg *g = (red_arg *)malloc(sizeof(red_arg)); if (!g) return NL_ERR_NOMEM;
  g->a = a;
  fast_run(maxf64_chunk, g, n, thread_count, n_workers);
  double d = -INFINITY;
  for (int j = 0; j < thread_count; j++) if (g->pd[j] > d || g->pd[j] != g->pd[j]) d = g->pd[j];
  *out = d; free(g); return NL_OK;
}
static void maxi64_chunk(void *p, int j, int64_t s, int64_t e) {
  red_arg *g = (red_arg *)p; const int64_t *restrict a = (const int64_t *)g->a;
  int64_t cur = INT64_MIN;
  for (int64_t i = s; i < e; i++) { if (a[i] > cur) cur = a[i]; }
  g->pi[j] = cur;
}
int nlo_fast_max_i64(const int64_t *a, int64_t n, int thread_count, int n_workers, int64_t *out) {
  if (thread_count > 256) return NL_ERR_INVALID;
  red_arg *g = (red_arg *)malloc(sizeof(red_arg)); if (!g) return NL_ERR_NOMEM;
  g->a = a;
  fast_run(maxi64_chunk, g, n, thread_count, n_workers);
  int64_t d = INT64_MIN;
  for (int j = 0; j < thread_count; j++) if (g->pi[j] > d) d = g->pi[j];
  *out = d; free(g); return NL_OK;
}
static void sumi64_chunk(void *p, int j, int64_t s, int64_t e) {
  red_arg *g = (red_arg *)p; const int64_t *restrict a = (const int64_t *)g->a;
  uint64_t cur = 0;
  for (int64_t i = s; i < e; i++) cur += (uint64_t)a[i];
  g->pi[j] = (int64_t)cur;
}
int nlo_fast_sum_i64(const int64_t *a, int64_t n, int thread_count, int n_workers, int64_t *out) {
  if (thread_count > 256) return NL_ERR_INVALID;
  red_arg *g = (red_arg *)malloc(sizeof(red_arg)); if (!g) return NL_ERR_NOMEM;
  g->a = a;
  fast_run(sumi64_chunk, g, n, thread_count, n_workers);
  uint64_t d = 0;
  for (int j = 0; j < thread_count; j++) d += (uint64_t)g->pi[j];
  *out = (int64_t)d; free(g); return NL_OK;
}
static void fmm_chunk(void *p, int j, int64_t s, int64_t e) {
  red_arg *g = (red_arg *)p; const float *restrict a = (const float *)g->a;
  const float t = g->t, k = g->k;
  double cur = -INFINITY;   /* ACC = double for floats */
  for (int64_t i = s; i < e; i++) {
    float l = a[i];
    if (!(l >= t)) continue;            /* filter: skip to for.inc */
    float v = l * k;
    if ((double)v > cur || v != v) cur = (double)v;
  }
  g->pf[j] = (float)cur;
}
int nlo_fast_filter_mul_max_f32(const float *a, float t, float k, int64_t n, int thread_count, int n_workers, float *out) {
  if (thread_count > 256) return NL_ERR_INVALID;
  red_arg *g = (red_arg *)malloc(sizeof(red_arg)); if (!g) return NL_ERR_NOMEM;
  g->a = a; g->t = t; g->k = k;
  fast_run(fmm_chunk, g, n, thread_count, n_workers);
  float d = -INFINITY;
  for (int j = 0; j < thread_count; j++) if (g->pf[j] > d || g->pf[j] != g->pf[j]) d = g->pf[j];
  *out = d; free(g); return NL_OK;
}
