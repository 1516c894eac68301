// optimizer.cpp — see optimizer.hpp.
#include "optimizer.hpp"

#include <cstring>

static bool dt_is_float(int t) { return t == NL_F32 || t == NL_F64; }
static bool dt_is_unsigned(int t) { return t == NL_BOOL || t == NL_U8 || t == NL_U16 || t == NL_U32 || t == NL_U64; }

// the value behind a bit pattern of type `dtype`, in the widest type of its kind
template <typename W>
static W bits_as(uint64_t bits, int dtype) {
  switch (dtype) {
    case NL_BOOL: return (W)((bits & 0xff) ? 1 : 0);
    case NL_I8: { int8_t v; memcpy(&v, &bits, 1); return (W)v; }
    case NL_I16: { int16_t v; memcpy(&v, &bits, 2); return (W)v; }
    case NL_I32: { int32_t v; memcpy(&v, &bits, 4); return (W)v; }
    case NL_I64: { int64_t v; memcpy(&v, &bits, 8); return (W)v; }
    case NL_U8: { uint8_t v; memcpy(&v, &bits, 1); return (W)v; }
    case NL_U16: { uint16_t v; memcpy(&v, &bits, 2); return (W)v; }
    case NL_U32: { uint32_t v; memcpy(&v, &bits, 4); return (W)v; }
    case NL_U64: { uint64_t v; memcpy(&v, &bits, 8); return (W)v; }
    case NL_F32: { float v; memcpy(&v, &bits, 4); return (W)v; }
    default: { double v; memcpy(&v, &bits, 8); return (W)v; }
  }
}

template <typename D>
static uint64_t store_as(uint64_t bits, int from) {
  D v = dt_is_float(from) ? (D)bits_as<double>(bits, from) : dt_is_unsigned(from) ? (D)bits_as<uint64_t>(bits, from) : (D)bits_as<int64_t>(bits, from);
  uint64_t out = 0;
  memcpy(&out, &v, sizeof(D));
  return out;
}

// getLLVMConstant (compiler.cpp:125-154) baked a size-1 input as an immediate of the column
// type; the conversion to the pipeline's value type happens here, on the host (C conversion rules)
uint64_t nljit_convert_scalar(uint64_t bits, int from, int to) {
  if (from == to) return bits;
  switch (to) {
    case NL_I8: return store_as<int8_t>(bits, from);
    case NL_I16: return store_as<int16_t>(bits, from);
    case NL_I32: return store_as<int32_t>(bits, from);
    case NL_I64: return store_as<int64_t>(bits, from);
    case NL_U8: return store_as<uint8_t>(bits, from);
    case NL_U16: return store_as<uint16_t>(bits, from);
    case NL_U32: return store_as<uint32_t>(bits, from);
    case NL_U64: return store_as<uint64_t>(bits, from);
    case NL_F32: return store_as<float>(bits, from);
    case NL_F64: return store_as<double>(bits, from);
    default: return (bits_as<double>(bits, from) != 0.0) ? 1 : 0;
  }
}
static uint64_t convert_scalar(uint64_t bits, int from, int to) { return nljit_convert_scalar(bits, from, to); }

struct Lowering {
  Pipeline *pipeline;
  LoweredPipeline *out;
  bool failed;
  bool too_big;
};

static DataSource *find_input(Pipeline *p, Operation *op) {
  for (Binding &e : p->inputs)
    if (e.operation == op) return e.source;
  return NULL;
}

static int find_output(Pipeline *p, Operation *op) {
  int k = 0;
  for (Binding &e : p->outputs) {
    if (e.operation == op) return k;
    k++;
  }
  return -1;
}

static int lower_op(Lowering &lw, Operation *op) {
  LoweredPipeline *o = lw.out;
  if (lw.failed || lw.too_big) return 0;
  nl_expr_node nd = {NL_OP_INPUT, -1, -1, -1, -1};
  if (op->kind == Operation::Column || op->kind == Operation::ChildResult) {
    DataSource *src = find_input(lw.pipeline, op);
    if (!src) {
      PyErr_SetString(PyExc_ValueError, "Column/Pipeline operation found, but not found in inputData.");   // compiler.cpp:225
      lw.failed = true;
      return 0;
    }
    // two leaves of the same thunk share one input slot (one column, read once per use)
    int k = -1;
    for (int i = 0; i < o->n_inputs; i++)
      if (o->input_sources[i] == src || (src->object && o->input_sources[i]->object == src->object)) { k = i; break; }
    if (k < 0) {
      if (o->n_inputs >= NL_MAX_INPUTS) { lw.too_big = true; return 0; }
      k = o->n_inputs++;
      o->input_sources[k] = src;
    }
    nd.input = k;
  } else {
    ThunkOperation *top = (op)->op;
    if (!top || top->opcode < 0) {
      PyErr_SetString(PyExc_ValueError, "operation has no kernel op-code and is not a pipeline root");
      lw.failed = true;
      return 0;
    }
    nd.op = top->opcode;
    if (op->kind == Operation::Apply1) {
      nd.lhs = lower_op(lw, op->lhs);
    } else {
      nd.lhs = lower_op(lw, op->lhs);
      nd.rhs = lower_op(lw, op->rhs);
    }
  }
  if (op->materialize) nd.output = find_output(lw.pipeline, op);
  if (lw.failed || lw.too_big) return 0;
  if (o->n_nodes >= NL_MAX_NODES) { lw.too_big = true; return 0; }
  o->nodes[o->n_nodes] = nd;
  return o->n_nodes++;
}

static int subtree_size(Operation *op) {
  switch (op->kind) {
    case Operation::Apply1: return 1 + subtree_size(op->lhs);
    case Operation::Apply2: return 1 + subtree_size(op->lhs) + subtree_size(op->rhs);
    default: return 1;
  }
}

// the interior node whose subtree is closest to half of the expression: cutting there keeps
// both halves as fused as possible
static void pick_cut(Operation *op, Operation *root, int total, Operation **best, int *best_dist) {
  if (op->kind != Operation::Apply1 && op->kind != Operation::Apply2) return;
  if (op != root) {
    int d = subtree_size(op) * 2 - total;
    if (d < 0) d = -d;
    if (!*best || d < *best_dist) { *best = op; *best_dist = d; }
  }
  if (op->kind == Operation::Apply1) pick_cut(op->lhs, root, total, best, best_dist);
  else {
    pick_cut(op->lhs, root, total, best, best_dist);
    pick_cut(op->rhs, root, total, best, best_dist);
  }
}

// every FILTER node below the root (their results become columns of a later pipeline)
static void collect_filters(Operation *op, Operation *root, std::vector<PyThunkObject *> *out) {
  if (op->kind != Operation::Apply1 && op->kind != Operation::Apply2) return;
  if (op != root && op->op && op->op->opcode == NL_OP_FILTER && op->thunk) out->push_back(op->thunk);
  if (op->lhs) collect_filters(op->lhs, root, out);
  if (op->rhs) collect_filters(op->rhs, root, out);
}

int LowerPipeline(Pipeline *pipeline, LoweredPipeline *out, std::vector<PyThunkObject *> *cuts_out) {
  memset(out, 0, sizeof(*out));
  Lowering lw = {pipeline, out, false, false};
  lower_op(lw, pipeline->root);
  if (lw.failed) return -1;
  out->n_outputs = (int)pipeline->outputs.size();
  if (out->n_outputs > NL_MAX_OUTPUTS) lw.too_big = true;

  int rc = NL_ERR_SPLIT;
  if (!lw.too_big) {
    // value type of the pipeline: its (non-boolean) columns decide; scalars adopt it
    int dtype = -1;
    for (int k = 0; k < out->n_inputs && dtype < 0; k++) {
      const int t = nljit_npy_to_nl(out->input_sources[k]->type);
      if (t != NL_BOOL && out->input_sources[k]->size != 1) dtype = t;
    }
    for (int k = 0; k < out->n_inputs && dtype < 0; k++) {
      const int t = nljit_npy_to_nl(out->input_sources[k]->type);
      if (t != NL_BOOL) dtype = t;
    }
    if (dtype < 0) dtype = NL_I64;
    out->dtype = dtype;
    for (int k = 0; k < out->n_inputs; k++) {
      DataSource *src = out->input_sources[k];
      const int t = nljit_npy_to_nl(src->type);
      nl_input_desc &d = out->inputs[k];
      d.dtype = t;
      d.imm_bits = 0;
      if (src->size != 1) {
        d.kind = NL_IN_ARRAY;
      } else if (src->storage && src->storage->host_scalar) {
        // size-1 inputs are baked as immediates (compiler.cpp:210-211)
        d.kind = NL_IN_SCALAR_IMM;
        if (t != NL_BOOL) { d.imm_bits = convert_scalar(src->storage->scalar_bits, t, dtype); d.dtype = dtype; }
        else d.imm_bits = src->storage->scalar_bits & 0xff;
      } else {
        // the result of a child reduction stays on the device: no host round trip between pipelines
        d.kind = NL_IN_SCALAR_DEV;
        if (t != NL_BOOL && t != dtype) {
          PyErr_SetString(PyExc_TypeError, "a device-resident scalar must have the dtype of the pipeline's columns");
          return -1;
        }
      }
    }
    rc = nl_plan_compile(out->nodes, out->n_nodes, out->inputs, out->n_inputs, &out->plan);
    if (rc == NL_OK) return 0;
    if (rc == NL_ERR_SPLIT_FILTER) {
      // `a[m] * b[m]`, `f[f < 10]` with f = a[a > 5] still lazy: legal NumPy, but one loop cannot index a
      // column by the loop counter and by the compaction counter at once (compiler.cpp:213-216 would read
      // column[compacted index]).  The filters below the root get pipelines of their own; the root then
      // runs over their (equally ragged) results.
      collect_filters(pipeline->root, pipeline->root, cuts_out);
      if (!cuts_out->empty()) return 1;
    }
    if (rc != NL_ERR_SPLIT) {
      nljit_raise(rc, "jit: pipeline cannot be compiled");
      return -1;
    }
  }
  // too large for one fused kernel: propose a cut
  Operation *best = NULL;
  int dist = 0;
  pick_cut(pipeline->root, pipeline->root, subtree_size(pipeline->root), &best, &dist);
  if (!best || !best->thunk) {
    nljit_raise(rc, "jit: pipeline cannot be split further");
    return -1;
  }
  cuts_out->push_back(best->thunk);
  return 1;
}
