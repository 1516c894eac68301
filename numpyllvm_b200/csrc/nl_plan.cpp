// nl_plan.cpp — host-only part of the C-ABI: lowering an operation tree to the
// accumulator-machine program ("CompilePipeline" without LLVM, compiler.cpp:309),
// the plan dump (debug_printer.cpp's role) and the range partition
// (scheduler.cpp:133-147).  No CUDA in this file; it is testable without a GPU.
#include "nl_abi.h"
#include "nl_internal.h"
#include "nl_step.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace nl {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static bool is_cmp(int op) { return op >= NL_OP_GE && op <= NL_OP_NE; }
static bool is_boolop(int op) { return op >= NL_OP_AND && op <= NL_OP_NOT; }
static bool is_vop(int op) { return op == NL_OP_MUL || op == NL_OP_ADD || op == NL_OP_SUB || op == NL_OP_DIV || op == NL_OP_FLOORDIV; }
static bool is_reduce(int op) { return op == NL_OP_MAX || op == NL_OP_MIN || op == NL_OP_SUM; }

struct Lower {
  const nl_expr_node *nodes;
  int n_nodes;
  const nl_input_desc *in;
  int n_in;
  nl_plan *plan;
  int rc = NL_OK;

  bool temp_used[NL_MAX_TEMPS] = {false, false};
  int cache_budget = 0;                 // temporaries that may hold re-read columns
  bool ptemp_used[NL_MAX_PTEMPS] = {false, false, false};
  int max_temps = 0;
  int input_uses[NL_MAX_INPUTS] = {0};
  int input_temp[NL_MAX_INPUTS];
  int acc_input = -1;                   // acc currently holds array/scalar input k, or -1
  int out_space[NL_MAX_OUTPUTS];
  bool has_filter_below[NL_MAX_NODES];  // node or a descendant is a FILTER

  // emission-order state of the reference's code generator (info.index_addr)
  int cur_space = NL_IDX_LOOP;

  bool fail(int code, const char *fmt, ...) {
    if (rc == NL_OK) {
      rc = code;
      va_list ap;
      va_start(ap, fmt);
      vsnprintf(g_err, sizeof(g_err), fmt, ap);
      va_end(ap);
    }
    return false;
  }

  bool is_pred(int n) const {
    const nl_expr_node &nd = nodes[n];
    if (nd.op == NL_OP_INPUT) return in[nd.input].dtype == NL_BOOL;
    return is_cmp(nd.op) || is_boolop(nd.op);
  }
  bool simple_leaf(int n) const { return nodes[n].op == NL_OP_INPUT && nodes[n].output < 0; }

  void emit(uint32_t s) {
    if (plan->n_steps >= NL_MAX_STEPS) { fail(NL_ERR_SPLIT, "pipeline needs more than %d steps", NL_MAX_STEPS); return; }
    plan->step[plan->n_steps++] = s;
  }
  int alloc_temp() {
    for (int j = 0; j < NL_MAX_TEMPS; j++)
      if (!temp_used[j]) {
        temp_used[j] = true;
        if (j + 1 > max_temps) max_temps = j + 1;
        return j;
      }
    fail(NL_ERR_SPLIT, "expression needs more than %d value temporaries", NL_MAX_TEMPS);
    return 0;
  }
  int alloc_ptemp() {
    for (int j = 0; j < NL_MAX_PTEMPS; j++)
      if (!ptemp_used[j]) { ptemp_used[j] = true; return j; }
    fail(NL_ERR_SPLIT, "mask expression needs more than %d predicate temporaries", NL_MAX_PTEMPS);
    return 0;
  }

  // ---- pass 1: the reference's emission order (LHS, RHS, self) ------------------
  void walk(int n) {
    const nl_expr_node &nd = nodes[n];
    has_filter_below[n] = false;
    if (nd.op == NL_OP_INPUT) {
      if (nd.input < 0 || nd.input >= n_in) { fail(NL_ERR_INVALID, "node %d: bad input index", n); return; }
      input_uses[nd.input]++;
      // compiler.cpp:213-216 loads column[info.index]; once a filter has run that is the
      // compacted index, which is ill-defined for a full-length column
      if (in[nd.input].kind == NL_IN_ARRAY && cur_space == NL_IDX_COMPACT)
        fail(NL_ERR_SPLIT_FILTER, "array input %d is loaded after a filter in the same pipeline", nd.input);
    } else {
      if (nd.lhs < 0 || nd.lhs >= n) { fail(NL_ERR_INVALID, "node %d: bad lhs", n); return; }
      walk(nd.lhs);
      if (nd.rhs >= 0) {
        if (nd.rhs >= n) { fail(NL_ERR_INVALID, "node %d: bad rhs", n); return; }
        walk(nd.rhs);
      }
      has_filter_below[n] = has_filter_below[nd.lhs] || (nd.rhs >= 0 && has_filter_below[nd.rhs]);
      if (nd.op == NL_OP_FILTER) { cur_space = NL_IDX_COMPACT; has_filter_below[n] = true; plan->flags |= NL_PLAN_HAS_FILTER; }
      if (is_reduce(nd.op)) cur_space = NL_IDX_SHARD;
    }
    if (nd.output >= 0) {
      if (nd.output >= NL_MAX_OUTPUTS) { fail(NL_ERR_SPLIT, "more than %d materialised outputs", NL_MAX_OUTPUTS); return; }
      int space = (nd.op == NL_OP_WHERE_STORE) ? NL_IDX_WHERE : cur_space;
      if (space == NL_IDX_COMPACT && !has_filter_below[n])
        fail(NL_ERR_SPLIT_FILTER, "node %d is stored at a compacted index but does not depend on the filter", n);
      out_space[nd.output] = space;
      plan->out[nd.output].index_space = space;
      plan->out[nd.output].dtype = is_pred(n) ? NL_BOOL : plan->dtype;
      if (nd.output + 1 > plan->n_outputs) plan->n_outputs = nd.output + 1;
    }
  }

  // value temporaries the expression itself needs under the L-first evaluation order below
  // (a Sethi-Ullman style count); what is left over may cache columns that are read twice
  int temps_needed(int n) const {
    const nl_expr_node &nd = nodes[n];
    if (nd.op == NL_OP_INPUT) return 0;
    if (nd.rhs < 0) return temps_needed(nd.lhs);
    if (is_vop(nd.op) || is_cmp(nd.op)) {
      if (simple_leaf(nd.rhs)) return temps_needed(nd.lhs);
      if (simple_leaf(nd.lhs)) return temps_needed(nd.rhs);
      const int l = temps_needed(nd.lhs), r = 1 + temps_needed(nd.rhs);
      return l > r ? l : r;
    }
    const int l = temps_needed(nd.lhs), r = temps_needed(nd.rhs);   // filter / bool ops / where: no value held across
    return l > r ? l : r;
  }

  // ---- pass 2: code generation ----------------------------------------------------
  uint32_t src_of_leaf(int n) {
    int k = nodes[n].input;
    if (input_temp[k] >= 0) return SRC_TEMP | (uint32_t)input_temp[k];
    return (uint32_t)k;
  }

  void gen_leaf_value(int n) {
    int k = nodes[n].input;
    if (in[k].dtype == NL_BOOL) { fail(NL_ERR_INVALID, "boolean input %d used as a value", k); return; }
    if (acc_input == k) return;
    if (input_temp[k] >= 0) {
      emit(make_step(S_LOADT, (uint32_t)input_temp[k]));
    } else {
      emit(make_step(S_LOAD, (uint32_t)k));
      // a column read more than once (a[a > t]) is kept in a temporary instead of re-read
      if (input_uses[k] > 1 && in[k].kind == NL_IN_ARRAY && cache_budget > 0) {
        for (int j = NL_MAX_TEMPS - 1; j >= 0; j--)
          if (!temp_used[j]) {
            temp_used[j] = true;
            cache_budget--;
            if (j + 1 > max_temps) max_temps = j + 1;
            input_temp[k] = j;
            emit(make_step(S_SAVE, (uint32_t)j));
            break;
          }
      }
    }
    acc_input = k;
  }

  static uint32_t vop_of(int op, bool reversed) {
    switch (op) {
      case NL_OP_MUL: return V_MUL;
      case NL_OP_ADD: return V_ADD;
      case NL_OP_DIV: return reversed ? V_RDIV : V_DIV;
      case NL_OP_FLOORDIV: return reversed ? V_RFDIV : V_FDIV;
      default: return reversed ? V_RSUB : V_SUB;
    }
  }
  static uint32_t cop_of(int op, bool reversed) {
    switch (op) {
      case NL_OP_GE: return reversed ? C_LE : C_GE;
      case NL_OP_GT: return reversed ? C_LT : C_GT;
      case NL_OP_LE: return reversed ? C_GE : C_LE;
      case NL_OP_LT: return reversed ? C_GT : C_LT;
      case NL_OP_EQ: return C_EQ;
      default: return C_NE;
    }
  }

  // leaves "L op R" operands ready: returns the source byte of the second operand and
  // whether the operands ended up reversed (acc = R, X = L)
  void gen_operands(int l, int r, uint32_t *src, bool *reversed, int *tmp_to_free) {
    *tmp_to_free = -1;
    if (is_pred(l) || is_pred(r)) { fail(NL_ERR_INVALID, "boolean operand used in arithmetic / comparison"); *src = 0; *reversed = false; return; }
    if (simple_leaf(r)) {
      gen_value(l);
      *src = src_of_leaf(r);
      *reversed = false;
    } else if (simple_leaf(l)) {
      gen_value(r);
      *src = src_of_leaf(l);
      *reversed = true;
    } else {
      gen_value(l);
      int j = alloc_temp();
      emit(make_step(S_SAVE, (uint32_t)j));
      gen_value(r);
      *src = SRC_TEMP | (uint32_t)j;
      *reversed = true;
      *tmp_to_free = j;
    }
  }

  void gen_value(int n) {
    if (rc != NL_OK) return;
    const nl_expr_node &nd = nodes[n];
    if (nd.op == NL_OP_INPUT) {
      gen_leaf_value(n);
    } else if (is_vop(nd.op)) {
      if (nd.op == NL_OP_DIV && plan->dtype != NL_F32 && plan->dtype != NL_F64) {
        fail(NL_ERR_UNSUPPORTED, "true division of integer columns yields floats (no dtype promotion): use floor division");
        return;
      }
      uint32_t src; bool rev; int tf;
      gen_operands(nd.lhs, nd.rhs, &src, &rev, &tf);
      emit(make_step(S_BIN, vop_of(nd.op, rev), src));
      if (tf >= 0) temp_used[tf] = false;
      acc_input = -1;
    } else if (nd.op == NL_OP_FILTER) {
      if (!is_pred(nd.rhs) || is_pred(nd.lhs)) { fail(NL_ERR_INVALID, "filter needs (value, mask)"); return; }
      // mask first: its evaluation clobbers acc, and nothing it does depends on the value
      gen_pred(nd.rhs);
      emit(make_step(S_FILTER));
      gen_value(nd.lhs);
    } else if (is_reduce(nd.op)) {
      if (n != n_nodes - 1) { fail(NL_ERR_UNSUPPORTED, "a reduction must be the pipeline root (parallelbreaker)"); return; }
      if (nd.output < 0) { fail(NL_ERR_INVALID, "reduction root has no output"); return; }
      if (is_pred(nd.lhs)) { fail(NL_ERR_UNSUPPORTED, "reduction of a boolean"); return; }
      gen_value(nd.lhs);
      uint32_t rop = nd.op == NL_OP_MAX ? R_MAX : nd.op == NL_OP_MIN ? R_MIN : R_SUM;
      emit(make_step(S_REDUCE, rop, (uint32_t)nd.output));
      plan->flags |= NL_PLAN_HAS_REDUCE;
      plan->reduce_op = nd.op;
      return;   // the reduction's own store is the partial write
    } else if (nd.op == NL_OP_WHERE_STORE) {
      if (n != n_nodes - 1 || nd.output < 0) { fail(NL_ERR_UNSUPPORTED, "masked assignment must be the pipeline root"); return; }
      if (has_filter_below[n]) { fail(NL_ERR_UNSUPPORTED, "filter inside a masked assignment"); return; }
      if (!is_pred(nd.rhs) || is_pred(nd.lhs)) { fail(NL_ERR_INVALID, "masked assignment needs (value, mask)"); return; }
      gen_pred(nd.rhs);
      gen_value(nd.lhs);
      emit(make_step(S_WHERE, (uint32_t)nd.output));
      plan->flags |= NL_PLAN_HAS_WHERE;
      return;
    } else {
      fail(NL_ERR_INVALID, "node %d: op %d does not produce a value", n, nd.op);
      return;
    }
    if (nd.output >= 0) emit(make_step(S_STORE, (uint32_t)nd.output, (uint32_t)out_space[nd.output]));
  }

  void gen_pred(int n) {
    if (rc != NL_OK) return;
    const nl_expr_node &nd = nodes[n];
    if (nd.op == NL_OP_INPUT) {
      if (in[nd.input].dtype != NL_BOOL) { fail(NL_ERR_INVALID, "mask input %d is not boolean", nd.input); return; }
      emit(make_step(S_PLOAD, (uint32_t)nd.input));
    } else if (is_cmp(nd.op)) {
      uint32_t src; bool rev; int tf;
      gen_operands(nd.lhs, nd.rhs, &src, &rev, &tf);
      emit(make_step(S_CMP, cop_of(nd.op, rev), src));
      if (tf >= 0) temp_used[tf] = false;
    } else if (nd.op == NL_OP_NOT) {
      if (!is_pred(nd.lhs)) { fail(NL_ERR_INVALID, "~ needs a boolean operand"); return; }
      gen_pred(nd.lhs);
      emit(make_step(S_PNOT));
    } else if (is_boolop(nd.op)) {
      if (!is_pred(nd.lhs) || !is_pred(nd.rhs)) { fail(NL_ERR_INVALID, "boolean operator needs boolean operands"); return; }
      gen_pred(nd.lhs);
      int j = alloc_ptemp();
      emit(make_step(S_PSAVE, (uint32_t)j));
      gen_pred(nd.rhs);
      uint32_t pop = nd.op == NL_OP_AND ? P_AND : nd.op == NL_OP_OR ? P_OR : P_XOR;
      emit(make_step(S_PBIN, pop, (uint32_t)j));
      ptemp_used[j] = false;
    } else {
      fail(NL_ERR_INVALID, "node %d: op %d does not produce a mask", n, nd.op);
      return;
    }
    if (nd.output >= 0) emit(make_step(S_PSTORE, (uint32_t)nd.output, (uint32_t)out_space[nd.output]));
  }
};

}  // namespace nl

using namespace nl;

extern "C" {

int nl_abi_version(void) { return NL_ABI_VERSION; }
const char *nl_last_error(void) { return g_err; }

size_t nl_dtype_size(int dtype) {
  switch (dtype) {
    case NL_BOOL: case NL_I8: case NL_U8: return 1;
    case NL_I16: case NL_U16: return 2;
    case NL_I32: case NL_U32: case NL_F32: return 4;
    case NL_I64: case NL_U64: case NL_F64: return 8;
  }
  return 0;
}

int nl_plan_compile(const nl_expr_node *nodes, int n_nodes, const nl_input_desc *inputs, int n_inputs,
                    nl_plan *plan) {
  if (!nodes || !plan || n_nodes <= 0 || (n_inputs > 0 && !inputs)) { set_error("nl_plan_compile: null argument"); return NL_ERR_INVALID; }
  if (n_nodes > NL_MAX_NODES) { set_error("pipeline has %d nodes (max %d)", n_nodes, NL_MAX_NODES); return NL_ERR_SPLIT; }
  if (n_inputs > NL_MAX_INPUTS) { set_error("pipeline has %d inputs (max %d)", n_inputs, NL_MAX_INPUTS); return NL_ERR_SPLIT; }
  memset(plan, 0, sizeof(*plan));
  plan->abi_version = NL_ABI_VERSION;

  // value type: every non-boolean column must agree; scalars arrive pre-converted
  int dtype = -1;
  for (int k = 0; k < n_inputs; k++) {
    if (nl_dtype_size(inputs[k].dtype) == 0) { set_error("input %d: unknown dtype %d", k, inputs[k].dtype); return NL_ERR_INVALID; }
    if (inputs[k].dtype == NL_BOOL || inputs[k].kind != NL_IN_ARRAY) continue;
    if (dtype >= 0 && inputs[k].dtype != dtype) { set_error("input %d: mixed column dtypes in one pipeline", k); return NL_ERR_UNSUPPORTED; }
    dtype = inputs[k].dtype;
  }
  for (int k = 0; k < n_inputs; k++) {
    if (inputs[k].dtype == NL_BOOL) continue;
    if (dtype < 0) dtype = inputs[k].dtype;
    if (inputs[k].dtype != dtype) { set_error("input %d: scalar dtype differs from the pipeline dtype (convert on the host)", k); return NL_ERR_UNSUPPORTED; }
  }
  if (dtype < 0) dtype = NL_I64;
  plan->dtype = dtype;
  plan->n_inputs = n_inputs;
  for (int k = 0; k < n_inputs; k++) plan->in[k] = inputs[k];

  Lower lw;
  lw.nodes = nodes; lw.n_nodes = n_nodes; lw.in = inputs; lw.n_in = n_inputs; lw.plan = plan;
  for (int k = 0; k < NL_MAX_INPUTS; k++) lw.input_temp[k] = -1;
  for (int k = 0; k < NL_MAX_OUTPUTS; k++) lw.out_space[k] = NL_IDX_LOOP;

  int root = n_nodes - 1;
  lw.walk(root);
  if (lw.rc != NL_OK) return lw.rc;
  if (nodes[root].output < 0) { set_error("pipeline root is not materialised"); return NL_ERR_INVALID; }
  lw.cache_budget = NL_MAX_TEMPS - lw.temps_needed(root);
  // A plan that needs no value temporary for its own evaluation stays without one: caching a
  // twice-read column would be its only temporary, which costs the interpreter a resident CTA per SM
  // and every compaction kernel half its rows per thread, while the second read of a row hits L1/L2
  // (and the super-tile compaction kernel reads its rows a second time from L2 by design).
  if (lw.temps_needed(root) == 0) lw.cache_budget = 0;
  if (lw.is_pred(root)) lw.gen_pred(root); else lw.gen_value(root);
  if (lw.rc != NL_OK) return lw.rc;

  // one compaction counter per pipeline: every compacted store has to follow the last filter
  int n_filters = 0, seen = 0;
  for (int i = 0; i < plan->n_steps; i++) if (step_op(plan->step[i]) == S_FILTER) n_filters++;
  for (int i = 0; i < plan->n_steps; i++) {
    uint32_t s = plan->step[i];
    if (step_op(s) == S_FILTER) seen++;
    if ((step_op(s) == S_STORE || step_op(s) == S_PSTORE) && step_b(s) == NL_IDX_COMPACT) {
      plan->flags |= NL_PLAN_HAS_COMPACT;
      if (seen != n_filters) { set_error("a filtered result is materialised between two filters: split the pipeline there"); return NL_ERR_SPLIT_FILTER; }
    }
  }
  // Dead saves: a column parked in a temporary "for later" whose later use found it still in the
  // accumulator.  Dropping them is what lets such plans run on the temporary-free interpreter.
  {
    int w = 0, max_t = 0;
    for (int i = 0; i < plan->n_steps; i++) {
      const uint32_t st = plan->step[i];
      if (step_op(st) == S_SAVE) {
        const uint32_t j = step_a(st);
        bool used = false;
        for (int k = i + 1; k < plan->n_steps && !used; k++) {
          const uint32_t t = plan->step[k], op = step_op(t);
          if (op == S_SAVE && step_a(t) == j) break;                        // overwritten before any read
          if (op == S_LOADT && step_a(t) == j) used = true;
          if ((op == S_BIN || op == S_CMP) && step_b(t) == (SRC_TEMP | j)) used = true;
        }
        if (!used) continue;
        if ((int)j + 1 > max_t) max_t = (int)j + 1;
      }
      plan->step[w++] = st;
    }
    for (int i = w; i < plan->n_steps; i++) plan->step[i] = 0;
    plan->n_steps = w;
    plan->n_temps = max_t;
  }
  return NL_OK;
}

static const char *dtype_name(int d) {
  static const char *n[] = {"bool", "i32", "i64", "f32", "f64", "i8", "i16", "u8", "u16", "u32", "u64"};
  return (d >= 0 && d < NL_N_DTYPES) ? n[d] : "?";
}

int nl_plan_dump(const nl_plan *plan, char *buf, size_t buflen) {
  if (!plan || !buf || buflen == 0) { set_error("nl_plan_dump: null argument"); return NL_ERR_INVALID; }
  size_t off = 0;
  auto put = [&](const char *fmt, ...) {
    if (off >= buflen) return;
    va_list ap;
    va_start(ap, fmt);
    int w = vsnprintf(buf + off, buflen - off, fmt, ap);
    va_end(ap);
    if (w > 0) off += (size_t)w;
    if (off >= buflen) off = buflen - 1;
  };
  static const char *vop[] = {"mul", "add", "sub", "rsub", "div", "rdiv", "floordiv", "rfloordiv"};
  static const char *cop[] = {"ge", "gt", "le", "lt", "eq", "ne"};
  static const char *pop[] = {"and", "or", "xor"};
  static const char *rop[] = {"max", "min", "sum"};
  static const char *space[] = {"loop", "compact", "shard", "where"};
  static const char *kind[] = {"array", "imm", "devscalar"};
  put("plan dtype=%s inputs=%d outputs=%d steps=%d temps=%d flags=%s%s%s%s\n", dtype_name(plan->dtype), plan->n_inputs,
      plan->n_outputs, plan->n_steps, plan->n_temps, (plan->flags & NL_PLAN_HAS_FILTER) ? "F" : "",
      (plan->flags & NL_PLAN_HAS_COMPACT) ? "C" : "", (plan->flags & NL_PLAN_HAS_REDUCE) ? "R" : "",
      (plan->flags & NL_PLAN_HAS_WHERE) ? "W" : "");
  for (int k = 0; k < plan->n_inputs; k++)
    put("  in%d: %s %s\n", k, dtype_name(plan->in[k].dtype), kind[plan->in[k].kind % 3]);
  for (int k = 0; k < plan->n_outputs; k++)
    put("  out%d: %s %s\n", k, dtype_name(plan->out[k].dtype), space[plan->out[k].index_space & 3]);
  for (int i = 0; i < plan->n_steps; i++) {
    uint32_t s = plan->step[i], a = step_a(s), b = step_b(s);
    char src[16];
    if (b & SRC_TEMP) snprintf(src, sizeof src, "t%u", b & 0x7f); else snprintf(src, sizeof src, "in%u", b);
    switch (step_op(s)) {
      case S_LOAD:   put("  %2d: acc = in%u\n", i, a); break;
      case S_LOADT:  put("  %2d: acc = t%u\n", i, a); break;
      case S_SAVE:   put("  %2d: t%u = acc\n", i, a); break;
      case S_BIN:    put("  %2d: acc = %s(acc, %s)\n", i, vop[a & 7], src); break;
      case S_CMP:    put("  %2d: pacc = %s(acc, %s)\n", i, cop[a % 6], src); break;
      case S_PLOAD:  put("  %2d: pacc = in%u\n", i, a); break;
      case S_PLOADT: put("  %2d: pacc = p%u\n", i, a); break;
      case S_PSAVE:  put("  %2d: p%u = pacc\n", i, a); break;
      case S_PBIN:   put("  %2d: pacc = %s(pacc, p%u)\n", i, pop[a % 3], b); break;
      case S_PNOT:   put("  %2d: pacc = not pacc\n", i); break;
      case S_FILTER: put("  %2d: filter (keep &= pacc)\n", i); break;
      case S_STORE:  put("  %2d: out%u[%s] = acc\n", i, a, space[b & 3]); break;
      case S_PSTORE: put("  %2d: out%u[%s] = pacc\n", i, a, space[b & 3]); break;
      case S_REDUCE: put("  %2d: out%u[shard] = %s over acc\n", i, b, rop[a % 3]); break;
      case S_WHERE:  put("  %2d: out%u[i] = acc where pacc\n", i, a); break;
      default:       put("  %2d: ?? %08x\n", i, s); break;
    }
  }
  return (int)off;
}

int nl_partition(int64_t size, int n_chunks, int64_t align, int chunk, int64_t *start_out, int64_t *end_out) {
  if (size < 0 || n_chunks <= 0 || chunk < 0 || chunk >= n_chunks || !start_out || !end_out) {
    set_error("nl_partition: bad argument");
    return NL_ERR_INVALID;
  }
  if (align < 1) align = 1;
  if (align & (align - 1)) { set_error("nl_partition: align must be a power of two"); return NL_ERR_INVALID; }
  // scheduler.cpp:133-139: increment = size / thread_count, the last chunk runs to `size`;
  // interior boundaries are rounded down to `align` so 128-bit accesses stay aligned
  int64_t inc = (size / n_chunks) & ~(align - 1);
  *start_out = inc * chunk;
  *end_out = (chunk == n_chunks - 1) ? size : inc * (chunk + 1);
  return NL_OK;
}

}  // extern "C"
