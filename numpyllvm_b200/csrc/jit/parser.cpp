// parser.cpp — cuts the thunk DAG into pipelines at the blocking operations; everything inside
// one pipeline is fused into a single kernel.  The RULES are the reference's (parser.cpp:81-173):
//   * an evaluated thunk is an input column;
//   * an operation "breaks" if it is a full breaker or if ANY operand is computed by a full or
//     parallel breaker; every unevaluated operand of a breaking operation becomes a child
//     pipeline whose output is this pipeline's input (so `sort(a*b*c) * (d*e*f)` is three
//     pipelines, parser.cpp:99-115);
//   * a thunk's cardinality comes from its operation's cardinality function;
//   * a node is materialised when its thunk is referenced outside this computation
//     (the reference tests `ob_refcnt > 1`, parser.cpp:66; here: more references than the DAG's own
//     operand slots account for) or it is a pipeline root;
//   * (new) a thunk that any consumer needs as a child pipeline gets ONE pipeline however often the
//     DAG reaches it, and every consumer reads that result as a column.
// The data structures are this engine's own (parser.hpp).
#include "parser.hpp"
#include "debug_printer.hpp"

#include <cassert>
#include <cstdlib>

namespace {

struct Parse {
  const std::set<PyThunkObject *> *cuts;   // extra breakers requested by the plan optimiser
  bool failed = false;
  // Facts about the DAG collected BEFORE any node is built (the nodes themselves take references):
  std::set<PyThunkObject *> seen;
  std::map<PyThunkObject *, Py_ssize_t> internal;   // operand slots inside this DAG that reference the thunk
  std::set<PyThunkObject *> held;                    // referenced from outside this DAG: keep the value
  std::set<PyThunkObject *> own_pipeline;            // some consumer needs it materialised by a pipeline of its own
  // one child pipeline per thunk and parse: a thunk reached twice through breakers (`(a-m)*(b-m)`,
  // `x.sort()*x`) is computed once and read back as a column by every consumer
  std::map<PyThunkObject *, std::pair<Pipeline *, DataSource *>> child_of;
  std::set<PyThunkObject *> stored;                  // thunks already materialised by some node of this parse
};

Pipeline *new_pipeline() {
  Pipeline *p = new Pipeline();
  p->root = NULL;
  p->parent = NULL;
  p->evaluated = false;
  p->inplace_target = NULL;
  char *n = getname("Pipe");
  p->name = n;
  free(n);
  return p;
}

void release(DataSource *s) {
  if (s && --s->refs == 0) {
    Py_XDECREF((PyObject *)s->object);
    storage_decref(s->storage);
    delete s;
  }
}

DataSource *column_of(PyThunkObject *thunk) {
  Py_INCREF(thunk);
  if (thunk->storage) storage_incref(thunk->storage);
  return new DataSource(thunk, thunk->storage, thunk->card.n, thunk->npy_type);
}

bool computed_by_breaker(PyObject *operand) {
  if (!PyThunk_Check(operand)) return false;
  ThunkOperation *op = ((PyThunkObject *)operand)->operation;
  return op != NULL && op->is_breaker();
}

bool operation_breaks(ThunkOperation *top) {
  bool breaks = top->klass == OpClass::FullBreaker;
  for (int i = 0; i < top->arity && !breaks; i++) breaks = computed_by_breaker(top->operand[i]);
  return breaks;
}

// First pass over the DAG.  For every unevaluated thunk: how many operand slots of this DAG point
// at it (so "somebody else holds it" is `refcount > internal references` — the reference's
// `ob_refcnt > 1`, parser.cpp:66, made independent of how many times the DAG itself or the
// interpreter's call sequence references the object) and whether any consumer breaks on it.
void survey(PyThunkObject *thunk, Parse &ps) {
  if (thunk->evaluated || !ps.seen.insert(thunk).second) return;
  ThunkOperation *top = thunk->operation;
  if (!top || top->arity < 1 || top->arity > 2) return;
  const bool breaks = operation_breaks(top);
  for (int i = 0; i < top->arity; i++) {
    if (!top->operand[i] || !PyThunk_Check(top->operand[i])) continue;
    PyThunkObject *operand = (PyThunkObject *)top->operand[i];
    ps.internal[operand]++;
    if (!operand->evaluated && (breaks || (ps.cuts && ps.cuts->count(operand)))) ps.own_pipeline.insert(operand);
    survey(operand, ps);
  }
}

// `sink`: when non-NULL the node is a pipeline root and its result goes to this column
Operation *parse_node(PyThunkObject *thunk, Pipeline *current, DataSource *sink, Parse &ps) {
  if (thunk->evaluated) {
    Operation *leaf = new Operation(Operation::Column, thunk);
    current->inputs.push_back(Binding{leaf, column_of(thunk)});
    return leaf;
  }
  ThunkOperation *top = thunk->operation;
  if (!top || top->arity < 1 || top->arity > 2) {
    PyErr_SetString(PyExc_ValueError, "thunk has neither data nor an operation");
    ps.failed = true;
    return new Operation(Operation::Column, thunk);
  }

  Operation *node = new Operation(top->arity == 1 ? Operation::Apply1 : Operation::Apply2, thunk, top);
  ssize_t sizes[2] = {0, 0};
  for (int i = 0; i < top->arity; i++) {
    PyThunkObject *operand = (PyThunkObject *)top->operand[i];
    Operation *sub;
    if (!operand->evaluated && ps.own_pipeline.count(operand)) {
      // the operand gets a pipeline of its own (once); here it is read back as a column
      Pipeline *child;
      DataSource *col;
      auto known = ps.child_of.find(operand);
      if (known == ps.child_of.end()) {
        child = new_pipeline();
        col = new DataSource(NULL, NULL, 0, 0);
        child->root = parse_node(operand, child, col, ps);
        child->parent = current;
        current->children.push_back(child);      // parse order is execution order: the first consumer's tree runs it
        ps.child_of[operand] = std::make_pair(child, col);
        col->refs++;
        release(col);                            // drop the construction reference; the child's output list keeps it alive
      } else {
        child = known->second.first;
        col = known->second.second;
        col->refs++;
      }
      sub = new Operation(Operation::ChildResult, operand);
      sub->child = child;
      current->inputs.push_back(Binding{sub, col});
    } else {
      sub = parse_node(operand, current, NULL, ps);
    }
    (i == 0 ? node->lhs : node->rhs) = sub;
    sizes[i] = operand->card.n;
  }
  thunk->card.n = thunk->card.fn(sizes);
  if (thunk->card.n < 0 && !ps.failed) {
    PyErr_SetString(PyExc_TypeError, "Incompatible cardinality.");
    ps.failed = true;
  }

  // somebody besides this computation holds the thunk: keep its value (once per parse)
  if (ps.held.count(thunk) && !ps.stored.count(thunk)) node->result_object = thunk;
  if (node->result_object || sink) {
    node->materialize = true;
    ps.stored.insert(thunk);
    DataSource *out = sink ? sink : new DataSource(NULL, NULL, 0, 0);
    out->object = thunk;
    out->size = thunk->card.n;
    out->type = thunk->npy_type;
    Py_INCREF(thunk);
    if (sink) out->refs++;
    current->outputs.push_back(Binding{node, out});
  }
  return node;
}

}  // namespace

void DestroyPipeline(Pipeline *pipeline) {
  if (!pipeline) return;
  for (Pipeline *c : pipeline->children) DestroyPipeline(c);
  delete pipeline->root;
  for (Binding &b : pipeline->inputs) release(b.source);
  for (Binding &b : pipeline->outputs) release(b.source);
  delete pipeline;
}

Pipeline *ParsePipeline(PyThunkObject *thunk, const std::set<PyThunkObject *> *cuts) {
  Pipeline *pipeline = new_pipeline();
  Parse ps;
  ps.cuts = cuts;
  survey(thunk, ps);
  for (PyThunkObject *t : ps.seen) {
    auto it = ps.internal.find(t);
    if (Py_REFCNT(t) > (it == ps.internal.end() ? 0 : it->second)) ps.held.insert(t);
  }
  DataSource *result = new DataSource(NULL, NULL, 0, 0);    // the root is always materialised
  pipeline->root = parse_node(thunk, pipeline, result, ps);
  release(result);
  if (ps.failed) {
    DestroyPipeline(pipeline);
    return NULL;
  }
  return pipeline;
}
