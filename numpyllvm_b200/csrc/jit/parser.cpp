// parser.cpp — cuts the thunk DAG into pipelines at the blocking operations; everything inside
// one pipeline is fused into a single kernel.  The RULES are the reference's (parser.cpp:81-173):
//   * an evaluated thunk is an input column;
//   * an operation "breaks" if it is a full breaker or if ANY operand is computed by a full or
//     parallel breaker; every unevaluated operand of a breaking operation becomes a child
//     pipeline whose output is this pipeline's input (so `sort(a*b*c) * (d*e*f)` is three
//     pipelines, parser.cpp:99-115);
//   * a thunk's cardinality comes from its operation's cardinality function;
//   * a node is materialised when its thunk is referenced outside this computation
//     (`ob_refcnt > 1`, parser.cpp:66) or it is a pipeline root.
// The data structures are this engine's own (parser.hpp).
#include "parser.hpp"
#include "debug_printer.hpp"

#include <cassert>
#include <cstdlib>

namespace {

struct Parse {
  const std::set<PyThunkObject *> *cuts;   // extra breakers requested by the plan optimiser
  bool failed = false;
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

// `sink`: when non-NULL the node is a pipeline root and its result goes to this column
Operation *parse_node(PyThunkObject *thunk, Pipeline *current, DataSource *sink,
                      std::set<PyThunkObject *> &stored, Parse &ps) {
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

  bool breaks = top->klass == OpClass::FullBreaker;
  for (int i = 0; i < top->arity && !breaks; i++) breaks = computed_by_breaker(top->operand[i]);

  Operation *node = new Operation(top->arity == 1 ? Operation::Apply1 : Operation::Apply2, thunk);
  node->op = top;
  ssize_t sizes[2] = {0, 0};
  for (int i = 0; i < top->arity; i++) {
    PyThunkObject *operand = (PyThunkObject *)top->operand[i];
    Operation *sub;
    const bool cut = ps.cuts && ps.cuts->count(operand);
    if ((breaks || cut) && !operand->evaluated) {
      // the operand gets a pipeline of its own; here it is read back as a column
      Pipeline *child = new_pipeline();
      DataSource *col = new DataSource(NULL, NULL, 0, 0);
      std::set<PyThunkObject *> child_stored;
      child->root = parse_node(operand, child, col, child_stored, ps);
      child->parent = current;
      current->children.push_back(child);
      sub = new Operation(Operation::ChildResult, operand);
      sub->child = child;
      col->refs++;
      current->inputs.push_back(Binding{sub, col});
      release(col);              // drop the construction reference; both lists hold their own
    } else {
      sub = parse_node(operand, current, NULL, stored, ps);
    }
    (i == 0 ? node->lhs : node->rhs) = sub;
    sizes[i] = operand->card.n;
  }
  thunk->card.n = thunk->card.fn(sizes);
  if (thunk->card.n < 0 && !ps.failed) {
    PyErr_SetString(PyExc_TypeError, "Incompatible cardinality.");
    ps.failed = true;
  }

  // somebody besides this computation holds the thunk: keep its value
  if (Py_REFCNT(thunk) > 1 && !stored.count(thunk)) node->result_object = thunk;
  if (node->result_object || sink) {
    node->materialize = true;
    stored.insert(thunk);
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
  std::set<PyThunkObject *> stored;
  DataSource *result = new DataSource(NULL, NULL, 0, 0);    // the root is always materialised
  pipeline->root = parse_node(thunk, pipeline, result, stored, ps);
  release(result);
  if (ps.failed) {
    DestroyPipeline(pipeline);
    return NULL;
  }
  return pipeline;
}
