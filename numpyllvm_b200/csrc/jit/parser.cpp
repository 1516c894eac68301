// parser.cpp — splits the thunk DAG into pipelines at the blocking operations; everything
// inside one pipeline is fused into a single kernel (reference: parser.cpp:6-198).  The rules
// are the reference's:
//   * an evaluated thunk is an input leaf (parser.cpp:84-90);
//   * an operation is a breaker if it is a full breaker or ANY parameter's operation is a full
//     or parallel breaker; then every unevaluated parameter becomes a child pipeline whose
//     output is this pipeline's input (parser.cpp:116-145);
//   * cardinality comes from the per-operation function (parser.cpp:149-151);
//   * an operation is materialised if its thunk is referenced from outside this computation
//     (`ob_refcnt > 1`, parser.cpp:66) or it is a pipeline root (parser.cpp:154-170).
#include "parser.hpp"
#include "debug_printer.hpp"

#include <cassert>
#include <cstdlib>

static Pipeline *CreatePipeline() {
  Pipeline *pipeline = new Pipeline();
  pipeline->operation = NULL;
  pipeline->parent = NULL;
  pipeline->inputData = new DataBlock();
  pipeline->outputData = new DataBlock();
  pipeline->name = getname("Pipe");
  pipeline->evaluated = false;
  pipeline->inplace_target = NULL;
  return pipeline;
}

static void source_decref(DataSource *s) {
  if (s && --s->refs == 0) {
    Py_XDECREF((PyObject *)s->object);
    storage_decref(s->storage);
    delete s;
  }
}

void DestroyPipeline(Pipeline *pipeline) {
  if (!pipeline) return;
  for (Pipeline *c : pipeline->children) DestroyPipeline(c);
  delete pipeline->operation;
  for (DataElement &e : pipeline->inputData->objects) source_decref(e.source);
  for (DataElement &e : pipeline->outputData->objects) source_decref(e.source);
  delete pipeline->inputData;
  delete pipeline->outputData;
  free(pipeline->name);
  delete pipeline;
}

static void AddChild(Pipeline *parent, Pipeline *child) {
  assert(!child->parent);
  child->parent = parent;
  parent->children.push_back(child);
}

struct ParseState {
  const std::set<PyThunkObject *> *cuts;
  bool failed;
};

static DataSource *source_for_thunk(PyThunkObject *thunk) {
  Py_INCREF(thunk);
  if (thunk->storage) storage_incref(thunk->storage);
  return new DataSource(thunk, thunk->storage, thunk->cardinality, thunk->type);
}

static Operation *OperationFromArgs(PyThunkObject *thunk, ThunkOperation *operation, Operation **operations, int op_count,
                                    std::set<PyThunkObject *> &stored) {
  PyThunkObject *result_object = NULL;
  if (Py_REFCNT(thunk) > 1 && !stored.count(thunk)) {
    /* the thunk is not only used in this calculation but elsewhere too, so we store the result */
    result_object = thunk;
  }
  switch (op_count) {
    case 1: return new UnaryOperation(result_object, thunk, operation, operations[0]);
    case 2: return new BinaryOperation(result_object, thunk, operation, operations[0], operations[1]);
  }
  return NULL;
}

static Operation *ParseOperationRecursive(PyThunkObject *thunk, Pipeline *current, DataSource *output_source,
                                          std::set<PyThunkObject *> &stored, ParseState &ps) {
  if (thunk->evaluated) {
    // already evaluated: an endpoint
    Operation *op = new ObjectOperation(thunk);
    current->inputData->objects.push_back(DataElement(op, source_for_thunk(thunk)));
    return op;
  }
  // not evaluated: there must be an operation that says how to compute it
  ThunkOperation *toperation = thunk->operation;
  if (!toperation || (int)toperation->gencode.type > 2 || (int)toperation->gencode.type < 1) {
    PyErr_SetString(PyExc_ValueError, "thunk has neither data nor an operation");
    ps.failed = true;
    return new ObjectOperation(thunk);
  }
  const int arity = (int)toperation->gencode.type;
  Operation *operations[OPTYPE_MAX_OPERATIONS] = {NULL, NULL, NULL};
  ssize_t child_cardinalities[OPTYPE_MAX_OPERATIONS] = {0, 0, 0};

  // if any of the children is a pipeline breaker, each child becomes a separate pipeline
  // (parser.cpp:99-129: sort(a*b*c) * (d*e*f) makes three pipelines)
  bool breaker = toperation->type == optype_fullbreaker;
  for (int i = 0; i < arity && !breaker; i++) {
    PyObject *param = toperation->gencode.parameter[i];
    if (PyThunk_Check(param) && ((PyThunkObject *)param)->operation != NULL &&
        (((PyThunkObject *)param)->operation->type == optype_fullbreaker ||
         ((PyThunkObject *)param)->operation->type == optype_parallelbreaker)) {
      breaker = true;
    }
  }
  for (int i = 0; i < arity; i++) {
    PyThunkObject *thunk_child = (PyThunkObject *)toperation->gencode.parameter[i];
    const bool cut = ps.cuts && ps.cuts->count(thunk_child);
    if ((breaker || cut) && !thunk_child->evaluated) {
      Pipeline *child = CreatePipeline();
      DataSource *source = new DataSource(NULL, NULL, 0, 0);
      std::set<PyThunkObject *> child_stored;
      // we materialise after a breaking operation
      child->operation = ParseOperationRecursive(thunk_child, child, source, child_stored, ps);
      // placeholder that receives the child pipeline's result as input
      operations[i] = new PipelineOperation(child, thunk_child);
      source->refs++;
      current->inputData->objects.push_back(DataElement(operations[i], source));
      source_decref(source);   // the parse-time reference; output list + input list hold theirs
      AddChild(current, child);
    } else {
      operations[i] = ParseOperationRecursive(thunk_child, current, NULL, stored, ps);
    }
    child_cardinalities[i] = thunk_child->cardinality;
  }
  thunk->cardinality = thunk->cardinality_func(child_cardinalities);
  if (thunk->cardinality < 0 && !ps.failed) {
    PyErr_SetString(PyExc_TypeError, "Incompatible cardinality.");
    ps.failed = true;
  }

  Operation *op = OperationFromArgs(thunk, toperation, operations, arity, stored);
  if (op->result_object || output_source) {
    op->materialize = true;
    stored.insert(thunk);
    if (output_source) {
      output_source->object = thunk;
      output_source->storage = NULL;
      output_source->size = thunk->cardinality;
      output_source->type = thunk->type;
      output_source->refs++;
      Py_INCREF(thunk);
    } else {
      output_source = new DataSource(thunk, NULL, thunk->cardinality, thunk->type);
      Py_INCREF(thunk);
    }
    current->outputData->objects.push_back(DataElement(op, output_source));
  }
  return op;
}

ThunkOperation *GetThunkOperation(Operation *op) {
  switch (op->Type()) {
    case OPTYPE_unop: return ((UnaryOperation *)op)->operation;
    case OPTYPE_binop: return ((BinaryOperation *)op)->operation;
    default: return NULL;
  }
}

Pipeline *ParsePipeline(PyThunkObject *thunk, const std::set<PyThunkObject *> *cuts) {
  Pipeline *pipeline = CreatePipeline();
  ParseState ps = {cuts, false};
  std::set<PyThunkObject *> stored;
  // the root is always materialised (parser.cpp:154: a pipeline root has an output source)
  DataSource *root = new DataSource(NULL, NULL, 0, 0);
  pipeline->operation = ParseOperationRecursive(thunk, pipeline, root, stored, ps);
  source_decref(root);
  if (ps.failed) {
    DestroyPipeline(pipeline);
    return NULL;
  }
  return pipeline;
}
