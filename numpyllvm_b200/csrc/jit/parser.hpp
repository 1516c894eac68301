// parser.hpp — DAG of thunks -> tree of fused pipelines (reference: parser.hpp:14-141).
#ifndef NLJIT_PARSER_H
#define NLJIT_PARSER_H

#include "gencode.hpp"
#include "thunk.hpp"
#include <set>
#include <vector>

class Pipeline;

enum operation_type {
  OPTYPE_nullop = 0,
  OPTYPE_unop = 1,
  OPTYPE_binop = 2,
  OPTYPE_pipeline = 100,
  OPTYPE_obj = 101,
  OPTYPE_unknown = 255
};

#define OPTYPE_MAX_OPERATIONS 3

class Operation {
public:
  /* Where the result of this operation is stored; NULL for an unnecessary intermediate
   * (parser.hpp:27-36).  `thunk` is the node the operation computes, materialised or not. */
  PyThunkObject *result_object;
  PyThunkObject *thunk;
  bool materialize;
  Operation(PyThunkObject *result, PyThunkObject *thunk) : result_object(result), thunk(thunk), materialize(false) {}
  virtual ~Operation() {}
  virtual operation_type Type() { return OPTYPE_unknown; }
};

/* input data from an evaluated thunk */
class ObjectOperation : public Operation {
public:
  ObjectOperation(PyThunkObject *thunk) : Operation(NULL, thunk) {}
  virtual operation_type Type() { return OPTYPE_obj; }
};

/* input data that is the result of a child pipeline */
class PipelineOperation : public Operation {
public:
  Pipeline *pipeline;
  PipelineOperation(Pipeline *pipeline, PyThunkObject *thunk) : Operation(NULL, thunk), pipeline(pipeline) {}
  virtual operation_type Type() { return OPTYPE_pipeline; }
};

class UnaryOperation : public Operation {
public:
  ThunkOperation *operation;
  Operation *LHS;
  UnaryOperation(PyThunkObject *result_object, PyThunkObject *thunk, ThunkOperation *operation, Operation *LHS)
      : Operation(result_object, thunk), operation(operation), LHS(LHS) {}
  virtual ~UnaryOperation() { delete LHS; }
  virtual operation_type Type() { return OPTYPE_unop; }
};

class BinaryOperation : public Operation {
public:
  ThunkOperation *operation;
  Operation *LHS, *RHS;
  BinaryOperation(PyThunkObject *result_object, PyThunkObject *thunk, ThunkOperation *operation, Operation *LHS, Operation *RHS)
      : Operation(result_object, thunk), operation(operation), LHS(LHS), RHS(RHS) {}
  virtual ~BinaryOperation() { delete LHS; delete RHS; }
  virtual operation_type Type() { return OPTYPE_binop; }
};

ThunkOperation *GetThunkOperation(Operation *op);

/* A column a pipeline reads or writes (parser.hpp:94-103): the owning thunk, its storage once
 * it exists, its size (the upper bound for filter results) and dtype. */
class DataSource {
public:
  PyThunkObject *object;
  Storage *storage;
  ssize_t size;
  int type;
  int refs;
  DataSource(PyThunkObject *object, Storage *storage, ssize_t size, int type)
      : object(object), storage(storage), size(size), type(type), refs(1) {}
};

class DataElement {
public:
  Operation *operation;
  DataSource *source;
  DataElement(Operation *op, DataSource *source) : operation(op), source(source) {}
};

class DataBlock {
public:
  std::vector<DataElement> objects;
};

class Pipeline {
public:
  Operation *operation;
  Pipeline *parent;
  std::vector<Pipeline *> children;
  DataBlock *inputData;
  DataBlock *outputData;
  char *name;
  bool evaluated;
  PyThunkObject *inplace_target;   /* masked assignment: the root's output aliases this thunk's storage */
};

/* `cuts`: thunks the plan optimiser wants materialised by their own pipeline (an expression too
 * large for one fused kernel); treated like pipeline breakers. */
Pipeline *ParsePipeline(PyThunkObject *thunk, const std::set<PyThunkObject *> *cuts = NULL);
void DestroyPipeline(Pipeline *pipeline);

#endif
