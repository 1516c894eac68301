// parser.hpp — the thunk DAG cut into a tree of fused pipelines.
//
// Same job as the reference's parser (parser.hpp:14-141): an operation tree per pipeline with
// its input and output columns.  The reference models nodes as a class hierarchy
// (Object/Pipeline/Nullary/Unary/BinaryOperation) and columns as DataBlock/DataElement; here a
// node is one tagged struct and a pipeline simply owns two binding lists.
#ifndef NLJIT_PARSER_H
#define NLJIT_PARSER_H

#include "thunk.hpp"
#include <map>
#include <set>
#include <string>
#include <vector>

struct Pipeline;

// one node of a pipeline's operation tree
struct Operation {
  enum Kind {
    Column,        // an evaluated thunk read as an input column
    ChildResult,   // the output of a child pipeline read as an input column
    Apply1,        // unary operator
    Apply2         // binary operator
  };
  Kind kind;
  PyThunkObject *thunk;          // the thunk this node computes / reads
  PyThunkObject *result_object;  // non-NULL: the value is also wanted outside this computation
  bool materialize;              // the node's value is stored to an output column
  ThunkOperation *op;            // Apply1/Apply2: the operator record
  Operation *lhs, *rhs;          // owned
  Pipeline *child;               // ChildResult: the pipeline that produces it

  // A node owns a reference to its thunk and to the operator record it runs: a child pipeline that
  // finishes marks its thunks evaluated, which drops their operations (and with them whole operand
  // sub-DAGs); later pipelines of the same tree still read both.  (The reference keeps raw pointers.)
  Operation(Kind k, PyThunkObject *t, ThunkOperation *o = NULL)
      : kind(k), thunk(t), result_object(NULL), materialize(false), op(o), lhs(NULL), rhs(NULL), child(NULL) {
    Py_XINCREF((PyObject *)thunk);
    RetainOperation(op);
  }
  ~Operation() {
    delete lhs;
    delete rhs;
    DestroyOperation(op);
    Py_XDECREF((PyObject *)thunk);
  }
  bool is_leaf() const { return kind == Column || kind == ChildResult; }
};

// a column a pipeline reads or writes: the owning thunk, its storage once it exists, its size
// (the upper bound for filter results) and NumPy type.  Shared between a child's output list
// and its parent's input list.
struct DataSource {
  PyThunkObject *object;
  Storage *storage;
  ssize_t size;
  int type;
  int refs;
  DataSource(PyThunkObject *object, Storage *storage, ssize_t size, int type)
      : object(object), storage(storage), size(size), type(type), refs(1) {}
};

struct Binding {
  Operation *operation;
  DataSource *source;
};

struct Pipeline {
  Operation *root;
  Pipeline *parent;
  std::vector<Pipeline *> children;
  std::vector<Binding> inputs;    // columns read (evaluated thunks and child results)
  std::vector<Binding> outputs;   // columns written (materialised nodes)
  std::string name;
  bool evaluated;
  PyThunkObject *inplace_target;  // masked assignment: the root's output aliases this thunk's storage
};

/* `cuts`: thunks the plan optimiser wants materialised by their own pipeline (an expression too
 * large for one fused kernel); treated like pipeline breakers. */
Pipeline *ParsePipeline(PyThunkObject *thunk, const std::set<PyThunkObject *> *cuts = NULL);
void DestroyPipeline(Pipeline *pipeline);

#endif
