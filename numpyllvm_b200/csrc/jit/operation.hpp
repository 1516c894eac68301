// operation.hpp — the per-operator plug-in record.
//
// Role of the reference's GenCodeInfo + ThunkOperation (gencode.hpp:15-41, operation.hpp:7-27):
// what an operator needs to tell the engine.  There the record carried LLVM emitter callbacks
// (gencode / initialize / initialize_data / finalize_data); here it carries the op-code the
// kernel library executes — the reduction's partial/finish protocol and the filter's index
// counter are implied by the op-code and live behind the C-ABI — plus, for a full pipeline
// breaker, the host (NumPy) function that stands in for a kernel.
#ifndef NLJIT_OPERATION_H
#define NLJIT_OPERATION_H

#include "pyapi.hpp"
#include <string>

// how an operation interacts with pipeline fusion (the reference's operator_type)
enum class OpClass : int {
  Vectorizable,     // runs inside the fused kernel, element by element
  FullBreaker,      // inputs are materialised, a host function produces the result (sort)
  ParallelBreaker   // ends a pipeline but is still computed by the kernel library (max, min, sum)
};

// host stand-in of a FullBreaker: host ndarray in, new host ndarray out
typedef PyObject *(*HostFunction)(PyObject *);

struct ThunkOperation {
  OpClass klass;
  int arity;             // 1 or 2 operands
  int opcode;            // nl_op run by the fused kernel, or -1 when only `base` exists
  HostFunction base;
  PyObject *operand[2];  // owned references (the reference INCREFs its parameters too, operation.cpp:15)
  std::string name;      // "*", ">=", "[]", "sum", ...
  int refs;              // the owning thunk holds one; every parser node that runs the record holds one, so a
                         // pipeline can still read it after the thunk was marked evaluated mid-evaluation

  bool is_breaker() const { return klass != OpClass::Vectorizable; }
};

ThunkOperation *NewUnaryOperation(PyObject *arg, OpClass klass, int opcode, HostFunction base, const char *name);
ThunkOperation *NewBinaryOperation(PyObject *left, PyObject *right, OpClass klass, int opcode, const char *name);
void RetainOperation(ThunkOperation *operation);
void DestroyOperation(ThunkOperation *operation);   /* drops one reference; frees the record (and its operand references) at zero */

#endif
