// thunk.hpp — the lazy array object of module `jit`.
//
// Plays the part of the reference's PyThunkObject (thunk.hpp:23-43): an evaluated thunk owns
// data, an unevaluated one owns the operation that will compute it.  The data is device-resident
// shard storage (storage.hpp) with a host ndarray cached on demand; dependents are tracked so an
// in-place assignment can force them first (thunk.cpp:81-88).
#ifndef NLJIT_THUNK_H
#define NLJIT_THUNK_H

#include "operation.hpp"
#include "storage.hpp"
#include <vector>

typedef ssize_t (*CardinalityFn)(ssize_t *operand_sizes);

enum class CardinalityKind : int { Unknown, Exact, UpperBound };

// how many elements a thunk has / will have
struct Cardinality {
  ssize_t n;             // -1 until known
  CardinalityKind kind;  // a filter only knows an upper bound before it runs
  CardinalityFn fn;      // result size from the operand sizes (unevaluated thunks)
};

ssize_t cardinality_same_as_operand(ssize_t *sizes);
ssize_t cardinality_broadcast(ssize_t *sizes);   // size-1 operands broadcast; -1 on mismatch

struct PyThunkObject {
  PyObject_HEAD
  bool evaluated;
  int npy_type;                               // NumPy type number of the elements
  Cardinality card;
  ThunkOperation *operation;                  // how to compute this thunk; released once evaluated
  Storage *storage;                           // device shards or a host scalar; NULL until evaluated / uploaded
  PyObject *host;                             // host ndarray: the wrapped copy when no device exists,
                                              // otherwise a cache filled by asnumpyarray() / str()
  std::vector<PyThunkObject *> *dependents;   // thunks computed from this one (borrowed; each removes
                                              // itself when it is evaluated or destroyed)
  char *name;
};

extern PyTypeObject PyThunk_Type;

#define PyThunk_Check(op) PyObject_TypeCheck((PyObject *)(op), &PyThunk_Type)

PyObject *PyThunk_FromArray(PyObject *unused, PyObject *input);
PyObject *PyThunk_FromOperation(ThunkOperation *operation, CardinalityFn fn, CardinalityKind kind, int npy_type);
PyObject *PyThunk_Coerce(PyObject *obj);           /* thunk, ndarray or Python scalar -> new reference to a thunk */
PyObject *PyThunk_AsType(PyThunkObject *thunk, int npy_type);   /* new reference: the thunk itself or a (lazy) conversion */

int PyThunk_Init(void);
PyObject *PyThunk_Evaluate(PyThunkObject *thunk);
bool PyThunk_IsEvaluated(PyThunkObject *thunk);
PyObject *PyThunk_AsArray(PyObject *thunk);        /* borrowed reference to the host ndarray */
int PyThunk_EnsureDevice(PyThunkObject *thunk);    /* deferred upload of a host-only thunk; 0 / -1 */
void PyThunk_InvalidateHost(PyThunkObject *thunk);

extern PyObject *PyThunk_LazyRichCompare(PyObject *self, PyObject *other, int cmp_op);
extern PyNumberMethods thunk_as_number;
extern PySequenceMethods thunk_as_sequence;
extern PyMappingMethods thunk_as_mapping;
extern struct PyMethodDef thunk_methods[];

char *getname(const char *base);

#endif
