// thunk.hpp — the lazy array object (reference: thunk.hpp:23-66).
#ifndef NLJIT_THUNK_H
#define NLJIT_THUNK_H

#include "operation.hpp"
#include "storage.hpp"
#include <vector>

typedef ssize_t (*cardinality_function)(ssize_t *inputs);

ssize_t default_cardinality_function(ssize_t *inputs);
ssize_t default_binary_cardinality_function(ssize_t *inputs);

typedef enum {
  cardinality_unknown = 255,
  cardinality_exact = 1,
  cardinality_upper_bound = 2
} cardinality_type_t;

struct PyThunkObject {
  PyObject_HEAD
  Storage *storage;          /* device shards (or a host scalar); NULL until evaluated / uploaded */
  PyObject *host;            /* host ndarray: the wrapped copy when no device is present, else a
                                cache filled by asnumpyarray()/str() */
  bool evaluated;
  ThunkOperation *operation; /* how to compute this thunk; released once evaluated */
  ssize_t cardinality;
  cardinality_type_t cardinality_type;
  cardinality_function cardinality_func;
  int type;                  /* NumPy type number */
  char *name;
  std::vector<PyThunkObject *> *children;  /* thunks computed from this one (borrowed; each child
                                              removes itself on dealloc) */
};

extern PyTypeObject PyThunk_Type;

#define PyThunk_CheckExact(op) (Py_TYPE(op) == &PyThunk_Type)
#define PyThunk_Check(op) PyObject_TypeCheck((PyObject *)(op), &PyThunk_Type)

PyObject *PyThunk_FromArray(PyObject *unused, PyObject *input);
PyObject *PyThunk_FromOperation(ThunkOperation *operation, cardinality_function cardinality_func,
                                cardinality_type_t cardinality_tpe, int type);
/* coerce an operand (thunk, ndarray, Python scalar) to a thunk; new reference */
PyObject *PyThunk_Coerce(PyObject *obj);

int PyThunk_Init(void);
void PyThunk_EvaluateChildren(PyThunkObject *thunk);
PyObject *PyThunk_Evaluate(PyThunkObject *thunk);
bool PyThunk_IsEvaluated(PyThunkObject *thunk);
PyObject *PyThunk_AsArray(PyObject *thunk);        /* borrowed reference to the host ndarray */
/* make sure an evaluated thunk's data is on the device (deferred upload); 0 / -1 */
int PyThunk_EnsureDevice(PyThunkObject *thunk);
void PyThunk_InvalidateHost(PyThunkObject *thunk);

extern PyObject *PyThunk_LazyRichCompare(PyObject *self, PyObject *other, int cmp_op);
extern PyNumberMethods thunk_as_number;
extern PySequenceMethods thunk_as_sequence;
extern PyMappingMethods thunk_as_mapping;
extern struct PyMethodDef thunk_methods[];

char *getname(const char *base);

#endif
