// thunk.cpp — the thunk object and its type (reference: thunk.cpp:9-195).
//
// Same surface as the reference: jit.thunk(ndarray) wraps (and copies) an array, operators
// build a lazy DAG, evaluate()/str()/asnumpyarray() force it.  Differences that are
// deliberate: the copy made on wrapping IS the host->device upload, children are tracked
// without dangling pointers, and a thunk drops its operation (and so its operands' buffers)
// once it is evaluated.
#include "thunk.hpp"
#include "parser.hpp"
#include "optimizer.hpp"
#include "scheduler.hpp"
#include "debug_printer.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

ssize_t cardinality_same_as_operand(ssize_t *inputs) { return inputs[0]; }

ssize_t cardinality_broadcast(ssize_t *inputs) {
  // if one input is a constant the other size is used (thunk.cpp:14-23)
  if (inputs[0] == 1) return inputs[1];
  if (inputs[1] == 1) return inputs[0];
  if (inputs[0] != inputs[1]) return -1;
  return inputs[0];
}

static PyThunkObject *thunk_alloc() {
  PyThunkObject *thunk = (PyThunkObject *)PyThunk_Type.tp_alloc(&PyThunk_Type, 0);
  if (!thunk) return NULL;
  thunk->storage = NULL;
  thunk->host = NULL;
  thunk->evaluated = false;
  thunk->operation = NULL;
  thunk->card.n = -1;
  thunk->card.kind = CardinalityKind::Unknown;
  thunk->card.fn = NULL;
  thunk->npy_type = NPY_NOTYPE;
  thunk->name = NULL;
  thunk->dependents = new std::vector<PyThunkObject *>();
  return thunk;
}

PyObject *PyThunk_FromArray(PyObject *unused, PyObject *input) {
  (void)unused;
  // the reference copies with NPY_ARRAY_ENSURECOPY (thunk.cpp:30); here the "copy" is the upload
  PyObject *arr = PyArray_FromAny(input, NULL, 0, 0, NPY_ARRAY_C_CONTIGUOUS | NPY_ARRAY_ALIGNED, NULL);
  if (arr == NULL || !PyArray_Check(arr)) {
    Py_XDECREF(arr);
    PyErr_SetString(PyExc_TypeError, "Expected a NumPy array as parameter.");
    return NULL;
  }
  PyArrayObject *a = (PyArrayObject *)arr;
  const int nl_type = nljit_npy_to_nl(PyArray_TYPE(a));
  if (nl_type < 0) {
    PyErr_Format(PyExc_TypeError, "thunk: dtype %R is outside the kernel library (bool, int8..int64, uint8..uint64, float32, float64)",
                 (PyObject *)PyArray_DESCR(a));
    Py_DECREF(arr);
    return NULL;
  }
  PyThunkObject *thunk = thunk_alloc();
  if (!thunk) { Py_DECREF(arr); return NULL; }
  thunk->evaluated = true;
  thunk->card.n = (ssize_t)PyArray_SIZE(a);
  thunk->card.kind = CardinalityKind::Exact;
  thunk->npy_type = PyArray_TYPE(a);
  thunk->name = getname("A");
  if (thunk->card.n == 1) {
    // scalars become size-1 thunks (thunk_as_number.cpp:52-59) and never need the device
    Storage *s = storage_new(nl_type, 1);
    s->host_scalar = true;
    memcpy(&s->scalar_bits, PyArray_DATA(a), nl_dtype_size(nl_type));
    thunk->storage = s;
    Py_DECREF(arr);
    return (PyObject *)thunk;
  }
  if (nljit_ensure_runtime() == 0) {
    thunk->storage = storage_from_host(a);
    Py_DECREF(arr);
    if (!thunk->storage) { Py_DECREF(thunk); return NULL; }
  } else {
    // no device in this process: keep a private host copy so graphs can still be built and
    // inspected; evaluating anything will raise the runtime error (there is no CPU fallback)
    PyErr_Clear();
    thunk->host = PyArray_NewCopy(a, NPY_CORDER);
    Py_DECREF(arr);
    if (!thunk->host) { Py_DECREF(thunk); return NULL; }
  }
  return (PyObject *)thunk;
}

PyObject *PyThunk_Coerce(PyObject *obj) {
  if (PyThunk_Check(obj)) { Py_INCREF(obj); return obj; }
  PyObject *arr = PyArray_FromAny(obj, NULL, 0, 0, 0, NULL);
  if (arr == NULL) {
    PyErr_Clear();
    PyErr_SetString(PyExc_TypeError, "could not coerse operand to thunk");   // thunk_as_number.cpp:55
    return NULL;
  }
  PyObject *t = PyThunk_FromArray(NULL, arr);
  Py_DECREF(arr);
  return t;
}

// astype(): conversion to another element type.  An evaluated host scalar is converted here; anything else gets a
// lazy full breaker whose work is one nl_cast pass per shard (scheduler.cpp: run_base_function) — the reference's
// only cast is the signed integer cast of its materialise store (compiler.cpp:256-260).
PyObject *PyThunk_AsType(PyThunkObject *thunk, int npy_type) {
  if (thunk->npy_type == npy_type) { Py_INCREF(thunk); return (PyObject *)thunk; }
  if (nljit_npy_to_nl(npy_type) < 0) {
    PyErr_SetString(PyExc_TypeError, "astype: dtype is outside the kernel library");
    return NULL;
  }
  if (npy_type == NPY_BOOL || thunk->npy_type == NPY_BOOL) {
    PyErr_SetString(PyExc_TypeError, "astype: conversions from or to bool are not supported (compare with 0 instead)");
    return NULL;
  }
  if (thunk->evaluated && thunk->storage && thunk->storage->host_scalar) {
    Storage *s = storage_new(nljit_npy_to_nl(npy_type), 1);
    s->host_scalar = true;
    s->scalar_bits = nljit_convert_scalar(thunk->storage->scalar_bits, thunk->storage->dtype, s->dtype);
    PyThunkObject *out = thunk_alloc();
    if (!out) { storage_decref(s); return NULL; }
    out->evaluated = true;
    out->card.n = 1;
    out->card.kind = CardinalityKind::Exact;
    out->npy_type = npy_type;
    out->name = getname("A");
    out->storage = s;
    return (PyObject *)out;
  }
  ThunkOperation *op = NewUnaryOperation((PyObject *)thunk, OpClass::FullBreaker, -1, NULL, "astype");
  if (!op) return PyErr_NoMemory();
  return PyThunk_FromOperation(op, cardinality_same_as_operand, thunk->card.kind, npy_type);
}

PyObject *PyThunk_FromOperation(ThunkOperation *operation, CardinalityFn cardinality_func,
                                CardinalityKind cardinality_tpe, int type) {
  PyThunkObject *thunk = thunk_alloc();
  if (!thunk) return NULL;
  thunk->operation = operation;
  thunk->card.kind = cardinality_tpe;
  thunk->card.fn = cardinality_func;
  thunk->npy_type = type;
  thunk->name = getname("operation");
  // register as a dependent of each parameter (thunk.cpp:65-70): assignment to a parameter
  // must evaluate us first
  for (int i = 0; i < (int)operation->arity; i++) {
    PyObject *parameter = operation->operand[i];
    if (PyThunk_Check(parameter)) ((PyThunkObject *)parameter)->dependents->push_back(thunk);
  }
  return (PyObject *)thunk;
}

static void release_operation(PyThunkObject *thunk) {
  ThunkOperation *op = thunk->operation;
  if (!op) return;
  for (int i = 0; i < (int)op->arity; i++) {
    PyObject *parameter = op->operand[i];
    if (parameter && PyThunk_Check(parameter)) {
      std::vector<PyThunkObject *> *ch = ((PyThunkObject *)parameter)->dependents;
      ch->erase(std::remove(ch->begin(), ch->end(), thunk), ch->end());
    }
  }
  thunk->operation = NULL;
  DestroyOperation(op);
}

void PyThunk_EvaluateChildren(PyThunkObject *thunk) {
  // evaluating a child unregisters it, so walk a snapshot (thunk.cpp:81-88)
  std::vector<PyThunkObject *> snapshot(*thunk->dependents);
  for (PyThunkObject *c : snapshot) Py_INCREF(c);
  for (PyThunkObject *c : snapshot) {
    if (!PyErr_Occurred()) PyThunk_Evaluate(c);
    Py_DECREF(c);
  }
}

PyObject *PyThunk_Evaluate(PyThunkObject *thunk) {
  if (PyThunk_IsEvaluated(thunk)) Py_RETURN_NONE;
  // split the DAG into standalone pipelines at the blocking operations (thunk.cpp:90-99)
  // ... and run them: children first, one fused kernel per pipeline per device shard
  if (ScheduleThunk(thunk) < 0) return NULL;
  if (!thunk->evaluated) {
    PyErr_SetString(PyExc_ValueError, "Something went wrong evaluating a Thunk.");   // thunk.cpp:108-111
    return NULL;
  }
  Py_RETURN_NONE;
}

// called by the scheduler once a pipeline output is computed (compiler.cpp:51)
void PyThunk_MarkEvaluated(PyThunkObject *thunk, Storage *storage) {
  if (thunk->storage) storage_decref(thunk->storage);
  thunk->storage = storage;
  thunk->evaluated = true;
  thunk->card.n = (ssize_t)storage->cardinality();
  thunk->card.kind = CardinalityKind::Exact;
  Py_CLEAR(thunk->host);
  release_operation(thunk);
}

bool PyThunk_IsEvaluated(PyThunkObject *thunk) { return thunk->evaluated; }

int PyThunk_EnsureDevice(PyThunkObject *thunk) {
  if (thunk->storage) return 0;
  if (!thunk->host) { PyErr_SetString(PyExc_ValueError, "thunk has no data"); return -1; }
  if (nljit_ensure_runtime() < 0) return -1;
  thunk->storage = storage_from_host((PyArrayObject *)thunk->host);
  return thunk->storage ? 0 : -1;
}

void PyThunk_InvalidateHost(PyThunkObject *thunk) {
  if (thunk->storage) Py_CLEAR(thunk->host);
}

PyObject *PyThunk_AsArray(PyObject *obj) {
  if (!PyThunk_Check(obj)) return NULL;
  PyThunkObject *thunk = (PyThunkObject *)obj;
  PyObject *r = PyThunk_Evaluate(thunk);
  if (r == NULL) return NULL;
  Py_DECREF(r);
  if (!thunk->host) {
    thunk->host = storage_to_host(thunk->storage);
    if (!thunk->host) return NULL;
    // The reference hands out its storage itself (thunk.cpp:115-124); this is a host COPY of device
    // storage, cached until the thunk is assigned to.  Read-only, so that a write meant for the thunk
    // fails loudly instead of silently diverging from the device data.
    PyArray_CLEARFLAGS((PyArrayObject *)thunk->host, NPY_ARRAY_WRITEABLE);
  }
  return thunk->host;
}

static PyObject *thunk_str(PyThunkObject *self) {
  PyObject *arr = PyThunk_AsArray((PyObject *)self);   // evaluates (thunk.cpp:126-133)
  if (arr == NULL) return NULL;
  return PyObject_Str(arr);
}

static void PyThunk_dealloc(PyThunkObject *self) {
  release_operation(self);
  // anything still registered as our dependent holds a reference to us through its operation,
  // so the list is empty here; clear defensively
  delete self->dependents;
  self->dependents = NULL;
  storage_decref(self->storage);
  Py_XDECREF(self->host);
  free(self->name);
  Py_TYPE(self)->tp_free((PyObject *)self);
}

static PyObject *thunk_new(PyTypeObject *type, PyObject *args, PyObject *kwds) {
  (void)type; (void)kwds;
  PyObject *input = NULL;
  if (!PyArg_ParseTuple(args, "O", &input)) return NULL;
  return PyThunk_FromArray(NULL, input);
}

PyTypeObject PyThunk_Type = {
    PyVarObject_HEAD_INIT(NULL, 0)
    "thunk",                 /* tp_name (thunk.cpp:144) */
    sizeof(PyThunkObject),   /* tp_basicsize */
};

int PyThunk_Init(void) {
  PyThunk_Type.tp_dealloc = (destructor)PyThunk_dealloc;
  PyThunk_Type.tp_as_number = &thunk_as_number;
  PyThunk_Type.tp_as_sequence = &thunk_as_sequence;
  PyThunk_Type.tp_as_mapping = &thunk_as_mapping;
  PyThunk_Type.tp_hash = (hashfunc)PyObject_HashNotImplemented;   /* unhashable (thunk.cpp:156) */
  PyThunk_Type.tp_str = (reprfunc)thunk_str;
  PyThunk_Type.tp_flags = Py_TPFLAGS_DEFAULT | Py_TPFLAGS_BASETYPE;
  PyThunk_Type.tp_doc = "Thunk.";
  PyThunk_Type.tp_richcompare = (richcmpfunc)PyThunk_LazyRichCompare;
  PyThunk_Type.tp_methods = thunk_methods;
  PyThunk_Type.tp_alloc = PyType_GenericAlloc;
  PyThunk_Type.tp_new = thunk_new;
  return PyType_Ready(&PyThunk_Type);
}
