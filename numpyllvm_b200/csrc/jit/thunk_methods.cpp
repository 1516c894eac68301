// thunk_methods.cpp — evaluate / isevaluated / asnumpyarray / sort / max / min / sum
// (reference: thunk_methods.cpp:10-199).  max and sum are "parallel breakers": they end a
// pipeline but are still computed inside the fused kernel; their LLVM emitters
// (llvm_generate_max/_sum, llvm_initialize_*, llvm_finalize_sum_data) are replaced by the
// NL_OP_MAX / NL_OP_SUM op-codes plus the deterministic grid finish and the NCCL allreduce
// behind the C-ABI.  `min` is the obvious sibling on the same slot.  sort stays what it is in
// the reference: a full breaker whose base function is NumPy's sort on a copy.
#include "thunk.hpp"

static PyObject *_thunk_evaluate(PyThunkObject *self, PyObject *args) {
  (void)args;
  if (PyThunk_Evaluate(self) == NULL) return NULL;
  Py_RETURN_NONE;
}

static PyObject *_thunk_isevaluated(PyThunkObject *self, PyObject *args) {
  (void)args;
  if (PyThunk_IsEvaluated(self)) Py_RETURN_TRUE;
  Py_RETURN_FALSE;
}

static PyObject *_thunk_asarray(PyThunkObject *self, PyObject *args) {
  (void)args;
  PyObject *arr = PyThunk_AsArray((PyObject *)self);
  if (arr == NULL) return NULL;     // the reference INCREFs a possible NULL here (thunk_methods.cpp:32)
  Py_INCREF(arr);
  return arr;
}

static PyObject *_thunk_inline_sort(PyObject *input) {
  // thunk_methods.cpp:37-42: quicksort of a copy
  PyArrayObject *output = (PyArrayObject *)PyArray_NewCopy((PyArrayObject *)input, NPY_CORDER);
  if (!output) return NULL;
  if (PyArray_Sort(output, 0, NPY_QUICKSORT) < 0) { Py_DECREF(output); return NULL; }
  return (PyObject *)output;
}

static PyObject *_thunk_sort(PyThunkObject *self, PyObject *unused) {
  (void)unused;
  if (!PyThunk_Check(self)) {
    PyErr_SetString(PyExc_TypeError, "Expected a thunk object as parameter.");
    return NULL;
  }
  ThunkOperation *op = NewUnaryOperation((PyObject *)self, OpClass::FullBreaker, -1,
                                                (HostFunction)_thunk_inline_sort, "sort");
  if (!op) return PyErr_NoMemory();
  return PyThunk_FromOperation(op, cardinality_same_as_operand, CardinalityKind::Exact, self->npy_type);
}

static ssize_t aggregate_cardinality_function(ssize_t *inputs) {
  (void)inputs;
  return 1;
}

static PyObject *lazy_reduce(PyThunkObject *self, int opcode, const char *name) {
  if (!PyThunk_Check(self)) {
    PyErr_SetString(PyExc_TypeError, "Expected a thunk object as parameter.");
    return NULL;
  }
  if (self->npy_type == NPY_BOOL) {
    PyErr_SetString(PyExc_TypeError, "reductions of boolean thunks are not supported");
    return NULL;
  }
  ThunkOperation *op = NewUnaryOperation((PyObject *)self, OpClass::ParallelBreaker, opcode, NULL, name);
  if (!op) return PyErr_NoMemory();
  return PyThunk_FromOperation(op, aggregate_cardinality_function, CardinalityKind::Exact, self->npy_type);
}

static PyObject *_thunk_max(PyThunkObject *self, PyObject *unused) { (void)unused; return lazy_reduce(self, NL_OP_MAX, "max"); }
static PyObject *_thunk_min(PyThunkObject *self, PyObject *unused) { (void)unused; return lazy_reduce(self, NL_OP_MIN, "min"); }
static PyObject *_thunk_sum(PyThunkObject *self, PyObject *unused) { (void)unused; return lazy_reduce(self, NL_OP_SUM, "sum"); }

// Host<->device edge: where the evaluated data lives, so a consumer can keep working
// on the GPUs instead of paying the D2H copy of asnumpyarray().  One (device, pointer, first global
// index, count, typestr) tuple per shard; the pointers stay valid while the thunk is alive and is
// not assigned to.  numpyllvm_b200.device_shards() wraps them as __cuda_array_interface__ objects.
static PyObject *_thunk_device_shards(PyThunkObject *self, PyObject *args) {
  (void)args;
  if (PyThunk_Evaluate(self) == NULL) return NULL;
  if (PyThunk_EnsureDevice(self) < 0) return NULL;
  Storage *s = self->storage;
  if (!s || s->host_scalar) {
    PyErr_SetString(PyExc_ValueError, "a size-1 thunk made from host data has no device buffer");
    return NULL;
  }
  const char *typestr = s->dtype == NL_I32 ? "<i4" : s->dtype == NL_I64 ? "<i8" : s->dtype == NL_F32 ? "<f4"
                        : s->dtype == NL_F64 ? "<f8" : "|b1";
  if (storage_settle(s) < 0 || nljit_sync_devices() < 0) return NULL;
  PyObject *list = PyList_New(0);
  if (!list) return NULL;
  for (const Shard &sh : s->shards) {
    if (s->replicated && sh.dev != 0) continue;          // an allreduced scalar is the same everywhere
    PyObject *t = Py_BuildValue("(iKLLs)", sh.dev, (unsigned long long)(uintptr_t)sh.ptr, (long long)sh.start,
                                (long long)(s->replicated ? 1 : sh.count), typestr);
    if (!t || PyList_Append(list, t) < 0) { Py_XDECREF(t); Py_DECREF(list); return NULL; }
    Py_DECREF(t);
  }
  return list;
}

struct PyMethodDef thunk_methods[] = {
    {"evaluate", (PyCFunction)_thunk_evaluate, METH_NOARGS, "evaluate() => "},
    {"isevaluated", (PyCFunction)_thunk_isevaluated, METH_NOARGS, "isevaluated() => "},
    {"asnumpyarray", (PyCFunction)_thunk_asarray, METH_NOARGS, "asnumpyarray() => "},
    {"sort", (PyCFunction)_thunk_sort, METH_NOARGS, "sort() => "},
    {"max", (PyCFunction)_thunk_max, METH_NOARGS, "max() => "},
    {"min", (PyCFunction)_thunk_min, METH_NOARGS, "min() => "},
    {"sum", (PyCFunction)_thunk_sum, METH_NOARGS, "sum() => "},
    {"device_shards", (PyCFunction)_thunk_device_shards, METH_NOARGS,
     "device_shards() => [(device, pointer, first_index, count, typestr), ...] of the evaluated thunk"},
    {NULL, NULL, 0, NULL} /* Sentinel */
};
