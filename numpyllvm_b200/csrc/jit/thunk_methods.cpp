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
  static const char *typestrs[NL_N_DTYPES] = {"|b1", "<i4", "<i8", "<f4", "<f8", "|i1", "<i2", "|u1", "<u2", "<u4", "<u8"};
  const char *typestr = typestrs[s->dtype];
  if (storage_settle(s) < 0 || nljit_sync_devices() < 0) return NULL;
  PyObject *list = PyList_New(0);
  if (!list) return NULL;
  for (const Shard &sh : s->shards) {
    if (s->replicated && sh.dev != 0) continue;          // an allreduced scalar is the same everywhere
    PyObject *t = Py_BuildValue("(iKLLs)", nl_device_cuda_id(sh.dev), (unsigned long long)(uintptr_t)sh.ptr, (long long)sh.start,
                                (long long)(s->replicated ? 1 : sh.count), typestr);
    if (!t || PyList_Append(list, t) < 0) { Py_XDECREF(t); Py_DECREF(list); return NULL; }
    Py_DECREF(t);
  }
  return list;
}

// __array__(dtype=None, copy=None): numpy.asarray(thunk) / numpy.array(thunk) materialise the thunk like
// asnumpyarray() (thunk.cpp:115-124) — the NumPy side of the host<->device edge
static PyObject *_thunk_array(PyThunkObject *self, PyObject *args, PyObject *kwargs) {
  static const char *kwlist[] = {"dtype", "copy", NULL};
  PyObject *dtype = Py_None, *copy = Py_None;
  if (!PyArg_ParseTupleAndKeywords(args, kwargs, "|OO", (char **)kwlist, &dtype, &copy)) return NULL;
  PyObject *arr = PyThunk_AsArray((PyObject *)self);
  if (arr == NULL) return NULL;
  if (copy == Py_False && dtype != Py_None) {
    PyErr_SetString(PyExc_ValueError, "a thunk cannot be converted to another dtype without a copy");
    return NULL;
  }
  if (dtype == Py_None && copy != Py_True) { Py_INCREF(arr); return arr; }     // the cached, read-only host copy
  PyArray_Descr *descr = NULL;
  if (dtype != Py_None) {
    if (!PyArray_DescrConverter(dtype, &descr)) return NULL;
  } else {
    descr = PyArray_DESCR((PyArrayObject *)arr);
    Py_INCREF(descr);
  }
  return PyArray_FromArray((PyArrayObject *)arr, descr, NPY_ARRAY_ENSURECOPY | NPY_ARRAY_WRITEABLE | NPY_ARRAY_FORCECAST);   // steals descr
}

// ---- DLPack export of one shard (consumers: torch.from_dlpack, cupy.from_dlpack, jax) -------------------
// The structs are the stable DLPack ABI (dlpack.h v0.8); the capsule keeps the thunk alive until the
// consumer's deleter runs.
extern "C" {
typedef struct { int32_t device_type; int32_t device_id; } NLDLDevice;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } NLDLDataType;
typedef struct { void *data; NLDLDevice device; int32_t ndim; NLDLDataType dtype; int64_t *shape; int64_t *strides; uint64_t byte_offset; } NLDLTensor;
typedef struct NLDLManagedTensor { NLDLTensor dl_tensor; void *manager_ctx; void (*deleter)(struct NLDLManagedTensor *); } NLDLManagedTensor;
}
struct DLOwner { NLDLManagedTensor m; int64_t shape[1]; PyObject *thunk; };

static void dl_deleter(NLDLManagedTensor *m) {
  DLOwner *o = (DLOwner *)m->manager_ctx;
  PyGILState_STATE g = PyGILState_Ensure();
  Py_XDECREF(o->thunk);
  PyGILState_Release(g);
  delete o;
}
static void dl_capsule_destructor(PyObject *cap) {
  // a capsule nobody consumed still owns the tensor (a consumer renames it to "used_dltensor")
  if (PyCapsule_IsValid(cap, "dltensor")) {
    NLDLManagedTensor *m = (NLDLManagedTensor *)PyCapsule_GetPointer(cap, "dltensor");
    if (m && m->deleter) m->deleter(m);
  }
}

// _dlpack(i) -> "dltensor" capsule of shard i (numpyllvm_b200.DeviceShard.__dlpack__ calls this)
static PyObject *_thunk_dlpack(PyThunkObject *self, PyObject *arg) {
  const long idx = PyLong_AsLong(arg);
  if (idx == -1 && PyErr_Occurred()) return NULL;
  if (PyThunk_Evaluate(self) == NULL) return NULL;
  if (PyThunk_EnsureDevice(self) < 0) return NULL;
  Storage *s = self->storage;
  if (!s || s->host_scalar) { PyErr_SetString(PyExc_ValueError, "a size-1 thunk made from host data has no device buffer"); return NULL; }
  if (idx < 0 || idx >= (long)s->shards.size()) { PyErr_SetString(PyExc_IndexError, "shard index out of range"); return NULL; }
  if (storage_settle(s) < 0 || nljit_sync_devices() < 0) return NULL;
  const Shard &sh = s->shards[(size_t)idx];
  DLOwner *o = new DLOwner();
  o->shape[0] = s->replicated ? 1 : sh.count;
  o->thunk = (PyObject *)self;
  Py_INCREF(o->thunk);
  o->m.dl_tensor.data = sh.ptr;
  o->m.dl_tensor.device.device_type = 2;                   // kDLCUDA
  o->m.dl_tensor.device.device_id = nl_device_cuda_id(sh.dev);
  o->m.dl_tensor.ndim = 1;
  const int dt = s->dtype;
  const bool uns = dt == NL_U8 || dt == NL_U16 || dt == NL_U32 || dt == NL_U64;
  o->m.dl_tensor.dtype.code = (dt == NL_F32 || dt == NL_F64) ? 2 : (dt == NL_BOOL ? 6 : (uns ? 1 : 0));   // kDLFloat / kDLBool / kDLUInt / kDLInt
  o->m.dl_tensor.dtype.bits = (uint8_t)(8 * nl_dtype_size(dt));
  o->m.dl_tensor.dtype.lanes = 1;
  o->m.dl_tensor.shape = o->shape;
  o->m.dl_tensor.strides = NULL;
  o->m.dl_tensor.byte_offset = 0;
  o->m.manager_ctx = o;
  o->m.deleter = dl_deleter;
  PyObject *cap = PyCapsule_New(&o->m, "dltensor", dl_capsule_destructor);
  if (!cap) { dl_deleter(&o->m); return NULL; }
  return cap;
}

// astype(dtype) => lazy conversion to another element type (NumPy's ndarray.astype)
static PyObject *_thunk_astype(PyThunkObject *self, PyObject *arg) {
  PyArray_Descr *descr = NULL;
  if (!PyArray_DescrConverter(arg, &descr)) return NULL;
  const int t = descr->type_num;
  Py_DECREF(descr);
  return PyThunk_AsType(self, t);
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
    {"astype", (PyCFunction)_thunk_astype, METH_O, "astype(dtype) => lazy conversion to another element type"},
    {"__array__", (PyCFunction)(void (*)(void))_thunk_array, METH_VARARGS | METH_KEYWORDS, "__array__(dtype=None, copy=None) => host ndarray (numpy.asarray(thunk))"},
    {"_dlpack", (PyCFunction)_thunk_dlpack, METH_O, "_dlpack(i) => DLPack capsule of shard i (see numpyllvm_b200.DeviceShard)"},
    {NULL, NULL, 0, NULL} /* Sentinel */
};
