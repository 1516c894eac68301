// thunk_as_mapping.cpp — t[mask] and t[index] = value (reference: thunk_as_mapping.cpp:9-97).
//
// Subscript: a boolean thunk index builds the lazy filter, exactly as in the reference; the
// LLVM emitters (llvm_generate_subscript / llvm_initialize_subscript) are replaced by the
// NL_OP_FILTER op-code.  Assignment keeps the reference's ordering — force every dependent of
// the target, force the target, then update it in place (thunk_as_mapping.cpp:80-86) — but
// the update runs on the device (the reference delegated to NumPy setitem on host storage).
#include "thunk.hpp"
#include <vector>
#include "scheduler.hpp"
#include "device_sort.hpp"

#include <algorithm>
#include <utility>
#include <vector>
#include <cstring>

static ssize_t subscript_cardinality_function(ssize_t *inputs) {
  return inputs[1];     // upper bound: the mask's size (thunk_as_mapping.cpp:32-35)
}

// a[idx] with an int64 index thunk: declared in the reference (cardinality function, IndexError
// FIXME) but without an emitter (thunk_as_mapping.cpp:51-55).  Here it is an eager operation: both
// operands are evaluated, every GPU gathers the elements its shard of `idx` names — from whichever
// GPU holds them, with peer loads over NVLink — and the result is an evaluated thunk laid out like
// `idx`.  Negative indices wrap as in NumPy; an index outside [-n, n) raises IndexError.
static PyObject *thunk_gather(PyThunkObject *self, PyThunkObject *index) {
  if (PyThunk_Evaluate(self) == NULL || PyThunk_Evaluate(index) == NULL) return NULL;
  if (PyThunk_EnsureDevice(self) < 0 || PyThunk_EnsureDevice(index) < 0) return NULL;
  Storage *col = self->storage, *idx = index->storage;
  if (col->host_scalar || idx->host_scalar || col->replicated || idx->replicated) {
    PyErr_SetString(PyExc_IndexError, "integer-array indexing of / by a one-element thunk is not supported");
    return NULL;
  }
  if (nljit_ensure_runtime() < 0) return NULL;
  if (storage_settle(col) < 0 || storage_settle(idx) < 0) return NULL;
  // peers read `col` with plain loads over NVLink: what other GPUs' streams still owe it must be done
  if (nljit_device_count() > 1 && nljit_sync_devices() < 0) return NULL;
  const int64_t total = col->cardinality();
  const void *ptrs[NL_MAX_DEVICES];
  int64_t starts[NL_MAX_DEVICES];
  int n_shards = 0;
  int64_t run = 0;
  for (const Shard &sh : col->shards) {
    ptrs[n_shards] = sh.ptr ? sh.ptr : (const void *)16;
    starts[n_shards] = run;
    run += sh.count;
    n_shards++;
  }
  Storage *out = storage_new(col->dtype, idx->bound);
  out->ragged = idx->ragged;
  const size_t es = nl_dtype_size(col->dtype);
  int rc = NL_OK;
  std::vector<int64_t *> bad(idx->shards.size(), (int64_t *)NULL);
  for (size_t g = 0; g < idx->shards.size() && rc == NL_OK; g++) {
    const Shard &is = idx->shards[g];
    Shard os = is;
    os.ptr = NULL;
    os.capacity = is.count;
    if (is.count > 0) rc = nl_malloc(is.dev, (size_t)is.count * es, &os.ptr);
    out->shards.push_back(os);
    if (rc != NL_OK || is.count == 0) continue;
    rc = nl_malloc(is.dev, sizeof(int64_t), (void **)&bad[g]);
    if (rc == NL_OK) rc = nl_memset(is.dev, 0, bad[g], 0, sizeof(int64_t));
    if (rc == NL_OK)
      rc = nl_gather(is.dev, 0, col->dtype, os.ptr, (const int64_t *)is.ptr, is.count, ptrs, starts, n_shards, total, bad[g]);
  }
  int64_t n_bad = 0;
  for (size_t g = 0; g < bad.size(); g++) {
    if (!bad[g]) continue;
    int64_t v = 0;
    if (rc == NL_OK) rc = nl_memcpy_d2h(idx->shards[g].dev, 0, &v, bad[g], sizeof(v));
    if (rc == NL_OK) rc = nl_stream_sync(idx->shards[g].dev, 0);
    n_bad += v;
    nl_free(idx->shards[g].dev, bad[g]);
  }
  if (rc != NL_OK) { storage_decref(out); nljit_raise(rc, "jit: gather failed"); return NULL; }
  if (n_bad > 0) {
    storage_decref(out);
    PyErr_Format(PyExc_IndexError, "%lld indices are out of bounds for a thunk of %lld elements", (long long)n_bad, (long long)total);
    return NULL;
  }
  // an evaluated thunk around the gathered storage
  npy_intp zero = 0;
  PyObject *empty = PyArray_SimpleNew(1, &zero, self->npy_type);
  if (!empty) { storage_decref(out); return NULL; }
  PyObject *t = PyThunk_FromArray(NULL, empty);
  Py_DECREF(empty);
  if (!t) { storage_decref(out); return NULL; }
  PyThunk_MarkEvaluated((PyThunkObject *)t, out);
  return t;
}

static PyObject *thunk_subscript(PyObject *self_o, PyObject *other) {
  PyThunkObject *self = (PyThunkObject *)self_o;
  if (PyArray_Check(other) && PyArray_ISINTEGER((PyArrayObject *)other) && PyArray_NDIM((PyArrayObject *)other) == 1) {
    // an integer ndarray as index: wrap it (as int64) and gather
    PyObject *as64 = PyArray_FromAny(other, PyArray_DescrFromType(NPY_INT64), 1, 1, NPY_ARRAY_C_CONTIGUOUS | NPY_ARRAY_FORCECAST, NULL);
    if (!as64) return NULL;
    PyObject *it = PyThunk_FromArray(NULL, as64);
    Py_DECREF(as64);
    if (!it) return NULL;
    PyObject *r = thunk_gather(self, (PyThunkObject *)it);
    Py_DECREF(it);
    return r;
  }
  if (!PyThunk_Check(other)) {
    PyErr_SetString(PyExc_IndexError, "FIXME: other types of array access");   // thunk_as_mapping.cpp:61
    return NULL;
  }
  PyThunkObject *other_thunk = (PyThunkObject *)other;
  if (other_thunk->npy_type == NPY_BOOL) {
    if (self->npy_type == NPY_BOOL) {
      PyErr_SetString(PyExc_IndexError, "filtering a boolean thunk is not supported");
      return NULL;
    }
    if (self->card.n > 1 && other_thunk->card.n > 1 && self->card.kind == CardinalityKind::Exact &&
        other_thunk->card.kind == CardinalityKind::Exact && self->card.n != other_thunk->card.n) {
      PyErr_SetString(PyExc_IndexError, "boolean index did not match the indexed thunk's cardinality");
      return NULL;
    }
    // boolean access: cardinality depends on the number of TRUE entries (upper bound)
    ThunkOperation *op = NewBinaryOperation(self_o, other, OpClass::Vectorizable, NL_OP_FILTER, "[]");
    if (!op) return PyErr_NoMemory();
    return PyThunk_FromOperation(op, subscript_cardinality_function, CardinalityKind::UpperBound, self->npy_type);
  }
  if (other_thunk->npy_type == NPY_INT64) return thunk_gather(self, other_thunk);
  if (other_thunk->npy_type == NPY_INT32) {
    PyErr_SetString(PyExc_IndexError, "integer-array indexing takes int64 indices");
    return NULL;
  }
  PyErr_SetString(PyExc_IndexError, "arrays used as indices must be of integer (or boolean) type");
  return NULL;
}

// ---- in-place assignment ---------------------------------------------------------------

// value -> bit pattern in the target's dtype (NumPy casting of a scalar on setitem)
static int scalar_bits_of(PyObject *value, int npy_type, uint64_t *bits) {
  PyArray_Descr *descr = PyArray_DescrFromType(npy_type);
  if (!descr) return -1;
  PyObject *arr = PyArray_FromAny(value, descr, 0, 0, NPY_ARRAY_FORCECAST, NULL);   // steals descr
  if (!arr) return -1;
  if (PyArray_SIZE((PyArrayObject *)arr) != 1) {
    Py_DECREF(arr);
    PyErr_SetString(PyExc_ValueError, "setting a thunk element with a sequence");
    return -1;
  }
  *bits = 0;
  memcpy(bits, PyArray_DATA((PyArrayObject *)arr), (size_t)PyArray_ITEMSIZE((PyArrayObject *)arr));
  Py_DECREF(arr);
  return 0;
}

// part of the global index run lo, lo+step, ... (count items) that falls into [s0, s1)
static bool intersect_run(int64_t lo, int64_t step, int64_t count, int64_t s0, int64_t s1, int64_t *kmin, int64_t *kmax) {
  if (count <= 0 || s1 <= s0) return false;
  int64_t a, b;
  if (step > 0) {
    a = (s0 - lo + step - 1) / step;              // ceil
    if (s0 <= lo) a = 0;
    if (s1 - 1 < lo) return false;
    b = (s1 - 1 - lo) / step;
  } else {
    const int64_t m = -step;
    a = (lo - (s1 - 1) + m - 1) / m;
    if (lo <= s1 - 1) a = 0;
    if (lo < s0) return false;
    b = (lo - s0) / m;
  }
  if (b > count - 1) b = count - 1;
  if (a > b) return false;
  *kmin = a;
  *kmax = b;
  return true;
}

static int assign_run(PyThunkObject *self, int64_t lo, int64_t step, int64_t count, PyObject *value) {
  Storage *st = self->storage;
  const int nl_type = st->dtype;
  const size_t es = nl_dtype_size(nl_type);
  if (st->host_scalar) {
    uint64_t bits;
    if (count != 1 || scalar_bits_of(value, self->npy_type, &bits) < 0) {
      if (!PyErr_Occurred()) PyErr_SetString(PyExc_IndexError, "index out of range for a size-1 thunk");
      return -1;
    }
    st->scalar_bits = bits;
    return 0;
  }
  if (st->ragged && st->shards.size() > 1) {
    // a filter result keeps per-shard runs of different lengths; an index into it means the concatenated array
    // (what the reference's gather leaves behind, compiler.cpp:36-48): re-cut it into the regular partition first
    Storage *regular = storage_repartition(st);
    if (!regular) return -1;
    storage_decref(self->storage);
    self->storage = regular;
    st = regular;
  }
  // a column-valued thunk of the target's dtype stays on the GPUs: the pieces each target shard needs
  // travel device-to-device (NVLink peer copies across GPUs), no host round trip
  if (PyThunk_Check(value) && ((PyThunkObject *)value)->npy_type == self->npy_type) {
    PyThunkObject *vt = (PyThunkObject *)value;
    if (PyThunk_Evaluate(vt) == NULL) return -1;
    if (vt->card.n > 1) {
      if (PyThunk_EnsureDevice(vt) < 0) return -1;
      Storage *vs = vt->storage;
      if (storage_settle(vs) < 0) return -1;
      if (vs == st) {
        // a[...] = a: the whole column onto itself is a no-op; anything else goes through a host copy below
        if (lo == 0 && step == 1 && count == vs->cardinality()) return 0;
        goto host_path;
      }
      if (vs->cardinality() != count) {
        PyErr_Format(PyExc_ValueError, "could not broadcast input array of size %lld into a selection of size %lld",
                     (long long)vs->cardinality(), (long long)count);
        return -1;
      }
      int rc = NL_OK;
      Py_BEGIN_ALLOW_THREADS
      for (int g = 0; g < nljit_device_count() && rc == NL_OK; g++) rc = nl_stream_sync(g, 0);   // the value is complete everywhere
      Py_END_ALLOW_THREADS
      std::vector<std::pair<int, void *>> temps;
      int64_t gstart = 0;
      for (Shard &sh : st->shards) {
        if (rc != NL_OK) break;
        const int64_t s0 = st->ragged ? gstart : sh.start, s1 = s0 + sh.count;
        gstart += sh.count;
        int64_t kmin, kmax;
        if (!intersect_run(lo, step, count, s0, s1, &kmin, &kmax)) continue;
        const int64_t llo = lo + kmin * step - s0, lcount = kmax - kmin + 1;
        if (step == 1) {
          rc = storage_copy_range(vs, kmin, kmax + 1, sh.dev, (char *)sh.ptr + (size_t)llo * es);   // straight into place
        } else {
          void *tmp = NULL;
          rc = nl_malloc(sh.dev, (size_t)lcount * es, &tmp);
          if (rc == NL_OK) temps.push_back(std::make_pair(sh.dev, tmp));
          if (rc == NL_OK) rc = storage_copy_range(vs, kmin, kmax + 1, sh.dev, tmp);
          if (rc == NL_OK) rc = nl_assign_copy(sh.dev, 0, sh.ptr, nl_type, llo, step, lcount, tmp);
        }
      }
      Py_BEGIN_ALLOW_THREADS
      for (int g = 0; g < nljit_device_count(); g++) {
        const int r2 = nl_stream_sync(g, 0);
        if (rc == NL_OK) rc = r2;
      }
      Py_END_ALLOW_THREADS
      for (auto &t : temps) nl_free(t.first, t.second);
      if (rc != NL_OK) { nljit_raise(rc, "jit: in-place assignment failed"); return -1; }
      return 0;
    }
  }
host_path:
  // scalar or array value?
  PyObject *varr = NULL;
  bool scalar = true;
  if (PyThunk_Check(value)) {
    PyThunkObject *vt = (PyThunkObject *)value;
    if (vt->card.n != 1 || !vt->evaluated) {
      varr = PyThunk_AsArray(value);           // forces it; borrowed
      if (!varr) return -1;
      Py_INCREF(varr);
      scalar = PyArray_SIZE((PyArrayObject *)varr) == 1;
    } else {
      varr = storage_to_host(vt->storage);
      if (!varr) return -1;
    }
  } else {
    PyArray_Descr *descr = PyArray_DescrFromType(self->npy_type);
    varr = PyArray_FromAny(value, descr, 0, 0, NPY_ARRAY_FORCECAST | NPY_ARRAY_C_CONTIGUOUS | NPY_ARRAY_ALIGNED, NULL);
    if (!varr) return -1;
    scalar = PyArray_SIZE((PyArrayObject *)varr) == 1;
  }
  if (PyArray_TYPE((PyArrayObject *)varr) != self->npy_type) {
    PyObject *cast = PyArray_Cast((PyArrayObject *)varr, self->npy_type);
    Py_DECREF(varr);
    if (!cast) return -1;
    varr = cast;
  }
  if (!scalar && (int64_t)PyArray_SIZE((PyArrayObject *)varr) != count) {
    PyErr_Format(PyExc_ValueError, "could not broadcast input array of size %lld into a selection of size %lld",
                 (long long)PyArray_SIZE((PyArrayObject *)varr), (long long)count);
    Py_DECREF(varr);
    return -1;
  }
  uint64_t bits = 0;
  if (scalar) memcpy(&bits, PyArray_DATA((PyArrayObject *)varr), es);
  const char *vdata = (const char *)PyArray_DATA((PyArrayObject *)varr);

  int rc = NL_OK;
  int64_t gstart = 0;     // global index of the shard's first element (== sh.start unless ragged on one device)
  for (Shard &sh : st->shards) {
    const int64_t s0 = st->ragged ? gstart : sh.start, s1 = s0 + sh.count;
    gstart += sh.count;
    int64_t kmin, kmax;
    if (!intersect_run(lo, step, count, s0, s1, &kmin, &kmax)) continue;
    const int64_t llo = lo + kmin * step - s0, lcount = kmax - kmin + 1;
    if (scalar) {
      rc = nl_assign_fill(sh.dev, 0, sh.ptr, nl_type, llo, step, lcount, bits);
    } else {
      void *tmp = NULL;
      rc = nl_malloc(sh.dev, (size_t)lcount * es, &tmp);
      if (rc == NL_OK) rc = nl_memcpy_h2d(sh.dev, 0, tmp, vdata + (size_t)kmin * es, (size_t)lcount * es);
      if (rc == NL_OK) rc = nl_assign_copy(sh.dev, 0, sh.ptr, nl_type, llo, step, lcount, tmp);
      if (rc == NL_OK) rc = nl_stream_sync(sh.dev, 0);
      if (tmp) nl_free(sh.dev, tmp);
    }
    if (rc != NL_OK) break;
  }
  Py_DECREF(varr);
  if (rc != NL_OK) { nljit_raise(rc, "jit: in-place assignment failed"); return -1; }
  return 0;
}

static int thunk_assign_subscript(PyObject *self_o, PyObject *ind, PyObject *value) {
  PyThunkObject *self = (PyThunkObject *)self_o;
  if (value == NULL) {
    PyErr_SetString(PyExc_TypeError, "thunk elements cannot be deleted");
    return -1;
  }

  // a[mask] = scalar with a mask that is a pure temporary (a[a > 5] = 0) is fused into one
  // in-place pass; a mask someone else can still observe is forced first like any dependent
  PyObject *where = NULL;
  bool fused_mask = false;
  if (PyThunk_Check(ind)) {
    PyThunkObject *mask = (PyThunkObject *)ind;
    if (mask->npy_type != NPY_BOOL) {
      PyErr_SetString(PyExc_IndexError, "only boolean thunks can be used as assignment masks");
      return -1;
    }
    PyObject *vt = PyThunk_Coerce(value);
    if (!vt) return -1;
    if (((PyThunkObject *)vt)->card.n != 1) {
      Py_DECREF(vt);
      PyErr_SetString(PyExc_ValueError, "masked assignment takes a scalar value");
      return -1;
    }
    if (((PyThunkObject *)vt)->npy_type != self->npy_type) {
      // cast the scalar to the target dtype (NumPy setitem casting)
      uint64_t bits;
      PyObject *host = PyThunk_AsArray(vt);
      if (!host || scalar_bits_of(host, self->npy_type, &bits) < 0) { Py_DECREF(vt); return -1; }
      Py_DECREF(vt);
      npy_intp one = 1;
      PyObject *arr = PyArray_SimpleNew(1, &one, self->npy_type);
      if (!arr) return -1;
      memcpy(PyArray_DATA((PyArrayObject *)arr), &bits, (size_t)PyArray_ITEMSIZE((PyArrayObject *)arr));
      vt = PyThunk_FromArray(NULL, arr);
      Py_DECREF(arr);
      if (!vt) return -1;
    }
    fused_mask = !mask->evaluated && Py_REFCNT(ind) == 1 && mask->dependents->empty();
    ThunkOperation *op = NewBinaryOperation(vt, ind, OpClass::ParallelBreaker, NL_OP_WHERE_STORE, "[]=");
    Py_DECREF(vt);
    if (!op) { PyErr_NoMemory(); return -1; }
    where = PyThunk_FromOperation(op, cardinality_broadcast, CardinalityKind::Exact, self->npy_type);
    if (!where) return -1;
  }

  // every thunk built from `self` must observe the OLD contents (thunk.cpp:81-88)
  {
    std::vector<PyThunkObject *> snapshot(*self->dependents);
    for (PyThunkObject *c : snapshot) Py_INCREF(c);
    for (PyThunkObject *c : snapshot) {
      const bool skip = fused_mask && where && (PyObject *)c == ind;
      // dependents reachable only through the fused mask are the assignment itself
      if (!skip && !PyErr_Occurred()) {
        PyObject *r = PyThunk_Evaluate(c);
        Py_XDECREF(r);
      }
      Py_DECREF(c);
    }
    if (fused_mask && where) {
      // children of the mask's own operands (e.g. `a` again) were handled above; the mask
      // expression may hang below other columns too — they are read, not written, so nothing to force
    }
    if (PyErr_Occurred()) { Py_XDECREF(where); return -1; }
  }
  PyObject *r = PyThunk_Evaluate(self);
  if (!r) { Py_XDECREF(where); return -1; }
  Py_DECREF(r);
  if (PyThunk_EnsureDevice(self) < 0) { Py_XDECREF(where); return -1; }
  // the update runs on the compute stream: behind the target's upload, and nobody may go on following
  // readiness marks that do not know about the write
  if (storage_settle(self->storage) < 0) { Py_XDECREF(where); return -1; }

  int rc = -1;
  if (where) {
    rc = ScheduleInPlace((PyThunkObject *)where, self);
    Py_DECREF(where);
  } else if (PySlice_Check(ind)) {
    Py_ssize_t start, stop, step;
    if (PySlice_Unpack(ind, &start, &stop, &step) < 0) return -1;
    const Py_ssize_t count = PySlice_AdjustIndices((Py_ssize_t)self->card.n, &start, &stop, step);
    rc = assign_run(self, (int64_t)start, (int64_t)step, (int64_t)count, value);
  } else if (PyIndex_Check(ind)) {
    Py_ssize_t i = PyNumber_AsSsize_t(ind, PyExc_IndexError);
    if (i == -1 && PyErr_Occurred()) return -1;
    if (i < 0) i += (Py_ssize_t)self->card.n;
    if (i < 0 || i >= (Py_ssize_t)self->card.n) {
      PyErr_Format(PyExc_IndexError, "index %zd is out of bounds for a thunk of size %zd", i, (Py_ssize_t)self->card.n);
      return -1;
    }
    rc = assign_run(self, (int64_t)i, 1, 1, value);
  } else {
    PyErr_SetString(PyExc_IndexError, "thunk assignment supports integer, slice and boolean-thunk indices");
    return -1;
  }
  if (rc == 0) PyThunk_InvalidateHost(self);
  return rc;
}

PyMappingMethods thunk_as_mapping = {
    (lenfunc)NULL,                          /* mp_length (NULL in the reference, thunk_as_mapping.cpp:89) */
    (binaryfunc)thunk_subscript,            /* mp_subscript */
    (objobjargproc)thunk_assign_subscript,  /* mp_ass_subscript */
};
