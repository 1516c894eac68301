// jitpackage.cpp — module `jit` (reference: jitpackage.cpp:10-39, ported from Py_InitModule3 to
// PyInit_jit).  jit.thunk(ndarray) is the one function the reference exports; the additions
// are host-side conveniences for the device edges and for inspecting plans.
#define NLJIT_IMPORT_ARRAY
#include "thunk.hpp"
#include "parser.hpp"
#include "optimizer.hpp"
#include "debug_printer.hpp"

#include <string>

static char module_docstring[] = "This module provides THUNKS.";
static char thunk_docstring[] = "thunk.thunk(array) => Creates a thunk array from a numpy array.";

static PyObject *jit_pinned_empty(PyObject *self, PyObject *args) {
  (void)self;
  Py_ssize_t n;
  PyObject *dtype_obj = NULL;
  if (!PyArg_ParseTuple(args, "n|O", &n, &dtype_obj)) return NULL;
  PyArray_Descr *descr = NULL;
  if (dtype_obj) {
    if (!PyArray_DescrConverter(dtype_obj, &descr)) return NULL;
  } else {
    descr = PyArray_DescrFromType(NPY_FLOAT64);
  }
  const int t = descr->type_num;
  Py_DECREF(descr);
  if (nljit_npy_to_nl(t) < 0) {
    PyErr_SetString(PyExc_TypeError, "pinned_empty: unsupported dtype");
    return NULL;
  }
  return nljit_pinned_ndarray(t, (int64_t)n);
}

static PyObject *jit_device_count(PyObject *self, PyObject *unused) {
  (void)self; (void)unused;
  if (nljit_ensure_runtime() < 0) return NULL;
  return PyLong_FromLong(nljit_device_count());
}

// synchronize(): everything enqueued so far (uploads, kernels, downloads) has completed on return
static PyObject *jit_synchronize(PyObject *self, PyObject *unused) {
  (void)self; (void)unused;
  if (nljit_ensure_runtime() < 0) return NULL;
  int rc = NL_OK;
  Py_BEGIN_ALLOW_THREADS
  for (int g = 0; g < nljit_device_count() && rc == NL_OK; g++) rc = nl_device_sync(g);
  Py_END_ALLOW_THREADS
  if (rc != NL_OK) { nljit_raise(rc, "jit: device synchronisation failed"); return NULL; }
  nljit_reap_uploads(true);
  Py_RETURN_NONE;
}

static PyObject *jit_set_chunk_bytes(PyObject *self, PyObject *arg) {
  (void)self;
  const long long b = PyLong_AsLongLong(arg);
  if (b == -1 && PyErr_Occurred()) return NULL;
  nljit_set_chunk_bytes((int64_t)b);
  Py_RETURN_NONE;
}

static PyObject *jit_launch_count(PyObject *self, PyObject *unused) {
  (void)self; (void)unused;
  return PyLong_FromUnsignedLongLong(nl_launch_count());
}

// plan(thunk) -> text: the pipelines a thunk would be split into and the kernel each would run.
// Pure host work (no device needed) for pipelines whose inputs are all evaluated.
static void describe(Pipeline *p, std::string &out) {
  for (Pipeline *c : p->children) describe(c, out);
  out += "pipeline ";
  out += p->name;
  ThunkOperation *root = (p->root)->op;
  if (root && root->opcode < 0) {
    out += ": base function ";
    out += root->name.c_str();
    out += "\n";
    return;
  }
  bool ready = true;
  for (Binding &e : p->inputs)
    if (e.operation->kind == Operation::ChildResult) ready = false;
  if (!ready) { out += ": fused kernel over the results of its child pipelines\n"; return; }
  LoweredPipeline lp;
  std::vector<PyThunkObject *> cut;
  int rc = LowerPipeline(p, &lp, &cut);
  if (rc == 1) { out += ": cannot run as one fused kernel, will be split\n"; return; }
  if (rc < 0) { PyErr_Clear(); out += ": cannot be compiled\n"; return; }
  char buf[4096], kernel[160];
  int is_static = 0;
  nl_plan_dump(&lp.plan, buf, sizeof buf);
  nl_plan_kernel(&lp.plan, 1, kernel, sizeof kernel, &is_static);
  out += " -> ";
  out += kernel;
  out += "\n";
  out += buf;
}

static PyObject *jit_plan(PyObject *self, PyObject *arg) {
  (void)self;
  if (!PyThunk_Check(arg)) { PyErr_SetString(PyExc_TypeError, "plan: expected a thunk"); return NULL; }
  PyThunkObject *t = (PyThunkObject *)arg;
  if (t->evaluated) return PyUnicode_FromString("evaluated\n");
  Pipeline *p = ParsePipeline(t, NULL);
  if (!p) return NULL;
  std::string out;
  describe(p, out);
  DestroyPipeline(p);
  return PyUnicode_FromString(out.c_str());
}

static PyMethodDef module_methods[] = {
    {"thunk", PyThunk_FromArray, METH_O, thunk_docstring},
    {"pinned_empty", jit_pinned_empty, METH_VARARGS, "pinned_empty(n, dtype) => 1-D ndarray in page-locked host memory (full-rate DMA)."},
    {"device_count", jit_device_count, METH_NOARGS, "device_count() => number of GPUs the arrays are sharded over."},
    {"synchronize", jit_synchronize, METH_NOARGS, "synchronize() => waits for every upload, kernel and download enqueued so far (evaluation is stream-ordered)."},
    {"set_chunk_bytes", jit_set_chunk_bytes, METH_O, "set_chunk_bytes(n) => size of the upload / chunked-evaluation / download chunks (0: default, 32 MiB)."},
    {"launch_count", jit_launch_count, METH_NOARGS, "launch_count() => kernels launched by the engine so far."},
    {"plan", jit_plan, METH_O, "plan(thunk) => text description of the fused pipelines and kernels."},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef jit_module = {PyModuleDef_HEAD_INIT, "jit", module_docstring, -1, module_methods, NULL, NULL, NULL, NULL};

extern "C" __attribute__((visibility("default"))) PyObject *PyInit_jit(void) {
  import_array();
  if (PyThunk_Init() < 0) return NULL;
  PyObject *m = PyModule_Create(&jit_module);
  if (m == NULL) return NULL;
  Py_INCREF(&PyThunk_Type);
  if (PyModule_AddObject(m, "thunk_type", (PyObject *)&PyThunk_Type) < 0) {
    Py_DECREF(&PyThunk_Type);
    Py_DECREF(m);
    return NULL;
  }
  // the reference starts its worker thread here (create_threads, jitpackage.cpp:33); the device
  // streams that replace it are created lazily on first use so that importing never needs a GPU
  return m;
}
