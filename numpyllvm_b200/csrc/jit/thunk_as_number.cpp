// thunk_as_number.cpp — lazy arithmetic and comparisons (reference: thunk_as_number.cpp:9-128).
// The reference implements `*` (nb_multiply) and `>=` only; `+`, `-` and the other five
// comparisons are filled in on the same slots with the same shape.  Instead of
// an LLVM emitter each operation records the op-code the kernel library executes.
#include "thunk.hpp"

static bool is_float_type(int t) { return t == NPY_FLOAT32 || t == NPY_FLOAT64; }

// The reference types the result after the LEFT operand (thunk_as_number.cpp:78, "todo: correct
// type") and would hand LLVM mismatched operand widths (Q3).  Here: scalars adopt the dtype of
// the array operand (NumPy's weak-scalar rule), two arrays must agree, and a float scalar is
// never silently truncated into an integer column.
static int binary_value_type(PyThunkObject *l, PyThunkObject *r, int *type_out) {
  const bool ls = l->card.n == 1, rs = r->card.n == 1;
  if (l->npy_type == NPY_BOOL || r->npy_type == NPY_BOOL) {
    PyErr_SetString(PyExc_TypeError, "arithmetic and comparisons on boolean thunks are not supported");
    return -1;
  }
  int t;
  if (ls && !rs) t = r->npy_type;
  else if (rs && !ls) t = l->npy_type;
  else {
    if (l->npy_type != r->npy_type && !(ls && rs)) {
      PyErr_SetString(PyExc_TypeError, "operands must have the same dtype (no implicit promotion between columns)");
      return -1;
    }
    t = l->npy_type;
  }
  PyThunkObject *scalar = (ls && !rs) ? l : (rs && !ls) ? r : NULL;
  if (scalar && is_float_type(scalar->npy_type) && !is_float_type(t)) {
    PyErr_SetString(PyExc_TypeError, "a float scalar cannot be combined with an integer column");
    return -1;
  }
  *type_out = t;
  return 0;
}

static int check_cardinality(PyThunkObject *l, PyThunkObject *r) {
  // thunk_as_number.cpp:60-65 meant this check but compared against the enum (Q9)
  if (l->card.n > 1 && r->card.n > 1 && l->card.kind == CardinalityKind::Exact &&
      r->card.kind == CardinalityKind::Exact && l->card.n != r->card.n) {
    PyErr_SetString(PyExc_TypeError, "Incompatible cardinality.");
    return -1;
  }
  return 0;
}

static PyObject *lazy_binary(PyObject *v, PyObject *w, int opcode, const char *opname) {
  if (!PyThunk_Check(v) && !PyThunk_Check(w)) Py_RETURN_NOTIMPLEMENTED;
  PyObject *lo = PyThunk_Coerce(v);
  if (!lo) return NULL;
  PyObject *ro = PyThunk_Coerce(w);
  if (!ro) { Py_DECREF(lo); return NULL; }
  PyThunkObject *left = (PyThunkObject *)lo, *right = (PyThunkObject *)ro;
  int type;
  if (binary_value_type(left, right, &type) < 0 || check_cardinality(left, right) < 0) {
    Py_DECREF(lo); Py_DECREF(ro);
    return NULL;
  }
  ThunkOperation *op = NewBinaryOperation(lo, ro, OpClass::Vectorizable, opcode, opname);
  Py_DECREF(lo); Py_DECREF(ro);      // the operation holds its own references
  if (!op) return PyErr_NoMemory();
  return PyThunk_FromOperation(op, cardinality_broadcast, CardinalityKind::Exact, type);
}

static PyObject *thunk_lazymultiply(PyObject *v, PyObject *w) { return lazy_binary(v, w, NL_OP_MUL, "*"); }
static PyObject *thunk_lazyadd(PyObject *v, PyObject *w) { return lazy_binary(v, w, NL_OP_ADD, "+"); }
static PyObject *thunk_lazysubtract(PyObject *v, PyObject *w) { return lazy_binary(v, w, NL_OP_SUB, "-"); }
// Division is not in the reference.  `/` exists for float thunks only — NumPy's int / int
// is a float and this engine keeps the left operand's dtype (Q3); `//` is NumPy's floor_divide.
static PyObject *thunk_lazytruedivide(PyObject *v, PyObject *w) {
  PyObject *r = lazy_binary(v, w, NL_OP_DIV, "/");
  if (r && r != Py_NotImplemented) {
    const int t = ((PyThunkObject *)r)->npy_type;
    if (t != NPY_FLOAT32 && t != NPY_FLOAT64) {
      Py_DECREF(r);
      PyErr_SetString(PyExc_TypeError, "true division of integer thunks would change the dtype; use // or convert to float first");
      return NULL;
    }
  }
  return r;
}
static PyObject *thunk_lazyfloordivide(PyObject *v, PyObject *w) { return lazy_binary(v, w, NL_OP_FLOORDIV, "//"); }
// -a == a * (-1): exact for floats (sign flip, also of zeros) and wrapping for ints like NumPy
static PyObject *thunk_lazynegative(PyObject *v) {
  PyThunkObject *l = (PyThunkObject *)v;
  if (l->npy_type == NPY_BOOL) {
    PyErr_SetString(PyExc_TypeError, "unary - is not defined on boolean thunks (use ~)");
    return NULL;
  }
  npy_intp one = 1;
  PyObject *m = PyArray_SimpleNew(1, &one, l->npy_type);
  if (!m) return NULL;
  PyObject *minus_one = PyLong_FromLong(-1);
  const int rc = minus_one ? PyArray_SETITEM((PyArrayObject *)m, (char *)PyArray_DATA((PyArrayObject *)m), minus_one) : -1;
  Py_XDECREF(minus_one);
  if (rc < 0) { Py_DECREF(m); return NULL; }
  PyObject *r = lazy_binary(v, m, NL_OP_MUL, "neg");
  Py_DECREF(m);
  return r;
}

PyObject *PyThunk_LazyRichCompare(PyObject *self, PyObject *other, int cmp_op) {
  if (!PyThunk_Check(self)) {
    PyErr_SetString(PyExc_ValueError, "Expected a thunk object.");
    return NULL;
  }
  int opcode;
  const char *name;
  switch (cmp_op) {
    case Py_GE: opcode = NL_OP_GE; name = ">="; break;   // the one the reference has (thunk_as_number.cpp:21)
    case Py_GT: opcode = NL_OP_GT; name = ">"; break;
    case Py_LE: opcode = NL_OP_LE; name = "<="; break;
    case Py_LT: opcode = NL_OP_LT; name = "<"; break;
    case Py_EQ: opcode = NL_OP_EQ; name = "=="; break;
    default:    opcode = NL_OP_NE; name = "!="; break;
  }
  PyObject *ro = PyThunk_Coerce(other);
  if (!ro) return NULL;
  PyThunkObject *left = (PyThunkObject *)self, *right = (PyThunkObject *)ro;
  int type;
  if (binary_value_type(left, right, &type) < 0 || check_cardinality(left, right) < 0) {
    Py_DECREF(ro);
    return NULL;
  }
  ThunkOperation *op = NewBinaryOperation(self, ro, OpClass::Vectorizable, opcode, name);
  Py_DECREF(ro);
  if (!op) return PyErr_NoMemory();
  return PyThunk_FromOperation(op, cardinality_broadcast, CardinalityKind::Exact,
                               NPY_BOOL /* returns a boolean array */);
}

// boolean thunks: & | ^ ~ combine masks inside the same fused pipeline
static PyObject *lazy_mask_binary(PyObject *v, PyObject *w, int opcode, const char *opname) {
  if (!PyThunk_Check(v) || !PyThunk_Check(w)) Py_RETURN_NOTIMPLEMENTED;
  PyThunkObject *l = (PyThunkObject *)v, *r = (PyThunkObject *)w;
  if (l->npy_type != NPY_BOOL || r->npy_type != NPY_BOOL) {
    PyErr_SetString(PyExc_TypeError, "&, | and ^ are defined on boolean thunks only");
    return NULL;
  }
  if (check_cardinality(l, r) < 0) return NULL;
  ThunkOperation *op = NewBinaryOperation(v, w, OpClass::Vectorizable, opcode, opname);
  if (!op) return PyErr_NoMemory();
  return PyThunk_FromOperation(op, cardinality_broadcast, CardinalityKind::Exact, NPY_BOOL);
}
static PyObject *thunk_lazyand(PyObject *v, PyObject *w) { return lazy_mask_binary(v, w, NL_OP_AND, "&"); }
static PyObject *thunk_lazyor(PyObject *v, PyObject *w) { return lazy_mask_binary(v, w, NL_OP_OR, "|"); }
static PyObject *thunk_lazyxor(PyObject *v, PyObject *w) { return lazy_mask_binary(v, w, NL_OP_XOR, "^"); }
static PyObject *thunk_lazyinvert(PyObject *v) {
  PyThunkObject *l = (PyThunkObject *)v;
  if (l->npy_type != NPY_BOOL) {
    PyErr_SetString(PyExc_TypeError, "~ is defined on boolean thunks only");
    return NULL;
  }
  ThunkOperation *op = NewUnaryOperation(v, OpClass::Vectorizable, NL_OP_NOT, NULL, "~");
  if (!op) return PyErr_NoMemory();
  return PyThunk_FromOperation(op, cardinality_same_as_operand, CardinalityKind::Exact, NPY_BOOL);
}

// Python 3 layout of the number protocol (the reference's table is the Python 2 one,
// thunk_as_number.cpp:83-123); every slot not named here stays NULL as in the reference.
PyNumberMethods thunk_as_number = {
    (binaryfunc)thunk_lazyadd,       /* nb_add */
    (binaryfunc)thunk_lazysubtract,  /* nb_subtract */
    (binaryfunc)thunk_lazymultiply,  /* nb_multiply */
    0,                               /* nb_remainder */
    0,                               /* nb_divmod */
    0,                               /* nb_power */
    (unaryfunc)thunk_lazynegative,   /* nb_negative */
    0,                               /* nb_positive */
    0,                               /* nb_absolute */
    0,                               /* nb_bool */
    (unaryfunc)thunk_lazyinvert,     /* nb_invert */
    0,                               /* nb_lshift */
    0,                               /* nb_rshift */
    (binaryfunc)thunk_lazyand,       /* nb_and */
    (binaryfunc)thunk_lazyxor,       /* nb_xor */
    (binaryfunc)thunk_lazyor,        /* nb_or */
    0,                               /* nb_int */
    0,                               /* nb_reserved */
    0,                               /* nb_float */
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0,    /* nb_inplace_add .. nb_inplace_or */
    (binaryfunc)thunk_lazyfloordivide,  /* nb_floor_divide */
    (binaryfunc)thunk_lazytruedivide,   /* nb_true_divide */
};
