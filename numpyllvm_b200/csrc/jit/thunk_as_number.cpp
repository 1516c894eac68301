// thunk_as_number.cpp — lazy arithmetic and comparisons (reference: thunk_as_number.cpp:9-128).
// The reference implements `*` (nb_multiply) and `>=` only; `+`, `-` and the other five
// comparisons are filled in on the same slots with the same shape.  Instead of
// an LLVM emitter each operation records the op-code the kernel library executes.
#include "thunk.hpp"

static bool is_float_type(int t) { return t == NPY_FLOAT32 || t == NPY_FLOAT64; }

// ---- operand types ------------------------------------------------------------------------------
// The reference types the result after the LEFT operand (thunk_as_number.cpp:78, "todo: correct type")
// and would hand LLVM mismatched operand widths (Q3).  Here the result type is NumPy's:
//   two columns (or NumPy scalars)   numpy.promote_types of the two dtypes
//   a Python int                     weak: the column's dtype (OverflowError if it does not fit, as in NumPy 2)
//   a Python float                   weak: the column's dtype if that is a float, float64 otherwise
//   /                                a float type: the promoted type, or float64 for integer operands
// and an operand of another type is converted first: a host scalar on the host, a column by a lazy
// astype() — a pass of its own (nl_cast) whose result the fused pipeline reads in the common type.
struct Weak { bool is_int, is_float; };

static Weak weakness(PyObject *obj) {
  Weak w = {false, false};
  if (PyThunk_Check(obj) || PyArray_Check(obj) || PyArray_IsScalar(obj, Generic)) return w;
  if (PyBool_Check(obj)) return w;
  if (PyLong_Check(obj)) w.is_int = true;
  else if (PyFloat_Check(obj)) w.is_float = true;
  return w;
}

static int result_type(PyThunkObject *l, Weak lw, PyThunkObject *r, Weak rw, int opcode, int *type_out) {
  if (l->npy_type == NPY_BOOL || r->npy_type == NPY_BOOL) {
    PyErr_SetString(PyExc_TypeError, "arithmetic and comparisons on boolean thunks are not supported");
    return -1;
  }
  const bool l_weak = lw.is_int || lw.is_float, r_weak = rw.is_int || rw.is_float;
  int t;
  if (l_weak != r_weak) {
    const int other = l_weak ? r->npy_type : l->npy_type;
    const Weak w = l_weak ? lw : rw;
    t = w.is_int ? other : (is_float_type(other) ? other : NPY_FLOAT64);
  } else if (l->npy_type == r->npy_type) {
    t = l->npy_type;
  } else {
    PyArray_Descr *a = PyArray_DescrFromType(l->npy_type), *b = PyArray_DescrFromType(r->npy_type);
    PyArray_Descr *c = (a && b) ? PyArray_PromoteTypes(a, b) : NULL;
    Py_XDECREF(a);
    Py_XDECREF(b);
    if (!c) return -1;
    t = c->type_num;
    Py_DECREF(c);
  }
  if (opcode == NL_OP_DIV && !is_float_type(t)) t = NPY_FLOAT64;      // int / int is a float64 in NumPy
  if (nljit_npy_to_nl(t) < 0 || t == NPY_BOOL) {
    PyErr_SetString(PyExc_TypeError, "the promoted dtype of the operands is outside the kernel library");
    return -1;
  }
  *type_out = t;
  return 0;
}

// a size-1 thunk of dtype `t` holding the Python scalar `value` (NumPy's own conversion: OverflowError when a
// Python int does not fit the column type)
static PyObject *scalar_thunk_of(PyObject *value, int t) {
  npy_intp one = 1;
  PyObject *arr = PyArray_SimpleNew(1, &one, t);
  if (!arr) return NULL;
  if (PyArray_SETITEM((PyArrayObject *)arr, (char *)PyArray_DATA((PyArrayObject *)arr), value) < 0) { Py_DECREF(arr); return NULL; }
  PyObject *th = PyThunk_FromArray(NULL, arr);
  Py_DECREF(arr);
  return th;
}

// new reference to `operand` as a thunk of dtype `t`
static PyObject *conform(PyObject *original, PyThunkObject *operand, Weak w, int t) {
  if (w.is_int || w.is_float) return scalar_thunk_of(original, t);
  if (operand->npy_type == t) { Py_INCREF(operand); return (PyObject *)operand; }
  return PyThunk_AsType(operand, t);
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
  if (result_type(left, weakness(v), right, weakness(w), opcode, &type) < 0 || check_cardinality(left, right) < 0) {
    Py_DECREF(lo); Py_DECREF(ro);
    return NULL;
  }
  PyObject *lc = conform(v, left, weakness(v), type);
  PyObject *rc = lc ? conform(w, right, weakness(w), type) : NULL;
  Py_DECREF(lo); Py_DECREF(ro);
  if (!lc || !rc) { Py_XDECREF(lc); Py_XDECREF(rc); return NULL; }
  ThunkOperation *op = NewBinaryOperation(lc, rc, OpClass::Vectorizable, opcode, opname);
  Py_DECREF(lc); Py_DECREF(rc);      // the operation holds its own references
  if (!op) return PyErr_NoMemory();
  return PyThunk_FromOperation(op, cardinality_broadcast, CardinalityKind::Exact, type);
}

static PyObject *thunk_lazymultiply(PyObject *v, PyObject *w) { return lazy_binary(v, w, NL_OP_MUL, "*"); }
static PyObject *thunk_lazyadd(PyObject *v, PyObject *w) { return lazy_binary(v, w, NL_OP_ADD, "+"); }
static PyObject *thunk_lazysubtract(PyObject *v, PyObject *w) { return lazy_binary(v, w, NL_OP_SUB, "-"); }
// Division is not in the reference.  `/` is NumPy's true_divide: the result is a float type (integer
// operands are converted to float64 first); `//` is NumPy's floor_divide.
static PyObject *thunk_lazytruedivide(PyObject *v, PyObject *w) { return lazy_binary(v, w, NL_OP_DIV, "/"); }
static PyObject *thunk_lazyfloordivide(PyObject *v, PyObject *w) { return lazy_binary(v, w, NL_OP_FLOORDIV, "//"); }
// floats: -a == a * (-1), exact (sign flip, also of zeros); integers: -a == 0 - a, wrapping like NumPy (also unsigned)
static PyObject *thunk_lazynegative(PyObject *v) {
  PyThunkObject *l = (PyThunkObject *)v;
  if (l->npy_type == NPY_BOOL) {
    PyErr_SetString(PyExc_TypeError, "unary - is not defined on boolean thunks (use ~)");
    return NULL;
  }
  npy_intp one = 1;
  PyObject *m = PyArray_SimpleNew(1, &one, l->npy_type);
  if (!m) return NULL;
  const bool flt = is_float_type(l->npy_type);
  PyObject *k = PyLong_FromLong(flt ? -1 : 0);
  const int rc = k ? PyArray_SETITEM((PyArrayObject *)m, (char *)PyArray_DATA((PyArrayObject *)m), k) : -1;
  Py_XDECREF(k);
  if (rc < 0) { Py_DECREF(m); return NULL; }
  PyObject *r = flt ? lazy_binary(v, m, NL_OP_MUL, "neg") : lazy_binary(m, v, NL_OP_SUB, "neg");
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
  const Weak none = {false, false};
  if (result_type(left, none, right, weakness(other), opcode, &type) < 0 || check_cardinality(left, right) < 0) {
    Py_DECREF(ro);
    return NULL;
  }
  PyObject *lc = conform(self, left, none, type);
  PyObject *rc = lc ? conform(other, right, weakness(other), type) : NULL;
  Py_DECREF(ro);
  if (!lc || !rc) { Py_XDECREF(lc); Py_XDECREF(rc); return NULL; }
  ThunkOperation *op = NewBinaryOperation(lc, rc, OpClass::Vectorizable, opcode, name);
  Py_DECREF(lc); Py_DECREF(rc);
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
