// thunk_as_sequence.cpp — reference: thunk_as_sequence.cpp:5-16 (every slot NULL there).
// sq_length is the one addition: len() of an evaluated thunk, or of a thunk whose cardinality
// is already exact, without forcing evaluation of anything else.
#include "thunk.hpp"

static Py_ssize_t thunk_length(PyObject *self) {
  PyThunkObject *t = (PyThunkObject *)self;
  if (!t->evaluated) {
    PyObject *r = PyThunk_Evaluate(t);    // a filter's length is only known after it ran
    if (!r) return -1;
    Py_DECREF(r);
  }
  return (Py_ssize_t)t->card.n;
}

PySequenceMethods thunk_as_sequence = {
    (lenfunc)thunk_length, /* sq_length */
    0,                     /* sq_concat */
    0,                     /* sq_repeat */
    0,                     /* sq_item */
    0,                     /* was_sq_slice */
    0,                     /* sq_ass_item */
    0,                     /* was_sq_ass_slice */
    0,                     /* sq_contains */
    0,                     /* sq_inplace_concat */
    0,                     /* sq_inplace_repeat */
};
