// operation.cpp — construction / destruction of operator records (see operation.hpp).
#include "operation.hpp"

static ThunkOperation *make(OpClass klass, int arity, int opcode, HostFunction base, const char *name, PyObject *a, PyObject *b) {
  ThunkOperation *op = new ThunkOperation();
  op->klass = klass;
  op->arity = arity;
  op->opcode = opcode;
  op->base = base;
  op->operand[0] = a;
  op->operand[1] = b;
  op->name = name;
  op->refs = 1;
  Py_XINCREF(a);
  Py_XINCREF(b);
  return op;
}

ThunkOperation *NewUnaryOperation(PyObject *arg, OpClass klass, int opcode, HostFunction base, const char *name) {
  return make(klass, 1, opcode, base, name, arg, NULL);
}

ThunkOperation *NewBinaryOperation(PyObject *left, PyObject *right, OpClass klass, int opcode, const char *name) {
  return make(klass, 2, opcode, NULL, name, left, right);
}

void RetainOperation(ThunkOperation *operation) {
  if (operation) operation->refs++;
}

void DestroyOperation(ThunkOperation *operation) {
  if (!operation) return;
  if (--operation->refs > 0) return;
  Py_XDECREF(operation->operand[0]);
  Py_XDECREF(operation->operand[1]);
  delete operation;
}
