// operation.cpp — operator records (reference: operation.cpp:4-42).  Operations own a
// reference to each parameter (operation.cpp:15, 31-32).
#include "operation.hpp"

#include <cstdlib>
#include <cstring>

ThunkOperation *ThunkOperation_FromUnary(PyObject *arg, operator_type type, int opcode,
                                         base_function_unary base, const char *opname) {
  ThunkOperation *op = (ThunkOperation *)calloc(1, sizeof(ThunkOperation));
  if (!op) return NULL;
  op->type = type;
  op->gencode.type = gentype_unary;
  op->gencode.opcode = opcode;
  op->gencode.base = base;
  op->gencode.parameter[0] = arg;
  op->gencode.parameter[1] = NULL;
  op->opname = strdup(opname);
  Py_INCREF(arg);
  return op;
}

ThunkOperation *ThunkOperation_FromBinary(PyObject *left, PyObject *right, operator_type type, int opcode,
                                          const char *opname) {
  ThunkOperation *op = (ThunkOperation *)calloc(1, sizeof(ThunkOperation));
  if (!op) return NULL;
  op->type = type;
  op->gencode.type = gentype_binary;
  op->gencode.opcode = opcode;
  op->gencode.base = NULL;
  op->gencode.parameter[0] = left;
  op->gencode.parameter[1] = right;
  op->opname = strdup(opname);
  Py_INCREF(left);
  Py_INCREF(right);
  return op;
}

void ThunkOperation_Destroy(ThunkOperation *operation) {
  if (!operation) return;
  for (int i = 0; i < (int)operation->gencode.type && i < 2; i++) Py_XDECREF(operation->gencode.parameter[i]);
  free(operation->opname);
  free(operation);
}
