// operation.hpp — reference: operation.hpp:7-27
#ifndef NLJIT_OPERATION_H
#define NLJIT_OPERATION_H

#include "gencode.hpp"

struct ThunkOperation {
  operator_type type;
  GenCodeInfo gencode;
  char *opname;
};

ThunkOperation *ThunkOperation_FromUnary(PyObject *arg, operator_type type, int opcode,
                                         base_function_unary base, const char *opname);
ThunkOperation *ThunkOperation_FromBinary(PyObject *left, PyObject *right, operator_type type, int opcode,
                                          const char *opname);
void ThunkOperation_Destroy(ThunkOperation *operation);

#endif
