// pyapi.hpp — the CPython 3.12 / NumPy 2 C-API as every translation unit of the `jit` extension
// sees it (one PY_ARRAY_UNIQUE_SYMBOL, imported in jitpackage.cpp), plus the engine's C ABI.
#ifndef NLJIT_PYAPI_H
#define NLJIT_PYAPI_H

#define PY_SSIZE_T_CLEAN
#include <Python.h>

#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#define PY_ARRAY_UNIQUE_SYMBOL nljit_ARRAY_API
#ifndef NLJIT_IMPORT_ARRAY
#define NO_IMPORT_ARRAY
#endif
#include <numpy/arrayobject.h>

#include <stdint.h>
#include <sys/types.h>

#include "nl_abi.h"

#endif
