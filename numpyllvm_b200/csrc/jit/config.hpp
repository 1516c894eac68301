// config.hpp — common includes of the `jit` extension (reference: config.hpp:1-12, ported to
// CPython 3.12 and the NumPy 2 C-API).
#ifndef NLJIT_CONFIG_H
#define NLJIT_CONFIG_H

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
