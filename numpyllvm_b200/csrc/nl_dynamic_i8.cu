// nl_dynamic_i8.cu — interpreter-driven kernels (plain, super-tile) for signed char.
#define NL_DYN_T signed char
#define NL_DYN_SUFFIX i8
#include "nl_dynamic.cuh"
