// nl_dynamic_i16.cu — interpreter-driven kernels (plain, super-tile) for short.
#define NL_DYN_T short
#define NL_DYN_SUFFIX i16
#include "nl_dynamic.cuh"
