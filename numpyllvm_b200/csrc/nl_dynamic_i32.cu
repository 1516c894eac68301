// nl_dynamic_i32.cu — interpreter-driven kernels (plain, staged pipeline, super-tile) for int32_t.
#define NL_DYN_T int32_t
#define NL_DYN_SUFFIX i32
#include "nl_dynamic.cuh"
