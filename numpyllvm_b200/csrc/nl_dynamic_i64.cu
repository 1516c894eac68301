// nl_dynamic_i64.cu — interpreter-driven kernels (plain, staged pipeline, super-tile) for long long.
#define NL_DYN_T long long
#define NL_DYN_SUFFIX i64
#include "nl_dynamic.cuh"
