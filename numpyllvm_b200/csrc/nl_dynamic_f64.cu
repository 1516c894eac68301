// nl_dynamic_f64.cu — interpreter-driven kernels (plain, staged pipeline, super-tile) for double.
#define NL_DYN_T double
#define NL_DYN_SUFFIX f64
#include "nl_dynamic.cuh"
