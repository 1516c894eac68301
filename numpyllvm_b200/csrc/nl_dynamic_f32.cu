// nl_dynamic_f32.cu — interpreter-driven kernels (plain, staged pipeline, super-tile) for float.
#define NL_DYN_T float
#define NL_DYN_SUFFIX f32
#include "nl_dynamic.cuh"
