// nl_dynamic_u32.cu — interpreter-driven kernels (plain, super-tile) for unsigned int.
#define NL_DYN_T unsigned int
#define NL_DYN_SUFFIX u32
#include "nl_dynamic.cuh"
