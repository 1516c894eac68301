// nl_dynamic_u64.cu — interpreter-driven kernels (plain, super-tile) for unsigned long long.
#define NL_DYN_T unsigned long long
#define NL_DYN_SUFFIX u64
#include "nl_dynamic.cuh"
