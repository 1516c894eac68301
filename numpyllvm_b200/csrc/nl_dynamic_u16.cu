// nl_dynamic_u16.cu — interpreter-driven kernels (plain, super-tile) for unsigned short.
#define NL_DYN_T unsigned short
#define NL_DYN_SUFFIX u16
#include "nl_dynamic.cuh"
