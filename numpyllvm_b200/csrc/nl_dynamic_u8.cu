// nl_dynamic_u8.cu — interpreter-driven kernels (plain, super-tile) for unsigned char.
#define NL_DYN_T unsigned char
#define NL_DYN_SUFFIX u8
#include "nl_dynamic.cuh"
