// nl_static_i32_b.cu — second half of the build-time shapes (nl_static.cuh: NL_STATIC_SHAPES_B) for int32_t.
#define NL_STATIC_T int32_t
#define NL_STATIC_DTYPE NL_I32
#define NL_STATIC_SECOND_HALF
#include "nl_static.cuh"

namespace nl {
NL_DEFINE_STATIC_TABLE(static_table_i32_b)
}  // namespace nl
