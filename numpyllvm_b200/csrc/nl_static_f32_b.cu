// nl_static_f32_b.cu — second half of the build-time shapes (nl_static.cuh: NL_STATIC_SHAPES_B) for float.
#define NL_STATIC_T float
#define NL_STATIC_DTYPE NL_F32
#define NL_STATIC_SECOND_HALF
#include "nl_static.cuh"

namespace nl {
NL_DEFINE_STATIC_TABLE(static_table_f32_b)
}  // namespace nl
