// nl_static_i64_b.cu — second half of the build-time shapes (nl_static.cuh: NL_STATIC_SHAPES_B) for long long.
#define NL_STATIC_T long long
#define NL_STATIC_DTYPE NL_I64
#define NL_STATIC_SECOND_HALF
#include "nl_static.cuh"

namespace nl {
NL_DEFINE_STATIC_TABLE(static_table_i64_b)
}  // namespace nl
