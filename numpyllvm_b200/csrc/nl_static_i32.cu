// nl_static_i32.cu — build-time instantiations of the pipeline shapes in nl_static.cuh for int32_t.
#define NL_STATIC_T int32_t
#define NL_STATIC_DTYPE NL_I32
#include "nl_static.cuh"

namespace nl {
NL_DEFINE_STATIC_TABLE(static_table_i32)
}  // namespace nl
