// nl_static_i64.cu — build-time instantiations of the pipeline shapes in nl_static.cuh for long long.
#define NL_STATIC_T long long
#define NL_STATIC_DTYPE NL_I64
#include "nl_static.cuh"

namespace nl {
NL_DEFINE_STATIC_TABLE(static_table_i64)
}  // namespace nl
