// nl_static_f64.cu — build-time instantiations of the pipeline shapes in nl_static.cuh for double.
#define NL_STATIC_T double
#define NL_STATIC_DTYPE NL_F64
#include "nl_static.cuh"

namespace nl {
NL_DEFINE_STATIC_TABLE(static_table_f64)
}  // namespace nl
