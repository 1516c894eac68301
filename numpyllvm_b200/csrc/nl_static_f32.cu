// nl_static_f32.cu — build-time instantiations of the pipeline shapes in nl_static.cuh for float.
#define NL_STATIC_T float
#define NL_STATIC_DTYPE NL_F32
#include "nl_static.cuh"

namespace nl {
NL_DEFINE_STATIC_TABLE(static_table_f32)
}  // namespace nl
