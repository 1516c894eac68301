// nl_family_f32.cu — build-time instantiations of the shape families of nl_family.cuh for float (first half).
#define NL_STATIC_T float
#define NL_STATIC_DTYPE NL_F32
#include "nl_family.cuh"

namespace nl {
NL_DEFINE_FAMILY_TABLE(family_table_f32)
}  // namespace nl
