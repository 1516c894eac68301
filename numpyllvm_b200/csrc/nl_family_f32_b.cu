// nl_family_f32_b.cu — build-time instantiations of the shape families of nl_family.cuh for float (second half: compaction).
#define NL_STATIC_T float
#define NL_STATIC_DTYPE NL_F32
#define NL_FAMILY_SECOND_HALF
#include "nl_family.cuh"

namespace nl {
NL_DEFINE_FAMILY_TABLE(family_table_f32_b)
}  // namespace nl
