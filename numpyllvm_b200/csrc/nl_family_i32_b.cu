// nl_family_i32_b.cu — build-time instantiations of the shape families of nl_family.cuh for int32_t (second half: compaction).
#define NL_STATIC_T int32_t
#define NL_STATIC_DTYPE NL_I32
#define NL_FAMILY_SECOND_HALF
#include "nl_family.cuh"

namespace nl {
NL_DEFINE_FAMILY_TABLE(family_table_i32_b)
}  // namespace nl
