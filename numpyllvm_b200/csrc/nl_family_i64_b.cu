// nl_family_i64_b.cu — build-time instantiations of the shape families of nl_family.cuh for long long (second half: compaction).
#define NL_STATIC_T long long
#define NL_STATIC_DTYPE NL_I64
#define NL_FAMILY_SECOND_HALF
#include "nl_family.cuh"

namespace nl {
NL_DEFINE_FAMILY_TABLE(family_table_i64_b)
}  // namespace nl
