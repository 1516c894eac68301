// nl_family_i64.cu — build-time instantiations of the shape families of nl_family.cuh for long long (first half).
#define NL_STATIC_T long long
#define NL_STATIC_DTYPE NL_I64
#include "nl_family.cuh"

namespace nl {
NL_DEFINE_FAMILY_TABLE(family_table_i64)
}  // namespace nl
