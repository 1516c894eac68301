// nl_family_f64_b.cu — build-time instantiations of the shape families of nl_family.cuh for double (second half: compaction).
#define NL_STATIC_T double
#define NL_STATIC_DTYPE NL_F64
#define NL_FAMILY_SECOND_HALF
#include "nl_family.cuh"

namespace nl {
NL_DEFINE_FAMILY_TABLE(family_table_f64_b)
}  // namespace nl
