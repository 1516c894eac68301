// nl_family_f64.cu — build-time instantiations of the shape families of nl_family.cuh for double (first half).
#define NL_STATIC_T double
#define NL_STATIC_DTYPE NL_F64
#include "nl_family.cuh"

namespace nl {
NL_DEFINE_FAMILY_TABLE(family_table_f64)
}  // namespace nl
