// nl_dynamic.cuh — the interpreter-driven instantiations of one dtype (included by nl_dynamic_<dtype>.cu
// with NL_DYN_T / NL_DYN_SUFFIX defined).
#pragma once
#include "nl_compact_pipe.cuh"
#include "nl_compact_super.cuh"
#include "nl_pipeline.cuh"

namespace nl {

#define NL_DYN_CAT2(a, b) a##b
#define NL_DYN_CAT(a, b) NL_DYN_CAT2(a, b)

template <typename T, class Prog>
static pipeline_fn dyn_pick_T(bool vec16, bool compact, bool reduce) {
  constexpr int V16 = wide_v<T>();
  if (vec16) {
    if (compact) return reduce ? pipeline_kernel<T, V16, kRows, true, true, Prog> : pipeline_kernel<T, V16, kRows, true, false, Prog>;
    return reduce ? pipeline_kernel<T, V16, kRows, false, true, Prog> : pipeline_kernel<T, V16, kRows, false, false, Prog>;
  }
  if (compact) return reduce ? pipeline_kernel<T, 1, kRows, true, true, Prog> : pipeline_kernel<T, 1, kRows, true, false, Prog>;
  return reduce ? pipeline_kernel<T, 1, kRows, false, true, Prog> : pipeline_kernel<T, 1, kRows, false, false, Prog>;
}

pipeline_fn NL_DYN_CAT(dynamic_kernel_, NL_DYN_SUFFIX)(bool vec16, bool compact, bool reduce, bool uses_temps) {
  return uses_temps ? dyn_pick_T<NL_DYN_T, DynProg>(vec16, compact, reduce) : dyn_pick_T<NL_DYN_T, DynProgLight>(vec16, compact, reduce);
}

pipeline_fn NL_DYN_CAT(dynamic_pipe_, NL_DYN_SUFFIX)(int rows) {
  if constexpr (sizeof(NL_DYN_T) >= 4) {
    constexpr int V16 = 16 / sizeof(NL_DYN_T);
    if (rows == kPipeRowsNarrow) return compact_pipe_kernel<NL_DYN_T, V16, DynSplit, kPipeRowsNarrow>;
    return compact_pipe_kernel<NL_DYN_T, V16, DynSplit>;
  } else {
    (void)rows;
    return nullptr;
  }
}

pipeline_fn NL_DYN_CAT(dynamic_super_, NL_DYN_SUFFIX)(bool uses_temps) {
  constexpr int V16 = wide_v<NL_DYN_T>();
  return uses_temps ? compact_super_kernel<NL_DYN_T, V16, kRows, DynSplit, true, kMinBlocks>
                    : compact_super_kernel<NL_DYN_T, V16, kRows, DynSplit, false, kMinBlocksLight>;
}

}  // namespace nl
