// nl_static.cuh — the pipeline shapes that are instantiated at build time.
//
// Each entry is the exact step sequence nl_plan_compile() emits for a thunk
// expression shape (input kinds — column, immediate, device scalar — stay run-time
// parameters, so `a * b`, `a * 2` and `res * a` with res on the device are all
// "mul2").  A plan whose steps match an entry runs the StaticProg instantiation;
// every other plan runs the dynamic interpreter with identical semantics
// (tests/test_gpu_parity.py runs both drivers on the same programs).
//
// To add a shape: append one X(...) line; tests/test_abi_host.py checks that the
// benchmark expressions resolve to a static kernel.
#pragma once
#include "nl_compact_pipe.cuh"
#include "nl_compact_super.cuh"

namespace nl {
namespace sp {
constexpr uint32_t L(uint32_t k) { return make_step(S_LOAD, k); }
constexpr uint32_t LT(uint32_t j) { return make_step(S_LOADT, j); }
constexpr uint32_t SV(uint32_t j) { return make_step(S_SAVE, j); }
constexpr uint32_t B(uint32_t vop, uint32_t src) { return make_step(S_BIN, vop, src); }
constexpr uint32_t C(uint32_t cop, uint32_t src) { return make_step(S_CMP, cop, src); }
constexpr uint32_t F() { return make_step(S_FILTER); }
constexpr uint32_t ST(uint32_t o, uint32_t space) { return make_step(S_STORE, o, space); }
constexpr uint32_t PST(uint32_t o, uint32_t space) { return make_step(S_PSTORE, o, space); }
constexpr uint32_t PSV(uint32_t j) { return make_step(S_PSAVE, j); }
constexpr uint32_t PB(uint32_t pop, uint32_t j) { return make_step(S_PBIN, pop, j); }
constexpr uint32_t R(uint32_t rop, uint32_t o) { return make_step(S_REDUCE, rop, o); }
constexpr uint32_t W(uint32_t o) { return make_step(S_WHERE, o); }
constexpr uint32_t T0 = SRC_TEMP | 0, T1 = SRC_TEMP | 1;
constexpr uint32_t LOOP = NL_IDX_LOOP, COMPACT = NL_IDX_COMPACT;

template <uint32_t... S> constexpr bool has_compact() {
  return (((step_op(S) == S_STORE || step_op(S) == S_PSTORE) && step_b(S) == NL_IDX_COMPACT) || ...);
}
template <uint32_t... S> constexpr bool has_reduce() { return ((step_op(S) == S_REDUCE) || ...); }
template <uint32_t... S> constexpr bool has_filter() { return ((step_op(S) == S_FILTER) || ...); }
// a shape that re-reads a value temporary keeps a second register set alive: 4 rows, not 8
template <uint32_t... S> constexpr bool reads_temp() {
  return ((step_op(S) == S_LOADT || ((step_op(S) == S_BIN || step_op(S) == S_CMP) && (step_b(S) & SRC_TEMP))) || ...);
}
// columns a shape loads (distinct inputs of its LOAD steps): they are what the staged pipeline keeps in
// a ring slot, and two or more take the narrow 4-row instantiation (half the slot size per column)
template <uint32_t... S> constexpr int loaded_columns() {
  constexpr uint32_t st[] = {S...};
  uint32_t seen = 0;
  for (uint32_t s : st)
    if (step_op(s) == S_LOAD) seen |= 1u << step_a(s);
  int n = 0;
  for (int k = 0; k < NL_MAX_INPUTS; k++) n += (seen >> k) & 1u;
  return n;
}
template <uint32_t... S> constexpr int pipe_rows_for() { return loaded_columns<S...>() > 1 ? kPipeRowsNarrow : kPipeRows; }
// the staged compaction pipeline is instantiated for the compaction shapes only
template <typename T, bool kOn, uint32_t... S> struct PipeFor { static constexpr pipeline_fn fn = nullptr; };
template <typename T, uint32_t... S> struct PipeFor<T, true, S...> {
  static constexpr pipeline_fn fn = &compact_pipe_kernel<T, 16 / sizeof(T), StaticSplit<S...>, pipe_rows_for<S...>()>;
};
// the super-tile compaction kernel: 8 rows per thread at 4 CTAs/SM, or 4 rows at 2 CTAs/SM when the shape
// keeps a value temporary alive
template <uint32_t... S> constexpr int super_rows_for() { return reads_temp<S...>() ? kRows : kRowsCompact; }
template <typename T, bool kOn, uint32_t... S> struct SuperFor { static constexpr pipeline_fn fn = nullptr; };
template <typename T, uint32_t... S> struct SuperFor<T, true, S...> {
  static constexpr pipeline_fn fn = &compact_super_kernel<T, 16 / sizeof(T), super_rows_for<S...>(), StaticSplit<S...>, reads_temp<S...>(),
                                                          (reads_temp<S...>() ? kMinBlocks : kMinBlocksCompact)>;
};
template <uint32_t... S> constexpr int rows_for() {
  // eight 16-byte rows per thread for the streaming shapes too (twice the loads in flight per
  // thread at 2 CTAs/SM: C1 +2 %, C3 +1 %, C2 +0.5 % over four rows at 3-5 CTAs/SM)
  // (not the filter-then-reduce shapes: their keep masks push them over the register budget, C5 -12 %)
  if (!has_compact<S...>() && !has_filter<S...>()) return kRowsStream;
  return (has_compact<S...>() && !reads_temp<S...>()) ? kRowsCompact : kRows;
}
}  // namespace sp

// X(name, steps...)
#define NL_CMP_SHAPES(X, tag, cop)                                                                          \
  /* a[a cmp s]            (Tests/filter.py family, C4) */                                                  \
  X(filter_##tag, sp::L(0), sp::C(cop, 1), sp::F(), sp::ST(0, sp::COMPACT))                       \
  /* a[a cmp s].sum()/max()/min()   (Tests/max.py family) */                                                \
  X(filter_##tag##_sum, sp::L(0), sp::C(cop, 1), sp::F(), sp::R(R_SUM, 0))                        \
  X(filter_##tag##_max, sp::L(0), sp::C(cop, 1), sp::F(), sp::R(R_MAX, 0))                        \
  X(filter_##tag##_min, sp::L(0), sp::C(cop, 1), sp::F(), sp::R(R_MIN, 0))                        \
  /* (a[a cmp s] * k).max()   (C5) */                                                                       \
  X(filter_##tag##_mul_max, sp::L(0), sp::C(cop, 1), sp::F(), sp::B(V_MUL, 2), sp::R(R_MAX, 0))   \
  X(filter_##tag##_mul_sum, sp::L(0), sp::C(cop, 1), sp::F(), sp::B(V_MUL, 2), sp::R(R_SUM, 0))   \
  /* b[a cmp s]: select one column by a predicate on another; inputs (a, b, s) or, as the jit    */        \
  /* extension numbers them (value first), (b, a, s)                                               */        \
  X(filter2_##tag, sp::L(0), sp::C(cop, 2), sp::F(), sp::L(1), sp::ST(0, sp::COMPACT))                       \
  X(filter2r_##tag, sp::L(1), sp::C(cop, 2), sp::F(), sp::L(0), sp::ST(0, sp::COMPACT))                      \
  /* a cmp s  -> materialised mask */                                                                       \
  X(mask_##tag, sp::L(0), sp::C(cop, 1), sp::PST(0, sp::LOOP))

/* a[(a lo_cmp s1) & (a hi_cmp s2)]: the range filter, and the same selecting another column */
#define NL_RANGE_SHAPES(X, tag, c1, c2)                                                                     \
  X(filter_range_##tag, sp::L(0), sp::C(c1, 1), sp::PSV(0), sp::C(c2, 2), sp::PB(P_AND, 0), sp::F(),         \
    sp::ST(0, sp::COMPACT))                                                                                  \
  X(filter_range_##tag##_sum, sp::L(0), sp::C(c1, 1), sp::PSV(0), sp::C(c2, 2), sp::PB(P_AND, 0), sp::F(),   \
    sp::R(R_SUM, 0))

#define NL_STATIC_SHAPES(X)                                                                                 \
  NL_RANGE_SHAPES(X, gt_lt, C_GT, C_LT)                                                                     \
  NL_RANGE_SHAPES(X, ge_lt, C_GE, C_LT)                                                                     \
  NL_RANGE_SHAPES(X, gt_le, C_GT, C_LE)                                                                     \
  NL_RANGE_SHAPES(X, ge_le, C_GE, C_LE)                                                                     \
  /* a * b, a * 2, res * a   (Tests/assignment.py, Tests/max.py:14, C2) */                                  \
  X(mul2, sp::L(0), sp::B(V_MUL, 1), sp::ST(0, sp::LOOP))                                                    \
  X(add2, sp::L(0), sp::B(V_ADD, 1), sp::ST(0, sp::LOOP))                                                    \
  X(sub2, sp::L(0), sp::B(V_SUB, 1), sp::ST(0, sp::LOOP))                                                    \
  /* d * e * f   (Tests/multiplication.py:26) */                                                            \
  X(mul3, sp::L(0), sp::B(V_MUL, 1), sp::B(V_MUL, 2), sp::ST(0, sp::LOOP))                                   \
  /* (d*e*f) * (a*b*c)   (C1) */                                                                            \
  X(mul6, sp::L(0), sp::B(V_MUL, 1), sp::B(V_MUL, 2), sp::SV(0), sp::L(3), sp::B(V_MUL, 4), sp::B(V_MUL, 5), \
    sp::B(V_MUL, sp::T0), sp::ST(0, sp::LOOP))                                                               \
  /* a.max() / a.min() / a.sum()   (C3) */                                                                  \
  X(reduce_max, sp::L(0), sp::R(R_MAX, 0))                                                                   \
  X(reduce_min, sp::L(0), sp::R(R_MIN, 0))                                                                   \
  X(reduce_sum, sp::L(0), sp::R(R_SUM, 0))                                                                   \
  /* a[a * k >= s]   (Tests/filter.py:13 exactly) */                                                        \
  X(filter_mul_ge, sp::L(0), sp::B(V_MUL, 1), sp::C(C_GE, 2), sp::F(), sp::L(0), sp::ST(0, sp::COMPACT)) \
  NL_CMP_SHAPES(X, ge, C_GE)                                                                                \
  NL_CMP_SHAPES(X, gt, C_GT)

// second half of the list: compiled in its own translation unit (nl_static_<dtype>_b.cu) so that the
// build spreads over more cores
#define NL_STATIC_SHAPES_B(X)                                                                               \
  NL_CMP_SHAPES(X, le, C_LE)                                                                                \
  NL_CMP_SHAPES(X, lt, C_LT)                                                                                \
  NL_CMP_SHAPES(X, eq, C_EQ)                                                                                \
  NL_CMP_SHAPES(X, ne, C_NE)

// one table row per shape (128-bit rows; misaligned launches use the dynamic V = 1 kernel)
// compaction shapes run 8 rows per thread (32 KB tiles): fewer, larger tiles keep the look-back
// chain off the critical path; everything else 4 rows
#define NL_STATIC_ROW(name, ...)                                                                            \
  {#name, &StaticProg<__VA_ARGS__>::matches, NL_STATIC_DTYPE, true,                                          \
   sp::rows_for<__VA_ARGS__>(),                                                                              \
   &pipeline_kernel<NL_STATIC_T, 16 / sizeof(NL_STATIC_T), sp::rows_for<__VA_ARGS__>(),                      \
                    sp::has_compact<__VA_ARGS__>(), sp::has_reduce<__VA_ARGS__>(), StaticProg<__VA_ARGS__>>,   \
   /* shapes that load two columns move twice the bytes per tile: the plain kernel already runs them at    */ \
   /* 6.2-6.6 TB/s (the staged pipeline: 5.1-5.9), so only single-column shapes get a pipeline            */ \
   sp::PipeFor<NL_STATIC_T, (sp::has_compact<__VA_ARGS__>() && !sp::has_reduce<__VA_ARGS__>() &&             \
                             sp::loaded_columns<__VA_ARGS__>() == 1), __VA_ARGS__>::fn,                        \
   sp::pipe_rows_for<__VA_ARGS__>(),                                                                         \
   sp::SuperFor<NL_STATIC_T, (sp::has_compact<__VA_ARGS__>() && !sp::has_reduce<__VA_ARGS__>()), __VA_ARGS__>::fn, \
   sp::super_rows_for<__VA_ARGS__>()},

#ifdef NL_STATIC_SECOND_HALF
#define NL_STATIC_LIST NL_STATIC_SHAPES_B
#else
#define NL_STATIC_LIST NL_STATIC_SHAPES
#endif

#define NL_DEFINE_STATIC_TABLE(fn_name)                                                                     \
  const StaticEntry *fn_name(int *n) {                                                                      \
    static const StaticEntry table[] = {NL_STATIC_LIST(NL_STATIC_ROW)};                                     \
    *n = (int)(sizeof(table) / sizeof(table[0]));                                                           \
    return table;                                                                                           \
  }

}  // namespace nl
