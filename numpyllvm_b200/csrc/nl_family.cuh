// nl_family.cuh — shape FAMILIES instantiated at build time (second tier of kernel-template selection:
// exact shapes of nl_static.cuh first, then these, then the step interpreter).
//
// A family fixes the KIND of every step (nl_step.h: family_key) and leaves operators and input
// indices to the launch, so one instantiation serves every thunk expression of that shape:
// `(a*2+1)[a > t]`, `(a-b/2)[c != d]`, `(x/y - z)[x <= 0]` all run `fc_v2` (filter, load, two
// operators, compacted store) at the speed of a build-time shape — the warp-uniform operator switch
// sits outside the element loop.  Compaction families get the super-tile kernel, too.
//
// To add a family: one X(...) line; tests/test_abi_host.py lists expressions per family.
#pragma once
#include "nl_compact_pipe.cuh"
#include "nl_compact_super.cuh"
#include "nl_static.cuh"

namespace nl {
namespace fm {
// family keys: which input / operator a step uses is NOT part of the shape; for BIN / CMP the operand kind is
// (nl_step.h: FamilyOperand) — B / C: scalar operand, Bc / Cc: column operand, Bx / Cx: decided per launch
constexpr uint32_t L = make_step(S_LOAD);
constexpr uint32_t B = make_step(S_BIN, 0, FK_SCALAR), Bc = make_step(S_BIN, 0, FK_COLUMN), Bx = make_step(S_BIN, 0, FK_EITHER);
constexpr uint32_t C = make_step(S_CMP, 0, FK_SCALAR), Cc = make_step(S_CMP, 0, FK_COLUMN), Cx = make_step(S_CMP, 0, FK_EITHER);
constexpr uint32_t BT0 = make_step(S_BIN, 0, SRC_TEMP | 0);     // operand: value temporary 0
constexpr uint32_t SV0 = make_step(S_SAVE, 0);
constexpr uint32_t F = make_step(S_FILTER);
constexpr uint32_t PSV0 = make_step(S_PSAVE, 0);
constexpr uint32_t PB0 = make_step(S_PBIN, 0, 0);
constexpr uint32_t ST = make_step(S_STORE, 0, NL_IDX_LOOP);
constexpr uint32_t STC = make_step(S_STORE, 0, NL_IDX_COMPACT);
constexpr uint32_t PST = make_step(S_PSTORE, 0, NL_IDX_LOOP);
constexpr uint32_t R = make_step(S_REDUCE, 0, 0);
constexpr uint32_t W = make_step(S_WHERE, 0);

template <uint32_t... K> constexpr bool uses_temp() { return ((step_op(K) == S_SAVE) || ...); }
// a step that may meet a column operand keeps two register sets alive
template <uint32_t... K> constexpr bool column_operand() {
  return (((step_op(K) == S_BIN || step_op(K) == S_CMP) && !(step_b(K) & SRC_TEMP) && step_b(K) != FK_SCALAR) || ...);
}
template <uint32_t... K> constexpr bool wide() { return !uses_temp<K...>() && !column_operand<K...>(); }
template <uint32_t... K> constexpr int operand_steps() { return (0 + ... + ((step_op(K) == S_BIN || step_op(K) == S_CMP) ? 1 : 0)); }
template <uint32_t... K> constexpr int rows_for() {
  // streaming families: 8 rows per thread while at most two steps may meet a column operand (each keeps a
  // register set in reserve; 128 registers at 2 CTAs/SM), 4 rows beyond that
  if (!sp::has_compact<K...>() && !sp::has_filter<K...>()) return (uses_temp<K...>() || operand_steps<K...>() > 2) ? kRows : kRowsStream;
  return (sp::has_compact<K...>() && wide<K...>()) ? kRowsCompact : kRows;
}
// compaction families on the super-tile kernel: 8 rows per thread when every operand is a scalar (64 registers
// at 4 CTAs/SM hold one register set of 8 rows), 4 rows otherwise
template <uint32_t... K> constexpr int super_rows_for() { return wide<K...>() ? kRowsCompact : kRows; }
template <typename T, bool kOn, uint32_t... K> struct SuperFor { static constexpr pipeline_fn fn = nullptr; };
template <typename T, uint32_t... K> struct SuperFor<T, true, K...> {
  static constexpr pipeline_fn fn = &compact_super_kernel<T, 16 / sizeof(T), super_rows_for<K...>(), FamilySplit<K...>,
                                                          uses_temp<K...>(), (uses_temp<K...>() ? kMinBlocks : kMinBlocksCompact)>;
};
}  // namespace fm

// X(name, keys...)
#define NL_FAMILY_SHAPES_A(X)                                                                       \
  /* elementwise chains: a op b [op c [op d ...]], operands columns or scalars */                    \
  X(ew1, fm::L, fm::Bx, fm::ST)                                                                      \
  X(ew2, fm::L, fm::Bx, fm::Bx, fm::ST)                                                              \
  X(ew3, fm::L, fm::Bx, fm::Bx, fm::Bx, fm::ST)                                                      \
  X(ew4, fm::L, fm::Bx, fm::Bx, fm::Bx, fm::Bx, fm::ST)                                              \
  X(ew5, fm::L, fm::Bx, fm::Bx, fm::Bx, fm::Bx, fm::Bx, fm::ST)                                      \
  /* (a op b) op (c op d), (a op b op c) op (d op e op f): two sub-chains joined through a temporary */ \
  X(ew1x1, fm::L, fm::Bx, fm::SV0, fm::L, fm::Bx, fm::BT0, fm::ST)                                   \
  X(ew2x1, fm::L, fm::Bx, fm::Bx, fm::SV0, fm::L, fm::Bx, fm::BT0, fm::ST)                           \
  X(ew2x2, fm::L, fm::Bx, fm::Bx, fm::SV0, fm::L, fm::Bx, fm::Bx, fm::BT0, fm::ST)                   \
  /* masks: a cmp b, (a op b) cmp c, (a cmp s1) & (a cmp s2) */                                       \
  X(mask1, fm::L, fm::Cx, fm::PST)                                                                   \
  X(mask_v1, fm::L, fm::Bx, fm::Cx, fm::PST)                                                         \
  X(mask_range, fm::L, fm::Cx, fm::PSV0, fm::Cx, fm::PB0, fm::PST)                                   \
  /* in-place masked assignment a[a cmp s] = v */                                                    \
  X(where1, fm::L, fm::Cx, fm::L, fm::W)                                                             \
  /* reductions of chains */                                                                         \
  X(red0, fm::L, fm::R)                                                                              \
  X(red1, fm::L, fm::Bx, fm::R)                                                                      \
  X(red2, fm::L, fm::Bx, fm::Bx, fm::R)                                                              \
  X(red3, fm::L, fm::Bx, fm::Bx, fm::Bx, fm::R)                                                      \
  /* filter -> reduce (the filter is a predicate on the reduction, nothing is compacted) */           \
  X(fr, fm::L, fm::Cx, fm::F, fm::R)                                                                 \
  X(fr_v1, fm::L, fm::Cx, fm::F, fm::Bx, fm::R)                                                      \
  X(fr_v2, fm::L, fm::Cx, fm::F, fm::Bx, fm::Bx, fm::R)                                              \
  X(fr_l, fm::L, fm::Cx, fm::F, fm::L, fm::R)                                                        \
  X(fr_l_v1, fm::L, fm::Cx, fm::F, fm::L, fm::Bx, fm::R)                                             \
  X(fr_p1_l, fm::L, fm::Bx, fm::Cx, fm::F, fm::L, fm::R)                                             \
  X(fr_range, fm::L, fm::C, fm::PSV0, fm::C, fm::PB0, fm::F, fm::R)

#define NL_FAMILY_SHAPES_B(X)                                                                       \
  /* filter -> compacted store.  `_l`: the value column is (re)loaded after the filter; `_vN`: N operators on the */ \
  /* value; `_pN`: N operators inside the predicate; a `c` marks a COLUMN operand (4-row tiles), no mark: scalars (8) */ \
  X(fc, fm::L, fm::C, fm::F, fm::STC)                                                                \
  X(fc_c, fm::L, fm::Cc, fm::F, fm::STC)                                                             \
  X(fc_v1, fm::L, fm::C, fm::F, fm::B, fm::STC)                                                      \
  X(fc_v2, fm::L, fm::C, fm::F, fm::B, fm::B, fm::STC)                                               \
  X(fc_v3, fm::L, fm::C, fm::F, fm::B, fm::B, fm::B, fm::STC)                                        \
  X(fc_v1c, fm::L, fm::C, fm::F, fm::Bc, fm::STC)                                                    \
  X(fc_l, fm::L, fm::C, fm::F, fm::L, fm::STC)                                                       \
  X(fc_c_l, fm::L, fm::Cc, fm::F, fm::L, fm::STC)                                                    \
  X(fc_l_v1, fm::L, fm::C, fm::F, fm::L, fm::B, fm::STC)                                             \
  X(fc_l_v2, fm::L, fm::C, fm::F, fm::L, fm::B, fm::B, fm::STC)                                      \
  X(fc_l_v3, fm::L, fm::C, fm::F, fm::L, fm::B, fm::B, fm::B, fm::STC)                               \
  X(fc_l_v1c, fm::L, fm::C, fm::F, fm::L, fm::Bc, fm::STC)                                           \
  X(fc_l_v2x, fm::L, fm::Cx, fm::F, fm::L, fm::Bx, fm::Bx, fm::STC)                                  \
  X(fc_p1_l, fm::L, fm::B, fm::C, fm::F, fm::L, fm::STC)                                             \
  X(fc_p2_l, fm::L, fm::B, fm::B, fm::C, fm::F, fm::L, fm::STC)                                      \
  X(fc_p1x_l, fm::L, fm::Bx, fm::Cx, fm::F, fm::L, fm::STC)                                          \
  X(fc_p1_l_v1, fm::L, fm::B, fm::C, fm::F, fm::L, fm::B, fm::STC)                                   \
  X(fc_range, fm::L, fm::C, fm::PSV0, fm::C, fm::PB0, fm::F, fm::STC)                                \
  X(fc_range_l, fm::L, fm::C, fm::PSV0, fm::C, fm::PB0, fm::F, fm::L, fm::STC)                       \
  X(fc_range_l_v1, fm::L, fm::C, fm::PSV0, fm::C, fm::PB0, fm::F, fm::L, fm::B, fm::STC)

#define NL_FAMILY_ROW(name, ...)                                                                            \
  {"family:" #name, &FamilyProg<__VA_ARGS__>::matches, NL_STATIC_DTYPE, true,                                \
   fm::rows_for<__VA_ARGS__>(),                                                                              \
   &pipeline_kernel<NL_STATIC_T, 16 / sizeof(NL_STATIC_T), fm::rows_for<__VA_ARGS__>(),                      \
                    sp::has_compact<__VA_ARGS__>(), sp::has_reduce<__VA_ARGS__>(), FamilyProg<__VA_ARGS__>>,   \
   nullptr, 0,                                                                                               \
   fm::SuperFor<NL_STATIC_T, (sp::has_compact<__VA_ARGS__>() && !sp::has_reduce<__VA_ARGS__>()), __VA_ARGS__>::fn, \
   fm::super_rows_for<__VA_ARGS__>()},

#ifdef NL_FAMILY_SECOND_HALF
#define NL_FAMILY_LIST NL_FAMILY_SHAPES_B
#else
#define NL_FAMILY_LIST NL_FAMILY_SHAPES_A
#endif

#define NL_DEFINE_FAMILY_TABLE(fn_name)                                                                     \
  const StaticEntry *fn_name(int *n) {                                                                      \
    static const StaticEntry table[] = {NL_FAMILY_LIST(NL_FAMILY_ROW)};                                     \
    *n = (int)(sizeof(table) / sizeof(table[0]));                                                           \
    return table;                                                                                           \
  }

}  // namespace nl
