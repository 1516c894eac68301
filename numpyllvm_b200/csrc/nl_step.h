// nl_step.h — encoding of the accumulator-machine program a plan carries.
//
// The reference emits, per pipeline, one scalar loop body by walking the
// operation tree (PerformOperation, compiler.cpp:202-269).  Here the same walk
// is done once on the host (nl_plan.cpp) and produces a short linear program;
// the pipeline kernel interprets it with every value held in registers.
//
// Machine state per thread, for the E = U*V elements it owns in a tile:
//   acc[E]        value accumulator (type T)
//   t0..t1[E]     value temporaries (NL_MAX_TEMPS)
//   pacc, p0..p2  predicate accumulator / temporaries, one bit per element
//   keep          elements still alive (in range and not filtered out)
// All control flow is warp-uniform: the program is the same for every thread.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define NL_HD __host__ __device__
#else
#define NL_HD
#endif

namespace nl {

enum StepOp : uint32_t {
  S_LOAD   = 1,   // a = input k                 acc  <- in[k]
  S_LOADT  = 2,   // a = temp j                  acc  <- t[j]
  S_SAVE   = 3,   // a = temp j                  t[j] <- acc
  S_BIN    = 4,   // a = VOp, b = src            acc  <- vop(acc, X)
  S_CMP    = 5,   // a = COp, b = src            pacc <- cop(acc, X)
  S_PLOAD  = 6,   // a = input k (bool)          pacc <- in[k] != 0
  S_PLOADT = 7,   // a = ptemp j                 pacc <- p[j]
  S_PSAVE  = 8,   // a = ptemp j                 p[j] <- pacc
  S_PBIN   = 9,   // a = POp, b = ptemp j        pacc <- pop(pacc, p[j])
  S_PNOT   = 10,  //                             pacc <- ~pacc
  S_FILTER = 11,  //                             keep &= pacc (+ compaction scan)
  S_STORE  = 12,  // a = output o, b = space     out[o][idx] <- acc
  S_PSTORE = 13,  // a = output o, b = space     out[o][idx] <- pacc (0x00/0x01)
  S_REDUCE = 14,  // a = ROp, b = output o       racc <- rop(racc, acc) under keep
  S_WHERE  = 15   // a = output o                if (pacc) out[o][i] <- acc
};

enum VOp : uint32_t { V_MUL = 0, V_ADD = 1, V_SUB = 2, V_RSUB = 3,              // RSUB: X - acc
                      V_DIV = 4, V_RDIV = 5, V_FDIV = 6, V_RFDIV = 7 };         // true / floor division, R: X op acc
enum COp : uint32_t { C_GE = 0, C_GT = 1, C_LE = 2, C_LT = 3, C_EQ = 4, C_NE = 5 };
enum POp : uint32_t { P_AND = 0, P_OR = 1, P_XOR = 2 };
enum ROp : uint32_t { R_MAX = 0, R_MIN = 1, R_SUM = 2 };

// operand source byte: bit 7 set -> temporary (low bits = j), else input k
constexpr uint32_t SRC_TEMP = 0x80;

NL_HD constexpr inline uint32_t make_step(uint32_t op, uint32_t a = 0, uint32_t b = 0, uint32_t c = 0) {
  return (op & 0xff) | ((a & 0xff) << 8) | ((b & 0xff) << 16) | ((c & 0xff) << 24);
}
NL_HD constexpr inline uint32_t step_op(uint32_t s) { return s & 0xff; }
NL_HD constexpr inline uint32_t step_a(uint32_t s) { return (s >> 8) & 0xff; }
NL_HD constexpr inline uint32_t step_b(uint32_t s) { return (s >> 16) & 0xff; }
NL_HD constexpr inline uint32_t step_c(uint32_t s) { return (s >> 24) & 0xff; }

// ---- shape families -------------------------------------------------------------------------
// A build-time instantiated kernel does not have to fix WHICH operator a step applies or WHICH input
// it reads — only what kind of step it is: the operator is a warp-uniform switch outside the element
// loop and the input an index into the parameter block, neither costs anything at HBM rate.  The
// family key of a step keeps the fields that shape the code (step kind, temporaries, output slot and
// index space) and drops the ones that stay run-time parameters:
//   LOAD / PLOAD     which input
//   BIN / CMP        which operator; which input when the operand is not a temporary — but whether that
//                    input is a scalar or a column IS part of the shape where registers are tight (a
//                    column operand needs a second register set): FK_SCALAR / FK_COLUMN / FK_EITHER
//   PBIN             and / or / xor
//   REDUCE           max / min / sum
// One family therefore serves `a[a > t]`, `a[a <= 3]`, ... and `(a*2+1)[a > t]`, `(a/3-1)[a != 0]`, ...
enum FamilyOperand : uint32_t { FK_SCALAR = 0, FK_COLUMN = 1, FK_EITHER = 2 };
// key of a plan's step; operand_is_column: the step's input operand (BIN / CMP, not a temporary) is an array
NL_HD constexpr inline uint32_t family_key(uint32_t s, bool operand_is_column) {
  switch (step_op(s)) {
    case S_LOAD: case S_PLOAD: return make_step(step_op(s));
    case S_BIN: case S_CMP:
      return (step_b(s) & SRC_TEMP) ? make_step(step_op(s), 0, step_b(s)) : make_step(step_op(s), 0, operand_is_column ? FK_COLUMN : FK_SCALAR);
    case S_PBIN: return make_step(S_PBIN, 0, step_b(s));
    case S_REDUCE: return make_step(S_REDUCE, 0, step_b(s));
    default: return s;
  }
}
// does a plan step with key `have` fit the family's key `want`?
NL_HD constexpr inline bool family_key_fits(uint32_t have, uint32_t want) {
  if (have == want) return true;
  const uint32_t op = step_op(want);
  return (op == S_BIN || op == S_CMP) && step_op(have) == op && !(step_b(want) & SRC_TEMP) && !(step_b(have) & SRC_TEMP) &&
         step_b(want) == FK_EITHER;
}
// operators a family kernel carries inline (floor division is a long out-of-line sequence: interpreter only)
NL_HD constexpr inline bool family_step_ok(uint32_t s) {
  return !(step_op(s) == S_BIN && (step_a(s) == V_FDIV || step_a(s) == V_RFDIV));
}

// Splitting a compaction plan at its filter into a count pass and a store pass (super-tile kernel): do
// the steps BEFORE the filter have to run again in the store pass?  Only if a value they produce is
// still used after it: the accumulator (the first step after the filter is not a load) or a temporary.
NL_HD constexpr inline bool store_pass_needs_prefix(const uint32_t *steps, int n, int filter) {
  if (filter + 1 >= n) return false;
  const uint32_t first = step_op(steps[filter + 1]);
  if (first != S_LOAD && first != S_LOADT) return true;
  for (int i = filter + 1; i < n; i++) {
    const uint32_t op = step_op(steps[i]);
    if (op == S_LOADT) return true;
    if ((op == S_BIN || op == S_CMP) && (step_b(steps[i]) & SRC_TEMP)) return true;
  }
  return false;
}

}  // namespace nl
