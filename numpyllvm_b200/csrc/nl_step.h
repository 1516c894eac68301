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

}  // namespace nl
