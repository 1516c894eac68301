// optimizer.hpp — plan-level optimiser: the stage between the parser and the kernel library.
// (The reference's optimizer.cpp is an LLVM IR pass list and does no plan-level work; with
// LLVM gone this stage is new: it lowers a pipeline's operation tree to the C-ABI's node
// format, folds host scalars into immediates, keeps reduction results on the device, asks
// nl_plan_compile() for the program, and proposes a cut when an expression is too large for
// one fused kernel.  Kernel-template selection itself happens inside the library.)
#ifndef NLJIT_OPTIMIZER_H
#define NLJIT_OPTIMIZER_H

#include "parser.hpp"
#include <vector>

struct LoweredPipeline {
  nl_expr_node nodes[NL_MAX_NODES];
  int n_nodes;
  nl_input_desc inputs[NL_MAX_INPUTS];
  DataSource *input_sources[NL_MAX_INPUTS];
  int n_inputs;
  int n_outputs;
  int dtype;          // nl_dtype of the value registers
  nl_plan plan;
};

// 0: lowered and compiled.  1: the expression needs splitting, `cuts_out` receives the thunk(s) to
// materialise in pipelines of their own — one near the middle of an over-large expression, or every
// filter below the root when a column meets a filter result inside one pipeline (a[m] * b[m]).
// -1: error, Python exception set.
int LowerPipeline(Pipeline *pipeline, LoweredPipeline *out, std::vector<PyThunkObject *> *cuts_out);
// bit pattern of a scalar of nl type `from` converted to nl type `to` (C conversion rules)
uint64_t nljit_convert_scalar(uint64_t bits, int from, int to);

#endif
