// gencode.hpp — the per-operator plugin record (reference: gencode.hpp:15-41).  The record
// keeps its shape; the LLVM emitter pointers (`gencode`, `initialize`) are replaced by the
// op-code the kernel library understands, and `initialize_data` / `finalize_data` by the
// reduction role implied by that op-code (partials + NCCL finish live behind the C-ABI).
#ifndef NLJIT_GENCODE_H
#define NLJIT_GENCODE_H

#include "config.hpp"

// NumPy fallback of a full pipeline breaker: host ndarray in, new host ndarray out
typedef PyObject *(*base_function_unary)(PyObject *);

typedef enum {
  gentype_unknown = 255,
  gentype_noargs = 0,
  gentype_unary = 1,
  gentype_binary = 2
} gencode_type;

struct GenCodeInfo {
  gencode_type type;
  int opcode;                 /* nl_op executed inside the fused kernel; -1: none (base function only) */
  base_function_unary base;   /* the NumPy function to call when there is no kernel op-code */
  PyObject *parameter[2];
};

typedef enum {
  optype_unknown,
  optype_vectorizable,     /* executed inside the fused loop */
  optype_fullbreaker,      /* everything is materialised and an external (NumPy) function runs */
  optype_parallelbreaker,  /* breaks pipelines but is still computed by the kernel library */
  optype_constant,
  optype_npyarray
} operator_type;

#endif
