// debug_printer.hpp — reference: debug_printer.hpp:1-17.  The reference prints the pipeline
// tree on every evaluate (thunk.cpp:101); here it is gated by the NL_DEBUG_PLAN environment
// variable and goes to stderr.
#ifndef NLJIT_DEBUG_PRINTER_H
#define NLJIT_DEBUG_PRINTER_H

#include "parser.hpp"

char *getname(const char *base);
bool debug_plan_enabled();
void PrintPipeline(Pipeline *pipeline);
void PrintPlan(Pipeline *pipeline, const nl_plan *plan);
void PrintThunk(PyThunkObject *thunk);

#endif
