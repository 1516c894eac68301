// scheduler.hpp — runs a tree of pipelines on the devices (reference: scheduler.hpp:12-54).
// The reference's task queue, worker thread and compile/execute tasks collapse into
// stream-ordered kernel launches: children first, one launch per device shard, then the
// finishing step (NCCL allreduce for reductions, count read-back for filters).
#ifndef NLJIT_SCHEDULER_H
#define NLJIT_SCHEDULER_H

#include "parser.hpp"

// executes the pipeline tree; 0 on success, -1 with a Python exception set.  On return every
// output thunk of every pipeline is evaluated and the devices are idle.
int ScheduleExecution(Pipeline *pipeline);
// parse + execute a thunk, splitting over-large expressions (what PyThunk_Evaluate calls)
int ScheduleThunk(PyThunkObject *thunk);
// evaluates `where` (an NL_OP_WHERE_STORE thunk) with its output aliasing `target`'s storage
int ScheduleInPlace(PyThunkObject *where, PyThunkObject *target);

void PyThunk_MarkEvaluated(PyThunkObject *thunk, Storage *storage);

#endif
