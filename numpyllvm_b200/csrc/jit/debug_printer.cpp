// debug_printer.cpp — names and plan dumps (reference: debug_printer.cpp:10-178).
#include "debug_printer.hpp"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

static std::atomic<long> g_name_counter{0};   // the reference's counter is not thread-safe (debug_printer.cpp:10-15)

char *getname(const char *base) {
  char buf[64];
  snprintf(buf, sizeof buf, "%s%ld", base, g_name_counter.fetch_add(1));
  return strdup(buf);
}

bool debug_plan_enabled() {
  static int on = -1;
  if (on < 0) {
    const char *e = getenv("NL_DEBUG_PLAN");
    on = (e && *e && strcmp(e, "0") != 0) ? 1 : 0;
  }
  return on == 1;
}

static void print_op(Operation *op, int depth) {
  for (int i = 0; i < depth; i++) fputs("  ", stderr);
  switch (op->kind) {
    case Operation::Column:
      fprintf(stderr, "[%s] size=%zd\n", op->thunk && op->thunk->name ? op->thunk->name : "?", op->thunk ? op->thunk->card.n : (ssize_t)-1);
      break;
    case Operation::ChildResult:
      fprintf(stderr, "<%s>\n", op->child->name);
      break;
    case Operation::Apply1:
      fprintf(stderr, "%s%s\n", op->op->name.c_str(), op->materialize ? "  -> stored" : "");
      print_op(op->lhs, depth + 1);
      break;
    case Operation::Apply2:
      fprintf(stderr, "%s%s\n", op->op->name.c_str(), op->materialize ? "  -> stored" : "");
      print_op(op->lhs, depth + 1);
      print_op(op->rhs, depth + 1);
      break;
    default:
      fputs("?\n", stderr);
  }
}

void PrintPipeline(Pipeline *pipeline) {
  for (Pipeline *c : pipeline->children) PrintPipeline(c);
  fprintf(stderr, "Pipeline %s (%zu inputs, %zu outputs)\n", pipeline->name, pipeline->inputs.size(),
          pipeline->outputs.size());
  print_op(pipeline->root, 1);
}

void PrintPlan(Pipeline *pipeline, const nl_plan *plan) {
  char buf[4096], kernel[160];
  int is_static = 0;
  nl_plan_dump(plan, buf, sizeof buf);
  nl_plan_kernel(plan, 1, kernel, sizeof kernel, &is_static);
  fprintf(stderr, "Plan of %s -> %s\n%s", pipeline->name, kernel, buf);
}

void PrintThunk(PyThunkObject *thunk) {
  if (thunk->evaluated || !thunk->operation) { fprintf(stderr, "%s", thunk->name ? thunk->name : "?"); return; }
  ThunkOperation *op = thunk->operation;
  fputc('(', stderr);
  if (op->arity == 2) {
    PrintThunk((PyThunkObject *)op->operand[0]);
    fprintf(stderr, " %s ", op->name.c_str());
    PrintThunk((PyThunkObject *)op->operand[1]);
  } else {
    fprintf(stderr, "%s ", op->name.c_str());
    PrintThunk((PyThunkObject *)op->operand[0]);
  }
  fputc(')', stderr);
}
