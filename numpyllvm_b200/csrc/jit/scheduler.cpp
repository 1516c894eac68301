// scheduler.cpp — see scheduler.hpp.  What remains of scheduler.cpp:13-219 and
// compiler.cpp:18-123 once compilation is a table lookup and execution is a kernel launch:
//   ScheduleCompilation / CompilePipeline  -> LowerPipeline() + nl_plan_compile()
//   ScheduleFunction   (alloc + 8 ranges)  -> output shards + one nl_launch_pipeline per device
//   ExecuteFunction    (call the loop)     -> the launch itself
//   JITFunctionDECREF  (gather, finalize)  -> nl_finish_reduce / shard counts, mark evaluated
#include "scheduler.hpp"
#include "optimizer.hpp"
#include "debug_printer.hpp"
#include "device_sort.hpp"

#include <algorithm>
#include <cstring>
#include <vector>

// per-device int64 scratch the kernels write output cardinalities to (compiler.cpp:491-508)
static int64_t *g_sizes[NL_MAX_DEVICES];
static int g_sizes_slot[NL_MAX_DEVICES];
static const int kSizeSlots = 16;

static int64_t *sizes_buffer(int dev) {
  if (!g_sizes[dev]) {
    void *p = NULL;
    if (nl_malloc(dev, sizeof(int64_t) * NL_MAX_OUTPUTS * kSizeSlots, &p) != NL_OK) return NULL;
    g_sizes[dev] = (int64_t *)p;
  }
  g_sizes_slot[dev] = (g_sizes_slot[dev] + 1) % kSizeSlots;
  return g_sizes[dev] + g_sizes_slot[dev] * NL_MAX_OUTPUTS;
}

// per-device scratch for the exclusive scan of the shard counts a fused filter leaves behind (world + 1 entries)
static int64_t *g_offsets[NL_MAX_DEVICES];
static int g_offsets_slot[NL_MAX_DEVICES];

static int64_t *offsets_buffer(int dev) {
  if (!g_offsets[dev]) {
    void *p = NULL;
    if (nl_malloc(dev, sizeof(int64_t) * (NL_MAX_DEVICES + 1) * kSizeSlots, &p) != NL_OK) return NULL;
    g_offsets[dev] = (int64_t *)p;
  }
  g_offsets_slot[dev] = (g_offsets_slot[dev] + 1) % kSizeSlots;
  return g_offsets[dev] + g_offsets_slot[dev] * (NL_MAX_DEVICES + 1);
}

static std::vector<PyThunkObject *> g_pending_cuts;

// the reference blocks on a semaphore while holding the GIL (thunk.cpp:106); here the GIL is released
static int sync_devices() { return nljit_sync_devices(); }

// the full pipeline breaker: materialise the child, run the NumPy base function on the host,
// upload the result (compiler.cpp:93-122)
static int run_base_function(Pipeline *pipeline) {
  ThunkOperation *top = (pipeline->root)->op;
  if (!top || (!top->base && top->name != "astype")) {
    PyErr_SetString(PyExc_ValueError, "pipeline root has neither a kernel op-code nor a base function");
    return -1;
  }
  if (pipeline->inputs.size() != 1) {
    PyErr_SetString(PyExc_ValueError, "Expected 1 parameter for a base function.");
    return -1;
  }
  DataSource *in = pipeline->inputs[0].source;
  if (!in->storage && in->object) {
    if (PyThunk_EnsureDevice(in->object) < 0) return -1;
    in->storage = in->object->storage;
    storage_incref(in->storage);
  }
  if (!in->storage) { PyErr_SetString(PyExc_ValueError, "base function input is not evaluated"); return -1; }
  if (top->name == "astype") {
    // conversion to the output thunk's element type: one nl_cast pass per shard, same layout as the input
    Binding &out = pipeline->outputs[0];
    Storage *src = in->storage;
    const int to = nljit_npy_to_nl(out.source->object->npy_type);
    Storage *dst = storage_new(to, src->bound);
    dst->ragged = src->ragged;
    dst->replicated = src->replicated;
    if (src->host_scalar) {
      dst->host_scalar = true;
      dst->scalar_bits = nljit_convert_scalar(src->scalar_bits, src->dtype, to);
    } else {
      if (nljit_ensure_runtime() < 0 || storage_settle(src) < 0) { storage_decref(dst); return -1; }
      const size_t es = nl_dtype_size(to);
      for (const Shard &sh : src->shards) {
        Shard o = sh;
        o.marks.clear();
        o.ptr = NULL;
        const int64_t cnt = src->replicated ? 1 : sh.count;
        o.capacity = cnt;
        int rc = NL_OK;
        if (cnt > 0) rc = nl_malloc(sh.dev, (size_t)cnt * es, &o.ptr);
        if (rc == NL_OK && cnt > 0) rc = nl_cast(sh.dev, 0, o.ptr, to, sh.ptr, src->dtype, cnt);
        dst->shards.push_back(o);
        if (rc != NL_OK) { nljit_raise(rc, "jit: dtype conversion failed"); storage_decref(dst); return -1; }
      }
    }
    out.source->storage = dst;                          // the pipeline's reference
    out.source->size = (ssize_t)dst->cardinality();
    storage_incref(dst);
    PyThunk_MarkEvaluated(out.source->object, dst);     // the thunk holds its own reference
    return 0;
  }
  if (top->name == "sort" && !in->storage->host_scalar && !in->storage->replicated && nljit_ensure_runtime() == 0) {
    // sort() runs where the data lives: radix sort per GPU, sample sort across GPUs (device_sort.cpp)
    Storage *sorted = device_sort(in->storage);
    if (!sorted) return -1;
    Binding &out = pipeline->outputs[0];
    out.source->storage = sorted;                       // the pipeline's reference
    out.source->size = (ssize_t)sorted->cardinality();
    storage_incref(sorted);
    PyThunk_MarkEvaluated(out.source->object, sorted);  // the thunk holds its own reference
    return 0;
  }
  PyObject *host = storage_to_host(in->storage);
  if (!host) return -1;
  PyObject *result = top->base(host);
  Py_DECREF(host);
  if (!result) return -1;
  PyObject *t = PyThunk_FromArray(NULL, result);     // uploads
  Py_DECREF(result);
  if (!t) return -1;
  PyThunkObject *tt = (PyThunkObject *)t;
  if (PyThunk_EnsureDevice(tt) < 0) { Py_DECREF(t); return -1; }
  Binding &out = pipeline->outputs[0];
  storage_incref(tt->storage);
  out.source->storage = tt->storage;
  storage_incref(tt->storage);
  out.source->size = (ssize_t)tt->storage->cardinality();
  PyThunk_MarkEvaluated(out.source->object, tt->storage);
  Py_DECREF(t);
  return 0;
}

static int execute_one(Pipeline *pipeline) {
  if (pipeline->evaluated) return 0;
  if (nljit_ensure_runtime() < 0) return -1;
  if (pipeline->root->thunk && pipeline->root->thunk->evaluated && pipeline->root->thunk->storage &&
      pipeline->root->kind != Operation::Column && !pipeline->inplace_target) {
    // the root was computed meanwhile (by an earlier pipeline of this tree): publish its storage to
    // whoever reads this pipeline's result as a column and skip the work
    for (Binding &e : pipeline->outputs)
      if (e.source->object && e.source->object->storage && !e.source->storage) {
        e.source->storage = e.source->object->storage;
        storage_incref(e.source->storage);
        e.source->size = (ssize_t)e.source->storage->cardinality();
      }
    pipeline->evaluated = true;
    return 0;
  }
  ThunkOperation *root_op = (pipeline->root)->op;
  if (root_op && root_op->opcode < 0) return run_base_function(pipeline);

  // inputs must be resident (deferred upload of host-only thunks)
  for (Binding &e : pipeline->inputs) {
    DataSource *src = e.source;
    if (!src->storage && src->object) {
      if (PyThunk_EnsureDevice(src->object) < 0) return -1;
      src->storage = src->object->storage;
      storage_incref(src->storage);
    }
    if (!src->storage) { PyErr_SetString(PyExc_ValueError, "pipeline input is not evaluated"); return -1; }
    src->size = (ssize_t)src->storage->cardinality();
  }

  // a one-element column that is neither a host scalar nor replicated (e.g. a filter that kept one
  // element) is broadcast like any size-1 input (compiler.cpp:210): fetch it and bake it
  for (Binding &e : pipeline->inputs) {
    Storage *s = e.source->storage;
    if (e.source->size == 1 && !s->host_scalar && !s->replicated) {
      PyObject *host = storage_to_host(s);
      if (!host) return -1;
      Storage *hs = storage_new(s->dtype, 1);
      hs->host_scalar = true;
      memcpy(&hs->scalar_bits, PyArray_DATA((PyArrayObject *)host), nl_dtype_size(s->dtype));
      Py_DECREF(host);
      storage_decref(s);
      e.source->storage = hs;
    }
  }

  LoweredPipeline lp;
  std::vector<PyThunkObject *> cut;
  int lrc = LowerPipeline(pipeline, &lp, &cut);
  if (lrc == 0) {
    // A purely elementwise pipeline can follow an upload that is still in flight chunk by chunk;
    // everything else (and every column on another layout) first orders the compute stream behind it.
    const bool chunkable = !(lp.plan.flags & (NL_PLAN_HAS_FILTER | NL_PLAN_HAS_COMPACT | NL_PLAN_HAS_REDUCE | NL_PLAN_HAS_WHERE));
    for (int k = 0; k < lp.n_inputs; k++) {
      Storage *s = lp.input_sources[k]->storage;
      if (!s || s->producer < 0) continue;
      if (!chunkable || lp.inputs[k].kind != NL_IN_ARRAY || s->ragged) { if (storage_settle(s) < 0) return -1; }
    }
  }
  if (lrc < 0) return -1;
  if (lrc == 1) {
    // handled by EvaluateWithCuts: re-parse with the proposed thunks as extra breakers
    g_pending_cuts = cut;
    return -2;
  }
  if (debug_plan_enabled()) PrintPlan(pipeline, &lp.plan);

  const int G = nljit_device_count();
  // A ragged column (a filter result: per-shard counts) that meets a column with another layout but
  // the same length is first re-cut into the regular partition — the reference's gather
  // (compiler.cpp:30-49) always leaves one contiguous array, so such pipelines are legal there.
  {
    Storage *first = NULL;
    bool mixed = false, same_len = true;
    for (int k = 0; k < lp.n_inputs; k++) {
      if (lp.inputs[k].kind != NL_IN_ARRAY) continue;
      Storage *s = lp.input_sources[k]->storage;
      if (!first) { first = s; continue; }
      if (s->cardinality() != first->cardinality()) same_len = false;
      bool same = s->shards.size() == first->shards.size();
      for (size_t g = 0; same && g < s->shards.size(); g++) same = s->shards[g].count == first->shards[g].count;
      if (!same) mixed = true;
    }
    if (mixed && same_len) {
      for (int k = 0; k < lp.n_inputs; k++) {
        if (lp.inputs[k].kind != NL_IN_ARRAY) continue;
        Storage *s = lp.input_sources[k]->storage;
        if (!s->ragged) continue;
        Storage *r = storage_repartition(s);
        if (!r) return -1;
        storage_decref(s);
        lp.input_sources[k]->storage = r;
      }
    }
  }
  // shard geometry: every column of a pipeline shares one partition (compiler.cpp:413, 428)
  Storage *shape = NULL;
  for (int k = 0; k < lp.n_inputs; k++) {
    Storage *s = lp.input_sources[k]->storage;
    if (lp.inputs[k].kind != NL_IN_ARRAY) continue;
    if (!shape) { shape = s; continue; }
    bool same = s->shards.size() == shape->shards.size();
    for (size_t g = 0; same && g < s->shards.size(); g++) same = s->shards[g].count == shape->shards[g].count;
    if (!same) {
      PyErr_SetString(PyExc_TypeError, "Incompatible cardinality.");
      return -1;
    }
  }
  if (!shape) {
    PyErr_SetString(PyExc_ValueError, "a pipeline needs at least one column input");
    return -1;
  }

  // outputs (ScheduleFunction, scheduler.cpp:107-122; no zero fill — the reference's own FIXME)
  std::vector<Storage *> outs(lp.n_outputs, (Storage *)NULL);
  auto fail = [&](void) { for (Storage *s : outs) if (s) storage_decref(s); return -1; };
  for (int k = 0; k < lp.n_outputs; k++) {
    const int space = lp.plan.out[k].index_space;
    if (space == NL_IDX_WHERE) {
      if (!pipeline->inplace_target || !pipeline->inplace_target->storage) {
        PyErr_SetString(PyExc_ValueError, "masked assignment without a target"); return fail();
      }
      Storage *t = pipeline->inplace_target->storage;
      bool same = t->shards.size() == shape->shards.size();
      for (size_t g = 0; same && g < t->shards.size(); g++) same = t->shards[g].count == shape->shards[g].count;
      if (!same) { PyErr_SetString(PyExc_IndexError, "boolean index did not match the assigned thunk's cardinality"); return fail(); }
      storage_incref(t);
      outs[k] = t;
      continue;
    }
    Storage *s = storage_new(lp.plan.out[k].dtype, space == NL_IDX_SHARD ? 1 : shape->bound);
    outs[k] = s;
    s->ragged = shape->ragged || space == NL_IDX_COMPACT;
    s->replicated = space == NL_IDX_SHARD;
    const size_t es = nl_dtype_size(s->dtype);
    for (int g = 0; g < G; g++) {
      Shard sh;
      sh.dev = g;
      sh.start = shape->shards[g].start;
      sh.capacity = space == NL_IDX_SHARD ? 1 : shape->shards[g].count;
      sh.count = sh.capacity;
      sh.ptr = NULL;
      if (sh.capacity > 0) {
        int rc = nl_malloc(g, (size_t)sh.capacity * es, &sh.ptr);
        if (rc != NL_OK) { nljit_raise(rc, "jit: device allocation failed"); return fail(); }
      }
      s->shards.push_back(sh);
    }
  }

  // one launch per device shard (the reference's 8 range tasks, scheduler.cpp:133-147)
  const bool fused_finish = G > 1 && (lp.plan.flags & NL_PLAN_HAS_REDUCE) && nl_comm_peer_ready();
  // a filter over several GPUs exchanges its shard counts inside the kernel too: every device ends up with
  // the exclusive scan of all counts (the memmove gather of compiler.cpp:36-48 as offsets), so the host reads
  // ONE small vector from ONE device instead of a count from each
  const bool fused_counts = G > 1 && (lp.plan.flags & NL_PLAN_HAS_COMPACT) && !(lp.plan.flags & NL_PLAN_HAS_REDUCE) && nl_comm_peer_ready();
  std::vector<int64_t *> sizes(G, (int64_t *)NULL);
  std::vector<int64_t *> offsets(G, (int64_t *)NULL);
  for (int g = 0; g < G; g++) {
    void *out_ptrs[NL_MAX_OUTPUTS] = {0};
    const void *in_ptrs[NL_MAX_INPUTS] = {0};
    for (int k = 0; k < lp.n_outputs; k++) {
      out_ptrs[k] = outs[k]->shards[g].ptr;
      if (!out_ptrs[k]) out_ptrs[k] = (void *)16;     // empty shard: never dereferenced
    }
    for (int k = 0; k < lp.n_inputs; k++) {
      Storage *s = lp.input_sources[k]->storage;
      if (lp.inputs[k].kind == NL_IN_SCALAR_IMM) continue;
      in_ptrs[k] = s->shards[g].ptr;
      if (!in_ptrs[k]) in_ptrs[k] = (const void *)16;
    }
    sizes[g] = sizes_buffer(g);
    if (!sizes[g]) { nljit_raise(NL_ERR_NOMEM, "jit: scratch allocation failed"); return fail(); }
    // Inputs whose upload is still in flight (jit.thunk() returns before PCIe is done): the kernel runs
    // chunk by chunk, each launch waiting for the upload of its own chunk only, and leaves marks on the
    // outputs so that asnumpyarray() can send chunk i home while chunk i + 1 is computed.
    std::vector<int64_t> bounds;
    for (int k = 0; k < lp.n_inputs; k++) {
      Storage *s = lp.input_sources[k]->storage;
      if (lp.inputs[k].kind != NL_IN_ARRAY || s->producer != NL_STREAM_H2D) continue;
      std::vector<ReadyMark> &marks = s->shards[g].marks;
      if (!marks.empty() && nl_event_query(marks.back().ev) == 1) continue;      // landed meanwhile
      for (const ReadyMark &m : marks) bounds.push_back(m.end);
    }
    const int64_t count = shape->shards[g].count;
    if (!bounds.empty()) {
      std::sort(bounds.begin(), bounds.end());
      bounds.erase(std::unique(bounds.begin(), bounds.end()), bounds.end());
      while (!bounds.empty() && bounds.back() >= count) bounds.pop_back();
      bounds.push_back(count);
      int64_t prev = 0;
      int rc = NL_OK;
      for (int64_t e : bounds) {
        if (e <= prev) continue;
        for (int k = 0; k < lp.n_inputs && rc == NL_OK; k++) {
          Storage *s = lp.input_sources[k]->storage;
          if (lp.inputs[k].kind != NL_IN_ARRAY || s->producer != NL_STREAM_H2D) continue;
          for (const ReadyMark &m : s->shards[g].marks)
            if (m.end >= e || &m == &s->shards[g].marks.back()) { rc = nl_stream_wait_event(g, NL_STREAM_COMPUTE, m.ev); break; }
        }
        if (rc == NL_OK) rc = nl_launch_pipeline(&lp.plan, out_ptrs, in_ptrs, prev, e, sizes[g], 0, g, NL_STREAM_COMPUTE);
        for (int k = 0; k < lp.n_outputs && rc == NL_OK; k++) {
          nl_event ev = NULL;
          rc = nl_event_create_sync(g, &ev);
          if (rc == NL_OK) rc = nl_event_record(ev, NL_STREAM_COMPUTE);
          if (ev) outs[k]->shards[g].marks.push_back(ReadyMark{e, ev});
          outs[k]->producer = NL_STREAM_COMPUTE;
        }
        if (rc != NL_OK) { nljit_raise(rc, "jit: pipeline launch failed"); return fail(); }
        prev = e;
      }
      continue;
    }
    for (int k = 0; k < lp.n_inputs; k++) {
      // the upload landed before we looked (or there never was one): plain stream order from here on
      Storage *s = lp.input_sources[k]->storage;
      if (lp.inputs[k].kind == NL_IN_ARRAY && s->producer == NL_STREAM_H2D && g == G - 1) storage_drop_marks(s);
    }
    // a reduction over several GPUs finishes inside its kernel: the partials travel through the
    // peers' NVLink mailboxes, no collective launch follows
    if (fused_counts) {
      offsets[g] = offsets_buffer(g);
      if (!offsets[g]) { nljit_raise(NL_ERR_NOMEM, "jit: scratch allocation failed"); return fail(); }
    }
    int rc = fused_finish   ? nl_launch_pipeline_allreduce(&lp.plan, out_ptrs, in_ptrs, 0, count, sizes[g], g)
             : fused_counts ? nl_launch_pipeline_allgather(&lp.plan, out_ptrs, in_ptrs, 0, count, sizes[g], offsets[g], g)
                            : nl_launch_pipeline(&lp.plan, out_ptrs, in_ptrs, 0, count, sizes[g], 0, g, 0);
    if (rc != NL_OK) { nljit_raise(rc, "jit: pipeline launch failed"); return fail(); }
  }

  // finish (JITFunctionDECREF, compiler.cpp:30-49; llvm_finalize_sum_data, thunk_methods.cpp:133-151)
  bool need_counts = false;
  for (int k = 0; k < lp.n_outputs; k++) {
    const int space = lp.plan.out[k].index_space;
    if (space == NL_IDX_SHARD) {
      void *ptrs[NL_MAX_DEVICES];
      for (int g = 0; g < G; g++) ptrs[g] = outs[k]->shards[g].ptr;
      if (fused_finish) continue;
      int rc = nl_finish_reduce(lp.plan.reduce_op, outs[k]->dtype, ptrs, 0);      // NCCL allreduce over NVLink
      if (rc != NL_OK) { nljit_raise(rc, "jit: reduction finish failed"); return fail(); }
    } else if (space == NL_IDX_COMPACT) {
      need_counts = true;
    }
  }
  if (need_counts && fused_counts) {
    // the kernels exchanged the counts among themselves: device 0 holds the scan of all of them
    int64_t host[NL_MAX_DEVICES + 1] = {0};
    int rc = nl_memcpy_d2h(0, 0, host, offsets[0], sizeof(int64_t) * (size_t)(G + 1));
    if (rc == NL_OK) { Py_BEGIN_ALLOW_THREADS rc = nl_stream_sync(0, 0); Py_END_ALLOW_THREADS }
    if (rc != NL_OK) { nljit_raise(rc, "jit: count read-back failed"); return fail(); }
    for (int k = 0; k < lp.n_outputs; k++)
      if (lp.plan.out[k].index_space == NL_IDX_COMPACT)
        for (int g = 0; g < G; g++) outs[k]->shards[g].count = host[g + 1] - host[g];
  } else if (need_counts) {
    // the shard counts are what the memmove gather of compiler.cpp:36-48 becomes: each shard's
    // run starts at the exclusive scan of the counts before it
    std::vector<int64_t> host((size_t)G * NL_MAX_OUTPUTS, 0);
    for (int g = 0; g < G; g++) {
      int rc = nl_memcpy_d2h(g, 0, &host[(size_t)g * NL_MAX_OUTPUTS], sizes[g], sizeof(int64_t) * (size_t)lp.n_outputs);
      if (rc != NL_OK) { nljit_raise(rc, "jit: count read-back failed"); return fail(); }
    }
    if (sync_devices() < 0) return fail();
    for (int k = 0; k < lp.n_outputs; k++)
      if (lp.plan.out[k].index_space == NL_IDX_COMPACT)
        for (int g = 0; g < G; g++) outs[k]->shards[g].count = host[(size_t)g * NL_MAX_OUTPUTS + k];
  }

  // mark the output thunks evaluated (compiler.cpp:51) and publish the storage to parents
  int k = 0;
  for (Binding &e : pipeline->outputs) {
    if (lp.plan.out[k].index_space != NL_IDX_WHERE) {
      storage_incref(outs[k]);
      e.source->storage = outs[k];
      e.source->size = (ssize_t)outs[k]->cardinality();
      storage_incref(outs[k]);
      PyThunk_MarkEvaluated(e.source->object, outs[k]);
    }
    k++;
  }
  for (Storage *s : outs) storage_decref(s);
  pipeline->evaluated = true;
  return 0;
}

static int execute_tree(Pipeline *pipeline) {
  // children first: a parent is only scheduled once all of its children are evaluated
  // (scheduler.cpp:13-50); stream order makes their results visible without host syncs
  for (Pipeline *c : pipeline->children) {
    int rc = execute_tree(c);
    if (rc != 0) return rc;
  }
  return execute_one(pipeline);
}

// Evaluation is stream-ordered: the launches are enqueued and this returns.  The data is waited for
// where it is read on the host (asnumpyarray / str / device_shards, the counts of a filter) — the
// reference's caller blocks on a semaphore here (thunk.cpp:106) because its results live in host
// memory; ours live behind a stream, and not waiting is what lets the upload, the kernel and the
// download of one expression overlap.
int ScheduleExecution(Pipeline *pipeline) {
  int rc = execute_tree(pipeline);
  if (rc == -2) return -2;
  if (rc < 0) return -1;
  return 0;
}

// parse -> execute, re-parsing with an extra breaker whenever the optimiser reports that an
// expression does not fit one fused kernel
static int EvaluateWithCuts(PyThunkObject *root, PyThunkObject *inplace_target) {
  std::set<PyThunkObject *> cuts;
  for (int attempt = 0; attempt < 32; attempt++) {
    Pipeline *pipeline = ParsePipeline(root, &cuts);
    if (!pipeline) return -1;
    pipeline->inplace_target = inplace_target;
    if (debug_plan_enabled()) PrintPipeline(pipeline);
    int rc = ScheduleExecution(pipeline);
    DestroyPipeline(pipeline);
    if (rc == -2) {
      bool grew = false;
      for (PyThunkObject *t : g_pending_cuts) grew |= cuts.insert(t).second;
      g_pending_cuts.clear();
      if (grew) continue;
      break;
    }
    return rc;
  }
  PyErr_SetString(PyExc_ValueError, "expression could not be split into fused pipelines");
  return -1;
}

int ScheduleThunk(PyThunkObject *thunk) { return EvaluateWithCuts(thunk, NULL); }

int ScheduleInPlace(PyThunkObject *where, PyThunkObject *target) { return EvaluateWithCuts(where, target); }
