// storage.cpp — device-resident thunk storage: shard allocation, host<->device edges and the
// pinned host pool.  Replaces the reference's NumPy-array storage (thunk.cpp:30 ENSURECOPY,
// scheduler.cpp:114-118 PyArray_ZEROS) with per-device shard buffers behind the C-ABI.
#include "storage.hpp"

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

static int g_runtime_state = 0;   // 0 = not tried, 1 = ok, -1 = failed
static int g_ndev = 0;

void nljit_raise(int rc, const char *what) {
  PyObject *exc = PyExc_RuntimeError;
  if (rc == NL_ERR_INVALID || rc == NL_ERR_UNSUPPORTED || rc == NL_ERR_SPLIT || rc == NL_ERR_SPLIT_FILTER) exc = PyExc_ValueError;
  if (rc == NL_ERR_NOMEM) exc = PyExc_MemoryError;
  PyErr_Format(exc, "%s: %s", what, nl_last_error());
}

int nljit_ensure_runtime() {
  if (g_runtime_state == 1) return 0;
  int want = 0;
  if (const char *env = getenv("NL_DEVICES")) want = atoi(env);
  int rc = nl_init(want, nullptr);
  if (rc != NL_OK) {
    nljit_raise(rc, "jit: cannot initialise the CUDA runtime");
    return -1;
  }
  g_ndev = nl_device_count();
  if (g_ndev > 1 && nl_comm_world_size() == 0) {
    // one process driving every GPU of the box: a local NCCL communicator over NVLink
    unsigned char id[NL_UNIQUE_ID_BYTES];
    rc = nl_comm_unique_id(id);
    if (rc == NL_OK) rc = nl_comm_init(id, g_ndev, 0);
    if (rc != NL_OK) {
      nljit_raise(rc, "jit: cannot create the NCCL communicator");
      return -1;
    }
  }
  g_runtime_state = 1;
  return 0;
}

int nljit_device_count() { return g_ndev; }

int nljit_npy_to_nl(int t) {
  switch (t) {
    case NPY_BOOL: return NL_BOOL;
    case NPY_INT32: return NL_I32;
    case NPY_INT64: return NL_I64;
#if NPY_LONGLONG != NPY_INT64
    case NPY_LONGLONG: return NL_I64;
#endif
    case NPY_FLOAT32: return NL_F32;
    case NPY_FLOAT64: return NL_F64;
    case NPY_INT8: return NL_I8;
    case NPY_INT16: return NL_I16;
    case NPY_UINT8: return NL_U8;
    case NPY_UINT16: return NL_U16;
    case NPY_UINT32: return NL_U32;
    case NPY_UINT64: return NL_U64;
#if NPY_ULONGLONG != NPY_UINT64
    case NPY_ULONGLONG: return NL_U64;
#endif
  }
  return -1;
}

int nljit_nl_to_npy(int t) {
  switch (t) {
    case NL_BOOL: return NPY_BOOL;
    case NL_I32: return NPY_INT32;
    case NL_I64: return NPY_INT64;
    case NL_F32: return NPY_FLOAT32;
    case NL_F64: return NPY_FLOAT64;
    case NL_I8: return NPY_INT8;
    case NL_I16: return NPY_INT16;
    case NL_U8: return NPY_UINT8;
    case NL_U16: return NPY_UINT16;
    case NL_U32: return NPY_UINT32;
    case NL_U64: return NPY_UINT64;
  }
  return NPY_NOTYPE;
}

Storage *storage_new(int dtype, int64_t bound) {
  Storage *s = new Storage();
  s->dtype = dtype;
  s->bound = bound;
  s->ragged = false;
  s->replicated = false;
  s->host_scalar = false;
  s->scalar_bits = 0;
  s->refs = 1;
  s->producer = -1;
  return s;
}

void storage_incref(Storage *s) { if (s) s->refs++; }

void storage_drop_marks(Storage *s) {
  for (Shard &sh : s->shards) {
    for (ReadyMark &m : sh.marks) nl_event_destroy(m.ev);
    sh.marks.clear();
  }
  s->producer = -1;
}

static int settle_rc(Storage *s) {
  int rc = NL_OK;
  if (s->producer == NL_STREAM_H2D)
    for (Shard &sh : s->shards)
      if (!sh.marks.empty() && rc == NL_OK) rc = nl_stream_wait_event(sh.dev, NL_STREAM_COMPUTE, sh.marks.back().ev);
  storage_drop_marks(s);
  return rc;
}

int storage_settle(Storage *s) {
  if (!s || s->producer < 0) return 0;
  int rc = settle_rc(s);
  if (rc != NL_OK) { nljit_raise(rc, "jit: stream ordering failed"); return -1; }
  return 0;
}

void storage_decref(Storage *s) {
  if (!s) return;
  if (--s->refs > 0) return;
  // buffers go back to the pool in compute-stream order: an upload still in flight must land first
  settle_rc(s);
  for (Shard &sh : s->shards)
    if (sh.ptr) nl_free(sh.dev, sh.ptr);
  delete s;
}

int storage_alloc_shards(Storage *s) {
  if (nljit_ensure_runtime() < 0) return -1;
  const int G = g_ndev;
  const size_t es = nl_dtype_size(s->dtype);
  s->shards.clear();
  for (int g = 0; g < G; g++) {
    int64_t st = 0, en = 0;
    // interior boundaries at multiples of 1024 elements keep every shard 128-bit aligned
    nl_partition(s->bound, G, 1024, g, &st, &en);
    Shard sh;
    sh.dev = g;
    sh.start = st;
    sh.capacity = en - st;
    sh.count = en - st;
    sh.ptr = nullptr;
    if (sh.capacity > 0) {
      int rc = nl_malloc(g, (size_t)sh.capacity * es, &sh.ptr);
      if (rc != NL_OK) {
        nljit_raise(rc, "jit: device allocation failed");
        for (Shard &o : s->shards) if (o.ptr) nl_free(o.dev, o.ptr);
        s->shards.clear();
        return -1;
      }
    }
    s->shards.push_back(sh);
  }
  return 0;
}

static int sync_all(int stream) {
  int rc = NL_OK;
  Py_BEGIN_ALLOW_THREADS
  for (int g = 0; g < g_ndev && rc == NL_OK; g++) rc = nl_stream_sync(g, stream);
  Py_END_ALLOW_THREADS
  return rc;
}

int nljit_sync_devices() {
  int rc = sync_all(NL_STREAM_COMPUTE);
  if (rc != NL_OK) { nljit_raise(rc, "jit: device synchronisation failed"); return -1; }
  nljit_reap_uploads(false);
  return 0;
}

static int64_t g_chunk_bytes = 0;

void nljit_set_chunk_bytes(int64_t b) { g_chunk_bytes = b > 0 ? b : 0; }

int64_t nljit_chunk_elems(size_t itemsize) {
  int64_t &bytes = g_chunk_bytes;
  if (!bytes) {
    bytes = (int64_t)32 << 20;            // 32 MiB: 0.6 ms of PCIe Gen5, 5 us of HBM
    if (const char *env = getenv("NL_CHUNK_MB")) { const long mb = atol(env); if (mb > 0) bytes = (int64_t)mb << 20; }
  }
  int64_t n = (bytes / (int64_t)(itemsize ? itemsize : 1)) & ~(int64_t)1023;
  return n < 1024 ? 1024 : n;
}

// ---- uploads in flight from page-locked sources ------------------------------------------
// The DMA engine reads such a source directly while jit.thunk() has already returned.  Until the
// last chunk has landed the source ndarray is kept alive and marked read-only (NumPy refuses writes
// through it), which is what keeps the reference's copy-on-wrap contract (thunk.cpp:30) observable.
struct Inflight {
  PyObject *source;
  bool restore_writeable;
  std::vector<nl_event> done;     // one per device that received a piece, recorded after its last chunk
};
static std::vector<Inflight> g_inflight;

void nljit_reap_uploads(bool block) {
  size_t w = 0;
  for (size_t i = 0; i < g_inflight.size(); i++) {
    Inflight &f = g_inflight[i];
    bool complete = true;
    for (nl_event ev : f.done) {
      if (block) { Py_BEGIN_ALLOW_THREADS nl_event_sync(ev); Py_END_ALLOW_THREADS }
      else if (nl_event_query(ev) == 0) { complete = false; break; }
    }
    if (!complete) { if (w != i) g_inflight[w] = f; w++; continue; }
    for (nl_event ev : f.done) nl_event_destroy(ev);
    if (f.restore_writeable) {
      // the same array may be on its way up a second time: the last upload to finish lifts the guard
      bool handed_on = false;
      for (size_t j = 0; j < g_inflight.size() && !handed_on; j++) {
        const bool alive = (j < i) ? (j < w) : (j > i);      // kept entries [0, w) and the not yet visited ones
        if (alive && g_inflight[j].source == f.source) { g_inflight[j].restore_writeable = true; handed_on = true; }
      }
      if (!handed_on) PyArray_ENABLEFLAGS((PyArrayObject *)f.source, NPY_ARRAY_WRITEABLE);
    }
    Py_DECREF(f.source);
  }
  g_inflight.resize(w);
}

Storage *storage_from_host(PyArrayObject *arr) {
  nljit_reap_uploads(false);
  const int dtype = nljit_npy_to_nl(PyArray_TYPE(arr));
  const int64_t n = (int64_t)PyArray_SIZE(arr);
  Storage *s = storage_new(dtype, n);
  if (storage_alloc_shards(s) < 0) { storage_decref(s); return nullptr; }
  const size_t es = nl_dtype_size(dtype);
  const char *src = (const char *)PyArray_DATA(arr);
  const bool pinned = nl_host_is_pinned(src) == 1;
  const int64_t chunk = nljit_chunk_elems(es);
  Inflight guard;
  guard.source = (PyObject *)arr;
  guard.restore_writeable = false;
  int rc = NL_OK;
  s->producer = NL_STREAM_H2D;
  Py_BEGIN_ALLOW_THREADS            // a pageable source is staged by the driver inside the call: do not hold the GIL for it
  for (Shard &sh : s->shards) {
    if (!sh.capacity || rc != NL_OK) continue;
    // the block may be recycled pool memory that work already enqueued on the compute stream still uses
    rc = nl_stream_wait(sh.dev, NL_STREAM_H2D, sh.dev, NL_STREAM_COMPUTE);
    for (int64_t off = 0; off < sh.capacity && rc == NL_OK; off += chunk) {
      const int64_t m = sh.capacity - off < chunk ? sh.capacity - off : chunk;
      rc = nl_memcpy_h2d(sh.dev, NL_STREAM_H2D, (char *)sh.ptr + (size_t)off * es, src + (size_t)(sh.start + off) * es, (size_t)m * es);
      nl_event ev = nullptr;
      if (rc == NL_OK) rc = nl_event_create_sync(sh.dev, &ev);
      if (rc == NL_OK) rc = nl_event_record(ev, NL_STREAM_H2D);
      if (ev) sh.marks.push_back(ReadyMark{off + m, ev});
    }
    if (pinned && rc == NL_OK) {
      nl_event ev = nullptr;
      rc = nl_event_create_sync(sh.dev, &ev);
      if (rc == NL_OK) rc = nl_event_record(ev, NL_STREAM_H2D);
      if (ev) guard.done.push_back(ev);
    }
  }
  Py_END_ALLOW_THREADS
  if (rc != NL_OK) {
    nljit_raise(rc, "jit: host to device copy failed");
    for (nl_event ev : guard.done) { nl_event_sync(ev); nl_event_destroy(ev); }
    sync_all(NL_STREAM_H2D);
    storage_decref(s);
    return nullptr;
  }
  if (pinned && !guard.done.empty()) {
    // copy-on-wrap (thunk.cpp:30) without waiting for PCIe: the source is guarded instead of copied
    if (PyArray_FLAGS(arr) & NPY_ARRAY_WRITEABLE) {
      PyArray_CLEARFLAGS(arr, NPY_ARRAY_WRITEABLE);
      guard.restore_writeable = true;
    }
    Py_INCREF(guard.source);
    g_inflight.push_back(guard);
  }
  return s;
}

// ---- pinned pool -----------------------------------------------------------------------
static std::mutex g_pool_mu;
static std::multimap<size_t, void *> g_pool;
static size_t g_pool_bytes = 0;
static const size_t kPoolCap = (size_t)24 << 30;

static size_t round_block(size_t bytes) {
  const size_t gran = (size_t)2 << 20;
  return ((bytes ? bytes : 1) + gran - 1) / gran * gran;
}

void *nljit_pinned_alloc(size_t bytes) {
  const size_t blk = round_block(bytes);
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    auto it = g_pool.find(blk);
    if (it != g_pool.end()) {
      void *p = it->second;
      g_pool.erase(it);
      g_pool_bytes -= blk;
      return p;
    }
  }
  void *p = nullptr;
  if (nl_host_alloc(blk, &p) != NL_OK) return nullptr;
  return p;
}

void nljit_pinned_free(void *p, size_t bytes) {
  if (!p) return;
  const size_t blk = round_block(bytes);
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (g_pool_bytes + blk <= kPoolCap) {
      g_pool.insert({blk, p});
      g_pool_bytes += blk;
      return;
    }
  }
  nl_host_free(p);
}

struct PinnedBlock { void *ptr; size_t bytes; };

static void pinned_capsule_free(PyObject *cap) {
  PinnedBlock *b = (PinnedBlock *)PyCapsule_GetPointer(cap, "nljit.pinned");
  if (b) {
    nljit_pinned_free(b->ptr, b->bytes);
    delete b;
  }
}

PyObject *nljit_pinned_ndarray(int npy_type, int64_t n) {
  if (nljit_ensure_runtime() < 0) return nullptr;
  PyArray_Descr *descr = PyArray_DescrFromType(npy_type);
  if (!descr) return nullptr;
  const size_t bytes = (size_t)n * (size_t)PyDataType_ELSIZE(descr);
  Py_DECREF(descr);
  void *p = nljit_pinned_alloc(bytes);
  if (!p) { nljit_raise(NL_ERR_NOMEM, "jit: pinned host allocation failed"); return nullptr; }
  npy_intp dims[1] = {(npy_intp)n};
  PyObject *arr = PyArray_SimpleNewFromData(1, dims, npy_type, p);
  if (!arr) { nljit_pinned_free(p, bytes); return nullptr; }
  PinnedBlock *b = new PinnedBlock{p, bytes};
  PyObject *cap = PyCapsule_New(b, "nljit.pinned", pinned_capsule_free);
  if (!cap || PyArray_SetBaseObject((PyArrayObject *)arr, cap) < 0) {
    Py_XDECREF(cap);
    Py_DECREF(arr);
    return nullptr;
  }
  return arr;
}

PyObject *storage_to_host(Storage *s) {
  const int npy = nljit_nl_to_npy(s->dtype);
  if (s->host_scalar) {
    npy_intp dims[1] = {1};
    PyObject *arr = PyArray_SimpleNew(1, dims, npy);
    if (!arr) return nullptr;
    memcpy(PyArray_DATA((PyArrayObject *)arr), &s->scalar_bits, nl_dtype_size(s->dtype));
    return arr;
  }
  const int64_t n = s->cardinality();
  const size_t es = nl_dtype_size(s->dtype);
  if (s->replicated) {
    npy_intp dims[1] = {1};
    PyObject *one = PyArray_SimpleNew(1, dims, npy);
    if (!one) return nullptr;
    int rc1 = nl_memcpy_d2h(s->shards[0].dev, 0, PyArray_DATA((PyArrayObject *)one), s->shards[0].ptr, es);
    if (rc1 == NL_OK) rc1 = sync_all(0);
    if (rc1 != NL_OK) { nljit_raise(rc1, "jit: device to host copy failed"); Py_DECREF(one); return nullptr; }
    return one;
  }
  PyObject *arr;
  if ((size_t)n * es >= ((size_t)1 << 20)) {
    arr = nljit_pinned_ndarray(npy, n);      // DMA target: full PCIe rate
  } else {
    npy_intp dims[1] = {(npy_intp)n};
    arr = PyArray_SimpleNew(1, dims, npy);
  }
  if (!arr) return nullptr;
  char *dst = (char *)PyArray_DATA((PyArrayObject *)arr);
  int64_t off = 0;
  int rc = NL_OK;
  for (Shard &sh : s->shards) {
    if (sh.count > 0 && rc == NL_OK) {
      // the exclusive scan of the shard counts is where each shard's run lands (compiler.cpp:36-48)
      char *to = dst + (size_t)off * es;
      if (sh.marks.empty()) {
        rc = nl_stream_wait(sh.dev, NL_STREAM_D2H, sh.dev, NL_STREAM_COMPUTE);
        if (rc == NL_OK) rc = nl_memcpy_d2h(sh.dev, NL_STREAM_D2H, to, sh.ptr, (size_t)sh.count * es);
      } else {
        // the column is still being produced chunk by chunk: each chunk leaves as soon as it is there
        int64_t prev = 0;
        for (const ReadyMark &m : sh.marks) {
          const int64_t end = m.end < sh.count ? m.end : sh.count;
          if (end <= prev || rc != NL_OK) continue;
          rc = nl_stream_wait_event(sh.dev, NL_STREAM_D2H, m.ev);
          if (rc == NL_OK) rc = nl_memcpy_d2h(sh.dev, NL_STREAM_D2H, to + (size_t)prev * es, (const char *)sh.ptr + (size_t)prev * es, (size_t)(end - prev) * es);
          prev = end;
        }
      }
    }
    off += sh.count;
  }
  if (rc == NL_OK) rc = sync_all(NL_STREAM_D2H);
  if (rc != NL_OK) { nljit_raise(rc, "jit: device to host copy failed"); Py_DECREF(arr); return nullptr; }
  storage_drop_marks(s);          // every chunk has been waited for and copied: the marks have all fired
  nljit_reap_uploads(false);
  return arr;
}
