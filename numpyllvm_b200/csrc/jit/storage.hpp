// storage.hpp — where a thunk's data lives.  The reference keeps one NumPy array per thunk
// (thunk.hpp:26 `storage`); here the array is sharded contiguously over the process's GPUs
// (the 8 range chunks of scheduler.cpp:133-147 become one shard per device) and a host
// ndarray is only materialised on demand (asnumpyarray / str).
#ifndef NLJIT_STORAGE_H
#define NLJIT_STORAGE_H

#include "pyapi.hpp"
#include <vector>

// Elements [0, end) of a shard are valid once `ev` has fired.  Marks exist while a column is being
// produced chunk by chunk on a stream — the upload of jit.thunk() on the copy stream, or a fused
// pipeline that follows such an upload chunk by chunk — so that the consumer (the kernel over chunk
// i, the download of chunk i) waits for its chunk only and PCIe in, HBM and PCIe out overlap.
struct ReadyMark {
  int64_t end;
  nl_event ev;
};

struct Shard {
  int dev;            // local device index
  void *ptr;          // device buffer of `capacity` elements (nullptr when capacity == 0)
  int64_t start;      // global index of the shard's first slot (in the partition of `cardinality_bound`)
  int64_t capacity;   // slots allocated
  int64_t count;      // valid elements (== capacity unless the shard holds a filter result)
  std::vector<ReadyMark> marks;   // ascending `end`; empty: the data is ordered on the compute stream
};

struct Storage {
  int dtype;                  // nl_dtype
  int64_t bound;              // size the partition was made for (the upper bound for filter results)
  std::vector<Shard> shards;  // one per device, empty for host-only scalars
  bool ragged;                // shard counts are not the partition sizes (filter result)
  bool replicated;            // one element, the same on every device (allreduced reduction result)
  // size-1 thunks made from host data never touch the device: value kept here
  bool host_scalar;
  uint64_t scalar_bits;       // bit pattern in `dtype`
  int refs;
  int producer;               // stream the marks were recorded on: NL_STREAM_H2D (upload in flight),
                              // NL_STREAM_COMPUTE (chunked evaluation) or -1 (no marks)

  int64_t cardinality() const {
    if (host_scalar || replicated) return 1;
    int64_t n = 0;
    for (const Shard &s : shards) n += s.count;
    return n;
  }
};

// runtime (lazy): devices are initialised on first use so `import jit` and graph building
// work without a GPU; anything that needs the device then fails loudly.
int nljit_ensure_runtime();                 // 0 or -1 with a Python exception set
int nljit_device_count();                   // after ensure_runtime
void nljit_raise(int rc, const char *what); // map an nl error code to a Python exception
int nljit_npy_to_nl(int npy_type);          // -1 if outside the kernel library
int nljit_nl_to_npy(int nl_type);

Storage *storage_new(int dtype, int64_t bound);
void storage_incref(Storage *s);
void storage_decref(Storage *s);

enum { NL_STREAM_COMPUTE = 0, NL_STREAM_H2D = 1, NL_STREAM_D2H = 2 };

// allocate device shards for `bound` elements with the engine's partition
int storage_alloc_shards(Storage *s);
// Upload a contiguous host array in chunks on the copy stream and return at once with the chunks'
// readiness marks (copy-on-wrap, thunk.cpp:30, without the wait: a pageable source has been consumed
// when this returns; a page-locked source is read by the DMA engine directly and stays guarded —
// referenced and marked read-only — until its last chunk has landed).
Storage *storage_from_host(PyArrayObject *arr);
// download into a fresh ndarray of exactly cardinality() elements (shards concatenated in order);
// chunk by chunk behind the producer's marks on the download stream; returns when the data is there
PyObject *storage_to_host(Storage *s);
// Order the compute stream behind an upload that is still in flight and forget the marks: what every
// consumer that is not chunk-aware, and every in-place writer, calls first.  0 / -1 (exception set).
int storage_settle(Storage *s);
void storage_drop_marks(Storage *s);
// give sources whose upload has completed back to their owners (restores the WRITEABLE flag);
// block: wait for all of them
void nljit_reap_uploads(bool block);
// every device's compute stream idle (GIL released); 0 / -1 (exception set)
int nljit_sync_devices();
// elements per upload / chunked-evaluation chunk for an item size (a multiple of 1024)
int64_t nljit_chunk_elems(size_t itemsize);
void nljit_set_chunk_bytes(int64_t bytes);     // 0: back to the default (32 MiB or NL_CHUNK_MB)

// pinned host memory pool (page-locking is expensive; blocks are reused)
void *nljit_pinned_alloc(size_t bytes);
void nljit_pinned_free(void *p, size_t bytes);
PyObject *nljit_pinned_ndarray(int npy_type, int64_t n);   // ndarray backed by pinned memory

#endif
