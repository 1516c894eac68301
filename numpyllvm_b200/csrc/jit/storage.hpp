// storage.hpp — where a thunk's data lives.  The reference keeps one NumPy array per thunk
// (thunk.hpp:26 `storage`); here the array is sharded contiguously over the process's GPUs
// (the 8 range chunks of scheduler.cpp:133-147 become one shard per device) and a host
// ndarray is only materialised on demand (asnumpyarray / str).
#ifndef NLJIT_STORAGE_H
#define NLJIT_STORAGE_H

#include "pyapi.hpp"
#include <vector>

struct Shard {
  int dev;            // local device index
  void *ptr;          // device buffer of `capacity` elements (nullptr when capacity == 0)
  int64_t start;      // global index of the shard's first slot (in the partition of `cardinality_bound`)
  int64_t capacity;   // slots allocated
  int64_t count;      // valid elements (== capacity unless the shard holds a filter result)
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

// allocate device shards for `bound` elements with the engine's partition
int storage_alloc_shards(Storage *s);
// upload a contiguous host array (blocks until the copy is complete: copy-on-wrap, thunk.cpp:30)
Storage *storage_from_host(PyArrayObject *arr);
// download into a fresh ndarray of exactly cardinality() elements (shards concatenated in order)
PyObject *storage_to_host(Storage *s);

// pinned host memory pool (page-locking is expensive; blocks are reused)
void *nljit_pinned_alloc(size_t bytes);
void nljit_pinned_free(void *p, size_t bytes);
PyObject *nljit_pinned_ndarray(int npy_type, int64_t n);   // ndarray backed by pinned memory

#endif
