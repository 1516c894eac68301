// device_sort.hpp — the sort() full breaker on the GPUs (the reference runs NumPy's sort on the
// host as the operator's `base` function, thunk_methods.cpp:37-62 / compiler.cpp:93-122).
#ifndef NLJIT_DEVICE_SORT_H
#define NLJIT_DEVICE_SORT_H

#include "storage.hpp"

// A new storage holding the elements of `in` in ascending order (NaNs last), partitioned like any
// regular column of that length.  `in` is left untouched (sort() returns a copy).
// NULL with a Python exception set on failure.
Storage *device_sort(Storage *in);

// A regular (canonical-partition) copy of a ragged storage, e.g. a filter result that has to meet
// an unfiltered column of the same length.  Order is preserved.
Storage *storage_repartition(Storage *in);

// Copy the global element range [lo, hi) of `src` (shards concatenated in order) to `dst` on local
// device `dst_dev` with device-to-device / NVLink peer copies, ordered on dst_dev's compute stream.
// The caller synchronises.  Returns an nl error code.
int storage_copy_range(Storage *src, int64_t lo, int64_t hi, int dst_dev, void *dst);

#endif
